"""world_size-2 gloo test of the N > 1 host logic (tree sharding, ensemble all-reduce, forest gather).
Each rank stands in for a GPU shard with per-tree pieces cut out of an oracle forest; the merged
result must equal the unsharded forest."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from oracle import port
from rfslam_b200 import shard
from rfslam_b200.synth import make_cpiu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _shard_of(full, n, ntree, ids):
    """What rfslam_b200_grow(tree_first, tree_step, finalize_ensembles = 0) returns for these trees."""
    noff = np.concatenate([[0], np.cumsum(2 * full["leafCount"].astype(np.int64) - 1)])
    loff = np.concatenate([[0], np.cumsum(full["leafCount"].astype(np.int64))])
    f = {k: np.concatenate([full[k][noff[b - 1]:noff[b]] for b in ids]) for k in shard.FOREST_KEYS}
    for k in shard.LEAF_KEYS:
        f[k] = np.concatenate([full[k][loff[b - 1]:loff[b]] for b in ids])
    f["tnCLAS"] = np.concatenate([full["tnCLAS"][2 * loff[b - 1]:2 * loff[b]] for b in ids])
    f["leafCount"] = full["leafCount"][ids - 1]; f["seed"] = full["seed"][ids - 1]; f["tree_ids"] = ids
    memb = full["nodeMembership"].reshape(ntree, n); boot = full["bootMembership"].reshape(ntree, n)
    all_num = np.zeros((2, n)); oob_num = np.zeros((2, n)); all_den = np.zeros(n, np.uint32); oob_den = np.zeros(n, np.uint32)
    for b in ids:
        clas = full["tnCLAS"][2 * loff[b - 1]:2 * loff[b]].reshape(-1, 2).astype(np.float64)
        prob = clas / clas.sum(axis=1, keepdims=True)
        pr = prob[memb[b - 1] - 1].T
        all_num += pr; all_den += 1
        oob = boot[b - 1] == 0
        oob_num[:, oob] += pr[:, oob]; oob_den[oob] += 1
    f["allEnsbCLS"] = all_num.reshape(-1); f["oobEnsbCLS"] = oob_num.reshape(-1); f["allEnsbDen"] = all_den; f["oobEnsbDen"] = oob_den
    return f


def _worker(rank, world, portno, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(portno)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, ntree = 1500, 7
    d = make_cpiu(n, 8, levels=32, seed=21)
    full = port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=ntree, nodesize=5, seed=-77)
    ids = shard.shard_tree_ids(rank, world, ntree)
    part = _shard_of(full, n, ntree, ids)
    merged = shard.allreduce_ensembles(part)
    forest = shard.gather_forest(part, dst=0)
    ok = np.allclose(merged["allEnsbCLS"], full["allEnsbCLS"], rtol=1e-12) and np.allclose(merged["oobEnsbCLS"], full["oobEnsbCLS"], rtol=1e-12)
    # the forest's perfClas from the reduced ensemble = the oracle's (and the reference's) on the unsharded forest
    perf = shard.perf_clas(d.y, merged["oobEnsbCLS"], merged["oobEnsbDen"], ntree)
    ok = ok and np.array_equal(perf.view(np.uint64), full["perfClas"].view(np.uint64))
    if rank == 0:
        for k in shard.FOREST_KEYS + shard.LEAF_KEYS + ("leafCount", "seed"):
            a, b = forest[k], full[k]
            ok = ok and a.shape == b.shape and np.array_equal(np.nan_to_num(a, nan=-1.0), np.nan_to_num(b, nan=-1.0))
        ok = ok and np.array_equal(forest["tree_ids"], np.arange(1, ntree + 1))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_merge_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    portno = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, portno, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_ids_cover_forest_once():
    for world in (1, 2, 3, 8):
        ids = np.concatenate([shard.shard_tree_ids(r, world, 13) for r in range(world)])
        assert sorted(ids.tolist()) == list(range(1, 14))
