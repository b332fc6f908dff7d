"""Tree sharding across GPUs: one process per GPU, rank r grows trees
r+1, r+1+G, ... of the SAME ntree-tree forest (chain seeds are derived for all ntree trees on
every rank, randomForestSRC.c:6664-6692, so tree b is identical for any G).  Nothing is exchanged
while trees grow; afterwards the ensemble numerators / denominators are summed
(randomForestSRC.c:944-949 does the same under per-row locks) and, if wanted, the variable-length
forest records are gathered to rank 0.  Works with any torch.distributed backend (NCCL on the
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

FOREST_KEYS = ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ")
LEAF_KEYS = ("tnRCNT", "tnACNT")


class Comm:
    """The library's own NCCL communicator (rfslam_b200_comm_init, csrc/comm.cu).  With it `rfsrc(..., comm=c)` /
    `ResidentCpiu.grow(..., comm=c)` all-reduce the ensemble sums on the DEVICE before they are finalized and sent
    home, and `gather_forest=True` gathers the shards' records to rank 0 over NVLink.  Rank 0 makes the 128-byte
    id; it travels through whatever the program has -- here any torch.distributed backend (gloo is enough)."""

    def __init__(self, rank: int, world: int, device: int, unique_id: bytes):
        import ctypes as C
        from . import capi
        self.rank, self.world, self.device = rank, world, device
        self.handle = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        capi.check(capi.lib().rfslam_b200_comm_init(C.byref(self.handle), rank, world, buf, device))

    @staticmethod
    def make_unique_id() -> bytes:
        import ctypes as C
        from . import capi
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().rfslam_b200_comm_unique_id(buf, 128))
        return buf.raw

    @classmethod
    def from_torch(cls, device: int, group=None) -> "Comm":
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.make_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(rank, world, device, box[0])

    def collectives(self):
        import ctypes as C
        from . import capi
        a, g = C.c_int64(0), C.c_int64(0)
        capi.lib().rfslam_b200_comm_collectives(self.handle, C.byref(a), C.byref(g))
        return {"allreduce": int(a.value), "gather": int(g.value)}

    def close(self):
        from . import capi
        if self.handle:
            capi.lib().rfslam_b200_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_tree_ids(rank: int, world: int, ntree: int) -> np.ndarray:
    """1-based ids of the trees rank `rank` grows: tree b -> rank (b - 1) mod world."""
    return np.arange(rank + 1, ntree + 1, world, dtype=np.int32)


def allreduce_ensembles(f: Dict[str, np.ndarray], *, device=None, group=None) -> Dict[str, np.ndarray]:
    """HOST-side form of the exchange (any torch.distributed backend; the CPU tests run it over gloo): sum the
    per-shard ensemble sums over ranks and finish them (randomForestSRC.c:8113-8125: numerator / denominator where
    the denominator is non-zero, 0 otherwise).  On GPUs pass a `Comm` to the grow call instead: the same sum is
    then taken on the device accumulators inside the library, with no host bounce."""
    import torch
    import torch.distributed as dist
    if f.get("finalized", False):
        raise ValueError("allreduce_ensembles needs the shards' SUMS: grow them with finalize=False")
    n = f["allEnsbDen"].size
    dev = device if device is not None else "cpu"
    # the arrays are page-locked (hostpool.cu): four straight copies, everything else on the device
    num = torch.cat([torch.from_numpy(np.ascontiguousarray(f[k], np.float64)).to(dev, non_blocking=True) for k in ("oobEnsbCLS", "allEnsbCLS")])
    den = torch.cat([torch.from_numpy(np.ascontiguousarray(f[k]).view(np.int32)).to(dev, non_blocking=True) for k in ("oobEnsbDen", "allEnsbDen")])
    dist.all_reduce(num, group=group)
    dist.all_reduce(den, group=group)
    out = dict(f)
    for k, (lo, dlo) in {"oobEnsbCLS": (0, 0), "allEnsbCLS": (2 * n, n)}.items():
        d = den[dlo:dlo + n].to(torch.float64)
        v = num[lo:lo + 2 * n].view(2, n)
        out[k] = torch.where(d > 0, v / torch.where(d > 0, d, torch.ones_like(d)), torch.zeros_like(v)).reshape(-1).cpu().numpy()
    den_h = den.cpu().numpy().view(np.uint32)
    out["oobEnsbDen"] = den_h[:n].copy()
    out["allEnsbDen"] = den_h[n:].copy()
    return out


def perf_clas(y: np.ndarray, oob_ens: np.ndarray, oob_den: np.ndarray, ntree: int) -> np.ndarray:
    """perfClas of a whole forest from its (all-reduced, finalized) OOB ensemble: what rfslam_b200_grow returns
    when it grows every tree itself (perf_kernel in grow.cu).  [ntree][3], NA_REAL except the last row =
    misclassification rate over rows with an OOB prediction (randomForestSRC.c:1067-1095, votes as in
    getMaxVote randomForestSRC.c:1175-1218: class 2 wins a tie) and per class over ALL rows of the class
    (randomForestSRC.c:1025-1066)."""
    na = np.array([0x7FF00000000007A2], np.uint64).view(np.float64)[0]
    out = np.full(ntree * 3, na, np.float64)
    n = oob_den.size
    p = np.asarray(oob_ens, np.float64).reshape(2, n)
    cls = np.asarray(y).astype(np.int64)                     # 1 / 2
    has = np.asarray(oob_den) != 0
    vote = np.where(p[0] <= p[1], 2, 1)
    correct = has & (vote == cls)
    if has.any():
        out[-3] = 1.0 - float(correct.sum()) / float(has.sum())
        for k in (1, 2):
            cnt = int((cls == k).sum())
            if cnt:
                out[-3 + k] = 1.0 - float((correct & (cls == k)).sum()) / float(cnt)
    return out


def gather_forest(f: Dict[str, np.ndarray], *, dst: int = 0, group=None) -> Optional[Dict[str, np.ndarray]]:
    """Gather the shards' forest records to rank `dst` and put the trees in ascending treeID order
    (records of one tree stay contiguous and in pre-order, randomForestSRC.c:10735-10806)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    payload = {k: f[k] for k in FOREST_KEYS + LEAF_KEYS + ("tnCLAS", "leafCount", "seed", "tree_ids")}
    gathered: List[Optional[dict]] = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    return merge_shards(gathered)


def merge_shards(parts: List[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    trees = []
    for p in parts:
        noff = np.concatenate([[0], np.cumsum(2 * p["leafCount"].astype(np.int64) - 1)])
        loff = np.concatenate([[0], np.cumsum(p["leafCount"].astype(np.int64))])
        for t, b in enumerate(p["tree_ids"]):
            trees.append((int(b), p, t, noff, loff))
    trees.sort(key=lambda x: x[0])
    out: Dict[str, List[np.ndarray]] = {k: [] for k in FOREST_KEYS + LEAF_KEYS + ("tnCLAS", "leafCount", "seed", "tree_ids")}
    for b, p, t, noff, loff in trees:
        for k in FOREST_KEYS:
            out[k].append(p[k][noff[t]:noff[t + 1]])
        for k in LEAF_KEYS:
            out[k].append(p[k][loff[t]:loff[t + 1]])
        out["tnCLAS"].append(p["tnCLAS"][2 * loff[t]:2 * loff[t + 1]])
        out["leafCount"].append(p["leafCount"][t:t + 1]); out["seed"].append(p["seed"][t:t + 1])
        out["tree_ids"].append(np.array([b], np.int32))
    return {k: np.concatenate(v) if v else np.zeros(0) for k, v in out.items()}
