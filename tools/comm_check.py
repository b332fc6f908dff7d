"""Tree-sharded growth with the library's NCCL communicator, checked against the same forest grown on one GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, torch.distributed as dist
import rfslam_b200
from rfslam_b200 import shard
from rfslam_b200.synth import make_cpiu

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")                        # the id only needs ANY channel; the data plane is the library's own NCCL
comm = shard.Comm.from_torch(local)
d = make_cpiu(30000, 12, levels=64, seed=5)
ntree = 13
part = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=ntree, nodesize=5, seed=-31, device=local, tree_first=rank + 1, tree_step=world,
                         comm=comm, gather_forest=True, split_stat=True)
full = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=ntree, nodesize=5, seed=-31, device=local, split_stat=True)
# every rank: the whole forest's ensembles and perfClas, reduced on the device
for k in ("oobEnsbCLS", "allEnsbCLS"):
    np.testing.assert_allclose(part[k], full[k], rtol=1e-12, atol=0, err_msg=k)
np.testing.assert_array_equal(part["oobEnsbDen"], full["oobEnsbDen"])
np.testing.assert_array_equal(part["perfClas"].view(np.uint64), full["perfClas"].view(np.uint64))
if rank == 0:                                            # rank 0: the gathered records = the single-GPU forest, tree by tree
    for k in ("tree_ids", "treeID", "nodeID", "parmID", "leafCount", "seed", "tnRCNT", "tnACNT", "tnCLAS"):
        np.testing.assert_array_equal(part[k], full[k], err_msg=k)
    np.testing.assert_array_equal(part["contPT"].view(np.uint64), full["contPT"].view(np.uint64))
    np.testing.assert_array_equal(part["splitStat"].view(np.uint64), full["splitStat"].view(np.uint64))
else:
    assert part["treeID"].size == 0 and list(part["tree_ids"]) == list(range(rank + 1, ntree + 1, world))
# records left sharded (gather_forest = False): each rank returns its own trees, ensembles still whole
sh = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=ntree, nodesize=5, seed=-31, device=local, tree_first=rank + 1, tree_step=world, comm=comm)
np.testing.assert_allclose(sh["oobEnsbCLS"], full["oobEnsbCLS"], rtol=1e-12, atol=0)
assert list(sh["tree_ids"]) == list(range(rank + 1, ntree + 1, world)) and sh["treeID"].size == int((2 * sh["leafCount"].astype(np.int64) - 1).sum())
print(f"rank {rank}: ok, collectives {comm.collectives()}", flush=True)
comm.close()
dist.barrier()
dist.destroy_process_group()
