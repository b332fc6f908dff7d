# time-slicing check: identical forests with and without slicing, at a size where trees outnumber resident CTAs
import sys, os, numpy as np
sys.path.insert(0, '.')
import rfslam_b200
from rfslam_b200.synth import make_cpiu
n, ntree = int(sys.argv[1]), int(sys.argv[2])
d = make_cpiu(n, 12, levels=64, seed=5)
ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
f1 = ds.grow(ntree=ntree, nodesize=5, seed=-77)
print("sliced kernel ms", f1['grow_kernel_ms'], flush=True)
os.environ["RFSLAM_NO_SLICING"] = "1"
f2 = ds.grow(ntree=ntree, nodesize=5, seed=-77)
print("plain  kernel ms", f2['grow_kernel_ms'], flush=True)
for k in ("treeID", "nodeID", "parmID", "leafCount", "tnRCNT", "tnACNT"):
    assert np.array_equal(f1[k], f2[k]), k
assert np.array_equal(f1["contPT"].view(np.uint64), f2["contPT"].view(np.uint64))
assert np.array_equal(f1["oobEnsbCLS"], f2["oobEnsbCLS"])
print("identical:", f1["treeID"].size, "nodes")
