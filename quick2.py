import sys, numpy as np
sys.path.insert(0, '.')
import rfslam_b200
from rfslam_b200.synth import make_cpiu
d = make_cpiu(1000000, 50, levels=64)
ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
for ntree in [int(a) for a in sys.argv[1:]]:
    f = ds.grow(ntree=ntree, nodesize=20, seed=-12345)
    print(ntree, f['grow_kernel_ms'], flush=True)
