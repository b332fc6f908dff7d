"""Kernel time against tree count: python tools/sweep_trees.py <rows> <covariates> <levels> <nodesize> <ntree> [<ntree> ...]"""
import os, time, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import rfslam_b200
from rfslam_b200.synth import make_cpiu
n, p, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
nodesize = int(sys.argv[4])
d = make_cpiu(n, p, levels=L)
print("K", int(d.strat.max()), flush=True)
ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
for ntree in [int(a) for a in sys.argv[5:]]:
    best = 1e30
    for rep in range(2):
        f = ds.grow(ntree=ntree, nodesize=nodesize, seed=-12345)
        best = min(best, f['grow_kernel_ms'])
    print(f"ntree {ntree} kernel {best:.1f} ms  trees/s {ntree/best*1e3:.1f}", flush=True)
