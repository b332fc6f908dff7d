"""Continuous covariates (> 256 distinct values): python tools/continuous_check.py <rows> <covariates> <ntree> <nodesize>"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import rfslam_b200
from rfslam_b200.synth import make_cpiu
n, p, ntree, nodesize = [int(v) for v in sys.argv[1:5]]
d = make_cpiu(n, p, levels=0, seed=3)
ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
for rep in range(2):
    t = time.time(); f = ds.grow(ntree=ntree, nodesize=nodesize, seed=-5); dt = time.time() - t
    print(f"continuous n={n} p={p} ntree={ntree}: wall {dt:.3f}s kernel {f['grow_kernel_ms']:.1f} ms, nodes {f['treeID'].size}, cand {f['split_candidates']:.3e}", flush=True)
