# config[0]-like: ~80k CPIUs x 30 covariates, continuous values (every numeric column has > 256 distinct values)
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import rfslam_b200
from rfslam_b200.synth import make_cpiu
n, p, levels, ntree, nodesize = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
d = make_cpiu(n, p, levels=levels, seed=3)
print("D per column:", [int(np.unique(c).size) for c in d.x][:12], flush=True)
ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
for rep in range(2):
    t = time.time(); f = ds.grow(ntree=ntree, nodesize=nodesize, seed=-5); dt = time.time() - t
    print(f"GPU  n={n} p={p} L={levels} ntree={ntree}: wall {dt:.3f}s kernel {f['grow_kernel_ms']:.1f} ms -> {ntree/dt:.1f} trees/s, nodes {f['treeID'].size}, cand {f['split_candidates']:.3e}", flush=True)
if len(sys.argv) > 6:
    import refdriver as rd
    nt = int(sys.argv[6])
    t = time.time(); r = rd.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=nt, nodesize=nodesize, seed=-5, membership=False); dt = time.time() - t
    print(f"REF  ntree={nt}: {dt:.2f}s -> {nt/dt:.2f} trees/s", flush=True)
