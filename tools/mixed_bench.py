"""Mixed numeric / unordered-factor table (configs[3]'s covariate kinds on configs[1]'s shape): python tools/mixed_bench.py [rows] [trees]"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import rfslam_b200
from rfslam_b200.synth import make_cpiu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000; trees = int(sys.argv[2]) if len(sys.argv) > 2 else 100
d = make_cpiu(n, 50, levels=64); p = 50
rng = np.random.default_rng(4)
xm = d.x.copy(); xt = ["R"] * p; xl = np.zeros(p, np.int32)
for c in range(2, p, 5):
    lv = 3 + (c // 5) % 6
    xm[c] = rng.integers(1, lv + 1, size=d.n).astype(np.float64); xt[c] = "C"; xl[c] = lv
ds = rfslam_b200.ResidentCpiu(xm, d.y, d.rt, d.strat, xvar_types=xt, xvar_levels=xl)
for rep in range(2):
    t0 = time.perf_counter(); f = ds.grow(ntree=trees, nodesize=20, seed=-12345); dt = time.perf_counter() - t0
    print(f"mixed: {trees / dt:.1f} trees/s, kernel {f['grow_kernel_ms']:.0f} ms, factor splits {int(f['mwcpSZ'].sum())}, nodes {f['treeID'].size}", flush=True)
