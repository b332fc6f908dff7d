"""Where a configs[2] step goes: python tools/c3_breakdown.py <trees> [rows] [covariates]  (table generated on the device)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import rfslam_b200
from rfslam_b200.synth import make_cpiu_torch
trees = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000; p = int(sys.argv[3]) if len(sys.argv) > 3 else 100
d = make_cpiu_torch(n, p, levels=64)
ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
for rep in range(2):
    t0 = time.perf_counter()
    f = ds.grow(ntree=trees, nodesize=20, seed=-12345)
    wall = time.perf_counter() - t0
    print(f"rep {rep}: wall {wall:.2f} s | device {f['total_device_ms']/1e3:.2f} s = tree kernel {f['grow_kernel_ms']/1e3:.2f} + bootstrap {f['bootstrap_ms']/1e3:.2f} "
          f"+ membership {f['membership_kernel_ms']/1e3:.2f} + rest {(f['total_device_ms']-f['grow_kernel_ms']-f['bootstrap_ms']-f['membership_kernel_ms'])/1e3:.2f} | host side {wall - f['total_device_ms']/1e3:.2f} s | "
          f"nodes {f['treeID'].size/1e6:.1f} M, forest bytes {sum(f[k].nbytes for k in ('treeID','nodeID','parmID','contPT','mwcpSZ','tnRCNT','tnACNT','tnCLAS'))/1e9:.2f} GB", flush=True)
    del f
