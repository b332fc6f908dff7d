"""Host-side share of a resident configs[1] step: RFSLAM_TIMING=1 python tools/step_overhead.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import rfslam_b200
from rfslam_b200.synth import make_cpiu
d = make_cpiu(1_000_000, 50, levels=64)
ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
f = None
for rep in range(4):
    f = None
    t0 = time.perf_counter(); f = ds.grow(ntree=500, nodesize=20, seed=-12345); dt = time.perf_counter() - t0
    print(f"rep {rep}: python wall {1e3*dt:.1f} ms, device {f['total_device_ms']:.1f}, kernel {f['grow_kernel_ms']:.1f}, bootstrap {f['bootstrap_ms']:.1f}, membership {f['membership_kernel_ms']:.1f}", flush=True)
