"""Tree-kernel plan sweep: python tools/variant_sweep.py <rows> <covariates> <levels> <ntree> <nodesize> "<ENV=V ENV=V>" ["<...>" ...]
Grows the same forest under each environment (CTA width, table / histogram slots, subtree tile ...) and prints the
kernel time plus a checksum of the forest: the results must not depend on the plan."""
import os, sys, time, zlib, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import rfslam_b200
from rfslam_b200.synth import make_cpiu, make_cpiu_torch
n, p, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ntree = int(sys.argv[4]); nodesize = int(sys.argv[5])
gen = make_cpiu_torch if n >= 4_000_000 else make_cpiu
d = gen(n, p, levels=L)
ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
KEYS = ("RFSLAM_CTA_THREADS", "RFSLAM_TAB_SLOTS", "RFSLAM_HIST_SLOTS", "RFSLAM_NO_SUBTREE", "RFSLAM_SUBTREE_MAX", "RFSLAM_NO_SLICING", "RFSLAM_QUANTUM_MS", "RFSLAM_TIMING", "RFSLAM_CARVEOUT", "RFSLAM_FULL_KERNEL")
for cfg in sys.argv[6:] or [""]:
    for k in KEYS: os.environ.pop(k, None)
    for kv in cfg.split():
        k, v = kv.split("="); os.environ[k] = v
    best = 1e30
    for rep in range(2):
        f = ds.grow(ntree=ntree, nodesize=nodesize, seed=-12345)
        best = min(best, f['grow_kernel_ms'])
    crc = zlib.crc32(f['parmID'].tobytes()) ^ zlib.crc32(f['contPT'].tobytes()) ^ zlib.crc32(f['tnRCNT'].tobytes())
    print(f"[{cfg or 'default'}] kernel {best:.1f} ms  trees/s {ntree/best*1e3:.1f}  nodes {f['treeID'].size} crc {crc:08x}", flush=True)
