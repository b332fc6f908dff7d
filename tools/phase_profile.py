"""Phase profile of the tree kernel: RFSLAM_PROFILE=1 python tools/phase_profile.py <rows> <covariates> <levels> <ntree> <nodesize> [torch] [reps]
("torch": synthesise the table on the GPU -- for 10M x 100, where the numpy generator takes minutes)"""
import os, time, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import rfslam_b200
from rfslam_b200.synth import make_cpiu, make_cpiu_torch
if len(sys.argv) > 6 and sys.argv[6] == "torch": make_cpiu = make_cpiu_torch
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
n, p, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ntree = int(sys.argv[4]); nodesize = int(sys.argv[5])
t=time.time(); d = make_cpiu(n, p, levels=L); print("gen", time.time()-t, "events", (d.y==2).mean(), "D", [int(np.unique(c[:100000]).size) for c in d.x][:12], flush=True)
t=time.time(); ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat); print("resident", time.time()-t, ds.hbm_bytes/1e6, "MB", flush=True)
for rep in range(reps):
    t=time.time(); f = ds.grow(ntree=ntree, nodesize=nodesize, seed=-12345); dt=time.time()-t
    print(f"grow wall {dt:.3f}s kernel {f['grow_kernel_ms']:.1f}ms total_dev {f['total_device_ms']:.1f}ms trees/s(kernel) {ntree/(f['grow_kernel_ms']/1e3):.1f} nodes {f['treeID'].size} cand {f['split_candidates']} cand/s {f['split_candidates']/(f['grow_kernel_ms']/1e3):.3e} algoGB {f['algorithmic_bytes']/1e9:.2f} GB/s {f['algorithmic_bytes']/1e9/(f['grow_kernel_ms']/1e3):.1f}", flush=True)
import ctypes as C
from rfslam_b200 import capi
buf = (C.c_uint64*96)()
k = capi.lib().rfslam_b200_last_profile(buf, 96)
if k:
    names = ["root","gate+draw","segpos+zero","hist","summary+wait","sm:presence","track","sm:rows|sort:walk","part_count","part_scatter","leaf","split_book","x","x","sort:radix","sm:cuts"]
    tot = sum(buf[i] for i in list(range(12))+[14,15])
    print("phase cycles (sum over trees), total %.3e" % tot)
    for i,nm in enumerate(names): print(f"  {nm:14s} {buf[i]:.3e} {100*buf[i]/max(tot,1):5.1f}%")
    print("node time by log2(len):")
    for c in range(0,28): 
        if buf[16+c]: print(f"  2^{c}: {buf[16+c]:.3e} {100*buf[16+c]/max(tot,1):5.1f}%")
    print("node counts by class pair:", [int(buf[48+i]) for i in range(14)])

if k >= 80:
    tots = sum(buf[64+i] for i in list(range(12))+[14,15])
    print("phases for nodes with < 64 entries: total %.3e (%.1f%% of all)" % (tots, 100*tots/max(tot,1)))
    for i,nm in enumerate(names):
        if buf[64+i]: print(f"  {nm:14s} {buf[64+i]:.3e} {100*buf[64+i]/max(tots,1):5.1f}%")

if k >= 96:
    tots = sum(buf[80+i] for i in list(range(12))+[14,15])
    print("phases for nodes with 64..512 entries: total %.3e (%.1f%% of all)" % (tots, 100*tots/max(tot,1)))
    for i,nm in enumerate(names):
        if buf[80+i]: print(f"  {nm:14s} {buf[80+i]:.3e} {100*buf[80+i]/max(tots,1):5.1f}%")
    totl = tot - tots - sum(buf[64+i] for i in list(range(12))+[14,15])
    print("phases for nodes with > 512 entries: total %.3e (%.1f%% of all)" % (totl, 100*totl/max(tot,1)))
    for i,nm in enumerate(names):
        v = buf[i] - buf[64+i] - buf[80+i]
        if v: print(f"  {nm:14s} {v:.3e} {100*v/max(totl,1):5.1f}%")
