import sys, time, numpy as np
sys.path.insert(0, '.')
import rfslam_b200
from rfslam_b200.synth import make_cpiu
d = make_cpiu(1000000, 50, levels=64)
for rep in range(3):
    t=time.time(); f = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=500, nodesize=20, seed=-12345); print("rfsrc wall", time.time()-t, f['grow_kernel_ms'], f['total_device_ms'], flush=True)
    del f
