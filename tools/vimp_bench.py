"""Permutation importance at configs[1]'s table: device time of the importance pass next to the grow call it rides on.
usage: python tools/vimp_bench.py [ntree] [n] [p] [mode]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import rfslam_b200
from rfslam_b200.synth import make_cpiu

ntree = int(sys.argv[1]) if len(sys.argv) > 1 else 148
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
p = int(sys.argv[3]) if len(sys.argv) > 3 else 50
mode = sys.argv[4] if len(sys.argv) > 4 else "permute"
d = make_cpiu(n, p, levels=64, seed=1)
kw = dict(splitrule=4, ntree=ntree, mtry=8, nodesize=20, seed=-7)
rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, **kw)                     # warm-up (pools, module load)
t0 = time.time(); plain = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, **kw); t1 = time.time()
imp = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, importance=mode, **kw); t2 = time.time()
print(json.dumps({"workload": f"{n} x {p}, {ntree} trees, importance = {mode}", "grow_call_s": round(t1 - t0, 3),
                  "grow_with_importance_call_s": round(t2 - t1, 3), "vimp_device_ms": round(imp["vimp_ms"], 1),
                  "permutations": ntree * p, "oob_rows_per_tree": round(0.3679 * n),
                  "row_walks_per_s": round(ntree * (p + 1) * 0.3679 * n / (imp["vimp_ms"] / 1e3)),
                  "top5": [int(i) for i in (-imp["importance"][:, 0]).argsort()[:5]]}))
