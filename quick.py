import sys; sys.path.insert(0,'.')
import numpy as np, rfslam_b200
from rfslam_b200.synth import make_cpiu
from oracle import port
d = make_cpiu(4000, 9, levels=64, seed=3)
got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=4, nodesize=5, seed=-11, membership=True)
want = port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=4, nodesize=5, seed=-11)
print("parm equal", np.array_equal(got["parmID"], want["parmID"]), got["treeID"].size)
