"""rfslam_b200 -- B200-native (sm_100a) implementation of rfSLAM's Poisson-split tree-growing hot path.

Only what the path needs lives here: csrc/ (the CUDA kernels + the C-ABI of include/rfslam_b200.h),
forest.py (the host-side mirror of the reference's rfsrc()/predict interface, over the C-ABI),
synth.py (the synthetic CPIU generator of SURVEY.md 8d), shard.py (tree sharding across GPUs) and cpiu.py (the mirror
of the reference's CPIU construction, utilities/cpiu.R, over the device kernels of csrc/cpiu.cu).
"""
from .forest import ResidentCpiu, bootstrap_counts, predict_rfsrc, rfsrc, score_node  # noqa: F401
from .capi import RfslamError  # noqa: F401
from .cpiu import cpiu_fc  # noqa: F401
