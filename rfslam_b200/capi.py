"""ctypes binding of include/rfslam_b200.h (the C-ABI drop-in boundary).

The product path is the CUDA library: loading fails loudly when librfslam_b200.so is missing, and
every compute entry returns RFSLAM_ENODEV without a GPU -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RFSLAM_LIB", os.path.join(HERE, "librfslam_b200.so"))

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pu = C.POINTER(C.c_uint32)


class GrowArgs(C.Structure):
    _fields_ = [
        ("trace_flag", C.c_int32), ("seed", C.c_int32), ("opt_low", C.c_uint32), ("opt_high", C.c_uint32),
        ("split_rule", C.c_int32), ("nsplit", C.c_int32), ("mtry", C.c_int32), ("htry", C.c_int32),
        ("ytry", C.c_int32), ("node_size", C.c_int32), ("node_depth", C.c_int32),
        ("cr_weight_size", C.c_int32), ("cr_weight", _pd),
        ("ntree", C.c_int32), ("observation_size", C.c_int32), ("y_size", C.c_int32),
        ("r_type", C.c_char_p), ("r_levels", _pi), ("r_data", _pd),
        ("x_size", C.c_int32), ("x_type", C.c_char_p), ("x_levels", _pi),
        ("risk_time", _pd), ("stratifier", _pi), ("max_strat", C.c_int32),
        ("var_categories", _pi), ("category_weights", _pd), ("n_categories", C.c_int32),
        ("to_fix", _pi), ("n_to_fix", C.c_int32),
        ("k_for_alpha", C.c_double), ("max_bootstrap_size", C.c_int32), ("bootstrap_size", _pi), ("bootstrap", _pi),
        ("case_weight", _pd), ("x_split_stat_weight", _pd), ("y_weight", _pd), ("x_weight", _pd), ("x_data", _pd),
        ("time_interest_size", C.c_int32), ("time_interest", _pd),
        ("n_impute", C.c_int32), ("perf_block", C.c_int32), ("num_threads", C.c_int32),
        ("device", C.c_int32), ("tree_first", C.c_int32), ("tree_step", C.c_int32), ("finalize_ensembles", C.c_int32),
        ("want_split_stat", C.c_int32), ("want_member_lists", C.c_int32),
        ("comm", C.c_void_p), ("gather_forest", C.c_int32),
    ]


class ForestOut(C.Structure):
    _fields_ = [
        ("ntree", C.c_int32), ("n", C.c_int32), ("total_nodes", C.c_int64), ("total_leaves", C.c_int64),
        ("tree_ids", _pi), ("treeID", _pi), ("nodeID", _pi), ("parmID", _pi), ("contPT", _pd), ("mwcpSZ", _pi),
        ("splitStat", _pd), ("leafCount", _pi), ("seed", _pi), ("tnRCNT", _pi), ("tnACNT", _pi), ("tnCLAS", _pi),
        ("nodeMembership", _pi), ("bootMembership", _pi), ("rmbrMembership", _pi), ("ambrMembership", _pi),
        ("oobEnsbCLS", _pd), ("allEnsbCLS", _pd), ("oobEnsbDen", _pu), ("allEnsbDen", _pu),
        ("split_candidates", C.c_int64), ("algorithmic_bytes", C.c_int64),
        ("grow_kernel_ms", C.c_double), ("total_device_ms", C.c_double), ("kernel_launches", C.c_int64), ("rmbr_stride", C.c_int64), ("perfClas", _pd),
        ("mwcpPT", _pi), ("mwcpCT", _pi), ("mwcp_total", C.c_int64),
        ("grow_kernel_algorithmic_bytes", C.c_int64), ("membership_algorithmic_bytes", C.c_int64), ("bootstrap_algorithmic_bytes", C.c_int64),
        ("membership_kernel_ms", C.c_double), ("bootstrap_ms", C.c_double), ("collective_ms", C.c_double),
        ("vimpClas", _pd), ("vimp_count", C.c_int32), ("vimp_ms", C.c_double),
    ]


class PredictArgs(C.Structure):
    """Mirrors rfslam_predict_args: the 58 rfsrcPredict arguments in order, lists flattened, then the build knobs."""
    _fields_ = [
        ("trace_flag", C.c_int32), ("seed", C.c_int32), ("opt_low", C.c_uint32), ("opt_high", C.c_uint32),
        ("ntree", C.c_int32), ("observation_size", C.c_int32), ("y_size", C.c_int32),
        ("r_type", C.c_char_p), ("r_levels", _pi), ("r_data", _pd),
        ("x_size", C.c_int32), ("x_type", C.c_char_p), ("x_levels", _pi), ("x_data", _pd),
        ("k_for_alpha", C.c_double), ("max_bootstrap_size", C.c_int32), ("bootstrap_size", _pi), ("bootstrap", _pi),
        ("case_weight", _pd), ("time_interest_size", C.c_int32), ("time_interest", _pd),
        ("total_node_count", C.c_int32), ("forest_seed", _pi), ("htry", C.c_int32),
        ("treeID", _pi), ("nodeID", _pi), ("parmID", _pi), ("contPT", _pd), ("mwcpSZ", _pi), ("mwcpPT", _pi),
        ("tnRMBR", _pi), ("tnAMBR", _pi), ("tnRCNT", _pi), ("tnACNT", _pi),
        ("tnSURV", _pd), ("tnMORT", _pd), ("tnNLSN", _pd), ("tnCSHZ", _pd), ("tnCIFN", _pd), ("tnREGR", _pd),
        ("tnCLAS", _pi), ("r_target", _pi), ("r_target_count", C.c_int32), ("ptn_count", C.c_int32),
        ("intr_predictor_size", C.c_int32), ("intr_predictor", _pi),
        ("partial_type", C.c_int32), ("partial_xvar", C.c_int32), ("partial_length", C.c_int32),
        ("sobservation_size", C.c_int32), ("sobservation_indv", _pi),
        ("fobservation_size", C.c_int32), ("fr_size", C.c_int32), ("fr_data", _pd), ("fx_data", _pd),
        ("perf_block", C.c_int32), ("num_threads", C.c_int32),
        ("tnCLAS_len", C.c_int64), ("mwcp_total", C.c_int64), ("device", C.c_int32), ("finalize_ensembles", C.c_int32),
        ("frow_first", C.c_int32), ("frow_count", C.c_int32),
    ]


class PredictOut(C.Structure):
    _fields_ = [("m", C.c_int32), ("ntree", C.c_int32), ("allEnsbCLS", _pd), ("allEnsbDen", _pu),
                ("oobEnsbCLS", _pd), ("oobEnsbDen", _pu), ("perfClas", _pd), ("leafCount", _pi),
                ("nodeMembership", _pi), ("bootMembership", _pi), ("kernel_ms", C.c_double), ("total_device_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("row_tree_visits", C.c_int64), ("node_visits", C.c_int64),
                ("vimpClas", _pd), ("vimp_count", C.c_int32), ("vimp_ms", C.c_double)]


# every symbol include/rfslam_b200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "rfslam_b200_grow", "rfslam_b200_free_forest", "rfslam_b200_dataset_create", "rfslam_b200_grow_resident",
    "rfslam_b200_dataset_destroy", "rfslam_b200_dataset_bytes", "rfslam_b200_predict", "rfslam_b200_free_predict",
    "rfslam_b200_rule_of_slot", "rfslam_b200_register_this", "rfslam_b200_register_defaults",
    "rfslam_b200_score_node", "rfslam_b200_bootstrap_counts", "rfslam_b200_last_error", "rfslam_b200_version",
    "rfslam_b200_device_count", "rfslam_b200_last_profile", "rfslam_b200_free_array", "rfslam_b200_host_pool_trim",
    "rfslam_b200_comm_unique_id", "rfslam_b200_comm_init", "rfslam_b200_comm_destroy", "rfslam_b200_comm_rank",
    "rfslam_b200_comm_world", "rfslam_b200_comm_collectives", "rfslam_b200_nccl_version",
    "rfslam_b200_cpiu_build", "rfslam_b200_free_cpiu",
]
# the reference's own plugin symbols, exported with its signatures (src/splitCustom.h:24-44, src/splitCustom.c:23-67)
REFERENCE_SYMBOLS = ["registerThis", "registerCustomFunctions"] + [f"poissonSplit{k}" for k in range(1, 7)]
CUSTOM_FUNCTION = C.CFUNCTYPE(C.c_double, C.c_uint, C.c_double, C.c_char_p, _pd, _pd, C.c_uint, C.c_uint, _pd, _pd, C.c_double, C.c_double,
                              C.c_uint, C.POINTER(_pd), C.c_uint)

_lib = None


class RfslamError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rfslam_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m rfslam_b200.build` (nvcc, sm_100a). "
                "rfslam_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.rfslam_b200_grow.restype = C.c_int; L.rfslam_b200_grow.argtypes = [C.POINTER(GrowArgs), C.POINTER(ForestOut)]
        L.rfslam_b200_free_forest.restype = None; L.rfslam_b200_free_forest.argtypes = [C.POINTER(ForestOut)]
        L.rfslam_b200_dataset_create.restype = C.c_int; L.rfslam_b200_dataset_create.argtypes = [C.POINTER(GrowArgs), C.POINTER(vp)]
        L.rfslam_b200_grow_resident.restype = C.c_int; L.rfslam_b200_grow_resident.argtypes = [vp, C.POINTER(GrowArgs), C.POINTER(ForestOut)]
        L.rfslam_b200_dataset_destroy.restype = None; L.rfslam_b200_dataset_destroy.argtypes = [vp]
        L.rfslam_b200_dataset_bytes.restype = C.c_int64; L.rfslam_b200_dataset_bytes.argtypes = [vp]
        L.rfslam_b200_predict.restype = C.c_int; L.rfslam_b200_predict.argtypes = [C.POINTER(PredictArgs), C.POINTER(PredictOut)]
        L.rfslam_b200_free_predict.restype = None; L.rfslam_b200_free_predict.argtypes = [C.POINTER(PredictOut)]
        L.rfslam_b200_rule_of_slot.restype = C.c_int; L.rfslam_b200_rule_of_slot.argtypes = [C.c_int, C.c_int]
        L.rfslam_b200_register_this.restype = C.c_int; L.rfslam_b200_register_this.argtypes = [C.c_int, C.c_int, C.c_int]
        L.rfslam_b200_register_defaults.restype = None; L.rfslam_b200_register_defaults.argtypes = []
        L.rfslam_b200_score_node.restype = C.c_int
        L.rfslam_b200_score_node.argtypes = [C.c_int, C.c_double, C.c_int, _pd, _pd, _pi, _pd, _pd, _pd, _pi, C.c_int]
        L.rfslam_b200_bootstrap_counts.restype = C.c_int
        L.rfslam_b200_bootstrap_counts.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _pi, C.c_int]
        L.rfslam_b200_last_error.restype = C.c_char_p; L.rfslam_b200_last_error.argtypes = []
        L.rfslam_b200_version.restype = C.c_char_p; L.rfslam_b200_version.argtypes = []
        L.rfslam_b200_device_count.restype = C.c_int; L.rfslam_b200_device_count.argtypes = []
        L.rfslam_b200_last_profile.restype = C.c_int; L.rfslam_b200_last_profile.argtypes = [C.POINTER(C.c_uint64), C.c_int]
        L.rfslam_b200_free_array.restype = None; L.rfslam_b200_free_array.argtypes = [vp]
        L.rfslam_b200_host_pool_trim.restype = None; L.rfslam_b200_host_pool_trim.argtypes = []
        L.rfslam_b200_comm_unique_id.restype = C.c_int; L.rfslam_b200_comm_unique_id.argtypes = [vp, C.c_int]
        L.rfslam_b200_comm_init.restype = C.c_int; L.rfslam_b200_comm_init.argtypes = [C.POINTER(vp), C.c_int, C.c_int, vp, C.c_int]
        L.rfslam_b200_comm_destroy.restype = None; L.rfslam_b200_comm_destroy.argtypes = [vp]
        L.rfslam_b200_comm_rank.restype = C.c_int; L.rfslam_b200_comm_rank.argtypes = [vp]
        L.rfslam_b200_comm_world.restype = C.c_int; L.rfslam_b200_comm_world.argtypes = [vp]
        L.rfslam_b200_comm_collectives.restype = C.c_int64; L.rfslam_b200_comm_collectives.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.rfslam_b200_nccl_version.restype = C.c_int; L.rfslam_b200_nccl_version.argtypes = []
        L.registerThis.restype = None; L.registerThis.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
        L.registerCustomFunctions.restype = None; L.registerCustomFunctions.argtypes = []
        for k in range(1, 7):
            f = getattr(L, f"poissonSplit{k}")
            f.restype = C.c_double
            f.argtypes = [C.c_uint, C.c_double, C.c_void_p, _pd, _pd, C.c_uint, C.c_uint, _pd, C.c_void_p, C.c_double, C.c_double, C.c_uint, C.c_void_p, C.c_uint]
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RfslamError(rc, lib().rfslam_b200_last_error().decode(errors="replace"))
