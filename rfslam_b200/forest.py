"""Host-side mirror of the reference's operator interface for the hot path.

``rfsrc()`` mirrors the argument names, defaults and option packing of the reference's R entry
point for the RF-SLAM configuration (/root/reference/R/rfsrc.R:1026-1067, :1360-1379, :1509-1652)
and returns the native output list under the reference's own names (randomForestSRC.c:9-79);
``predict_rfsrc()`` mirrors R/generic.predict.rfsrc.R:437-534.  Both go through the C-ABI
(include/rfslam_b200.h) -- the same calls a cgo/JNI/.Call binding would make.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import Dict, Optional, Sequence

import numpy as np

from . import capi
from .capi import ForestOut, GrowArgs, PredictArgs, PredictOut, _pd, _pi

SPLITRULES = {f"poisson.split{k}": k for k in range(1, 7)}


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_pd)


def _ip(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_pi)




def _take(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype) if ptr else None
    return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)


def _own(struct, field, count, dtype):
    """Adopt a callee-owned output array without copying: the struct's pointer is cleared (so
    rfslam_b200_free_* skips it) and rfslam_b200_free_array returns the block to the library's
    page-locked pool when the numpy array is collected."""
    ptr = getattr(struct, field)
    if not ptr:
        return None
    if count == 0:
        return np.zeros(0, dtype)
    addr = C.cast(ptr, C.c_void_p).value
    arr = np.ctypeslib.as_array(ptr, shape=(count,))
    assert arr.dtype == np.dtype(dtype), (field, arr.dtype, dtype)
    setattr(struct, field, type(ptr)())
    weakref.finalize(arr, capi.lib().rfslam_b200_free_array, addr)
    return arr


# perf.type -> optLow bits beside OPT_PERF (R/utilities.R:260-272): "brier" 2^3, "g.mean" 2^14
# importance -> optLow bits (R/utilities.R:150-208): OPT_VIMP 2^25, per-tree adds OPT_VIMP_LEOB 2^24, permutation OPT_VIMP_TYP1 2^8
IMPORTANCE_BITS = {None: 0, False: 0, "none": 0, True: (1 << 25) + (1 << 24) + (1 << 8), "permute": (1 << 25) + (1 << 24) + (1 << 8),
                   "permute.ensemble": (1 << 25) + (1 << 8)}
PERF_BITS = {"misclass": 0, "default": 0, "standard": 0, "brier": 1 << 3, "g.mean": 1 << 14, "gmean": 1 << 14}


class _Packed:
    """GrowArgs plus the numpy buffers it points into (kept alive for the call)."""

    def __init__(self, x, y, risk_time, stratifier, *, splitrule, ntree, mtry, nodesize, nodedepth, seed, k_for_alpha,
                 membership, samp, sampsize, split_wt, device, tree_first, tree_step, finalize, nsplit=0, xvar_types=None,
                 split_stat=False, member_lists=False, var_categories=None, category_weights=None, bootstrap=None, comm=None,
                 gather_forest=False, xvar_levels=None, perf_type="misclass", rfq=False, importance=None):
        x = np.ascontiguousarray(x, np.float64)
        assert x.ndim == 2, "x must be [p, n]: p contiguous columns of n (randomForestSRC.c:22190)"
        p, n = x.shape
        if isinstance(splitrule, str):
            if splitrule not in SPLITRULES:
                raise ValueError(f"splitrule {splitrule!r}: only {sorted(SPLITRULES)} are on the accelerated path")
            rule = SPLITRULES[splitrule]
        else:
            rule = int(splitrule)
        slot = rule + 1                                           # R/rfsrc.R:1360-1379 "poisson.splitN" -> "custom(N+1)"
        if mtry is None:
            mtry = int(math.ceil(math.sqrt(p)))                   # R/data.utilities.R:391-405
        if risk_time is None:
            risk_time = np.ones(n)                                # R/rfsrc.R:1509
        if stratifier is None:
            stratifier = np.ones(n, np.int32)                     # R/rfsrc.R:1510
        self.keep = dict(
            x=x, y=np.ascontiguousarray(y, np.float64), rt=np.ascontiguousarray(risk_time, np.float64),
            st=np.ascontiguousarray(stratifier, np.int32), rlev=np.array([2], np.int32),
            xlev=np.zeros(p, np.int32) if xvar_levels is None else np.ascontiguousarray(xvar_levels, np.int32),
            cat=np.ones(p, np.int32) if var_categories is None else np.ascontiguousarray(var_categories, np.int32),
            catw=np.array([1.0]) if category_weights is None else np.ascontiguousarray(category_weights, np.float64), tofix=np.zeros(1, np.int32), cw=np.ones(n), yw=np.array([1.0]),
            xw=np.ones(p), sw=np.ascontiguousarray(np.ones(p) if split_wt is None else split_wt, np.float64))
        k = self.keep
        # the C side reads n (per row) / p (per covariate) elements from each of these: check before it does
        if xvar_types is not None and len(xvar_types) != p:
            raise ValueError(f"xvar_types: expected {p} type characters")
        for name, want in (("y", n), ("rt", n), ("st", n), ("cat", p), ("sw", p), ("xlev", p)):
            if k[name].shape != (want,):
                raise ValueError(f"{name}: expected {want} elements for x of shape [p={p}, n={n}], got shape {k[name].shape}")
        if k["catw"].ndim != 1 or k["catw"].size < int(k["cat"].max()):
            raise ValueError("category_weights must hold one weight per variable category")
        a = GrowArgs()
        a.trace_flag = 0
        a.seed = int(seed)
        a.opt_low = (1 << 5) + (1 << 2)                           # forest + perf (R/utilities.R:89-103, :260-263)
        a.opt_low |= PERF_BITS[perf_type] | ((1 << 15) if rfq else 0)   # perf.type / rfq (R/utilities.R:254-290)
        if importance not in IMPORTANCE_BITS:
            raise ValueError(f"importance {importance!r}: only permutation importance (\"permute\", \"permute.ensemble\") is on the accelerated path")
        a.opt_low |= IMPORTANCE_BITS[importance]
        a.opt_high = 256 * (slot - 1) + (1 << 16)                 # split.cust bits + terminal.qualts (R/utilities.R:403-459)
        if membership:
            a.opt_high |= (1 << 6)                                # R/utilities.R:333-348
        if bootstrap not in (None, "by.root", "by.user"):
            raise ValueError(f"bootstrap {bootstrap!r}: only by.root and by.user (samp) are on the accelerated path")
        if samp is not None:
            a.opt_low |= (1 << 19) + (1 << 20)                    # bootstrap = "by.user" (R/utilities.R:43-45)
            k["samp"] = np.ascontiguousarray(samp, np.int32)
            assert k["samp"].shape == (ntree, n)
            a.bootstrap = _ip(k["samp"])
            k["bs"] = k["samp"].sum(axis=1).astype(np.int32)
        else:
            k["bs"] = np.full(ntree, n if sampsize is None else int(sampsize), np.int32)
        a.split_rule = 16
        a.nsplit = int(nsplit)
        a.mtry = int(mtry); a.htry = 0; a.ytry = 1
        a.node_size = int(nodesize); a.node_depth = int(nodedepth)
        a.cr_weight_size = 0
        a.ntree = int(ntree); a.observation_size = n; a.y_size = 1
        self._rtype = b"C"; a.r_type = self._rtype
        a.r_levels = _ip(k["rlev"]); a.r_data = _dp(k["y"])
        a.x_size = p
        self._xtype = ("".join(xvar_types) if xvar_types is not None else "R" * p).encode(); a.x_type = self._xtype
        a.x_levels = _ip(k["xlev"])
        a.risk_time = _dp(k["rt"]); a.stratifier = _ip(k["st"]); a.max_strat = int(k["st"].max())
        a.var_categories = _ip(k["cat"]); a.category_weights = _dp(k["catw"]); a.n_categories = int(np.unique(k["cat"]).size)
        a.to_fix = _ip(k["tofix"]); a.n_to_fix = 0
        a.k_for_alpha = float(k_for_alpha)
        a.max_bootstrap_size = int(k["bs"].max()); a.bootstrap_size = _ip(k["bs"])
        a.case_weight = _dp(k["cw"]); a.x_split_stat_weight = _dp(k["sw"]); a.y_weight = _dp(k["yw"]); a.x_weight = _dp(k["xw"])
        a.x_data = _dp(k["x"])
        a.time_interest_size = 0; a.n_impute = 1; a.perf_block = int(ntree); a.num_threads = -1
        a.device = int(device); a.tree_first = int(tree_first); a.tree_step = int(tree_step)
        a.finalize_ensembles = 1 if finalize else 0
        a.want_split_stat = 1 if split_stat else 0
        a.want_member_lists = 1 if member_lists else 0
        a.comm = comm.handle if comm is not None else None
        a.gather_forest = 1 if (gather_forest and comm is not None) else 0
        self.args = a
        self.n, self.p = n, p


def _unpack_forest(o: ForestOut, finalized: bool = True, ntree_total: int = 0) -> Dict[str, np.ndarray]:
    n, T, NN, NL = o.n, o.ntree, o.total_nodes, o.total_leaves
    res = {
        "tree_ids": _own(o, "tree_ids", T, np.int32),
        "treeID": _own(o, "treeID", NN, np.int32), "nodeID": _own(o, "nodeID", NN, np.int32),
        "parmID": _own(o, "parmID", NN, np.int32), "contPT": _own(o, "contPT", NN, np.float64),
        "mwcpSZ": _own(o, "mwcpSZ", NN, np.int32), "splitStat": _own(o, "splitStat", NN, np.float64),
        "mwcpPT": _own(o, "mwcpPT", int(o.mwcp_total), np.int32) if o.mwcpPT else np.zeros(0, np.int32),
        "mwcpCT": _own(o, "mwcpCT", T, np.int32),
        "leafCount": _own(o, "leafCount", T, np.int32), "seed": _own(o, "seed", T, np.int32),
        "tnRCNT": _own(o, "tnRCNT", NL, np.int32), "tnACNT": _own(o, "tnACNT", NL, np.int32),
        "tnCLAS": _own(o, "tnCLAS", 2 * NL, np.int32),
        "oobEnsbCLS": _own(o, "oobEnsbCLS", 2 * n, np.float64), "allEnsbCLS": _own(o, "allEnsbCLS", 2 * n, np.float64),
        "oobEnsbDen": _own(o, "oobEnsbDen", n, np.uint32), "allEnsbDen": _own(o, "allEnsbDen", n, np.uint32),
        "split_candidates": int(o.split_candidates), "algorithmic_bytes": int(o.algorithmic_bytes),
        "grow_kernel_ms": float(o.grow_kernel_ms), "total_device_ms": float(o.total_device_ms),
        "kernel_launches": int(o.kernel_launches),
        "grow_kernel_algorithmic_bytes": int(o.grow_kernel_algorithmic_bytes), "membership_algorithmic_bytes": int(o.membership_algorithmic_bytes),
        "bootstrap_algorithmic_bytes": int(o.bootstrap_algorithmic_bytes), "membership_kernel_ms": float(o.membership_kernel_ms),
        "bootstrap_ms": float(o.bootstrap_ms), "collective_ms": float(o.collective_ms),
        "finalized": bool(finalized),             # False: oob/allEnsbCLS are the shard's sums (see shard.allreduce_ensembles)
    }
    if o.rmbrMembership:
        res["rmbrMembership"] = _own(o, "rmbrMembership", T * int(o.rmbr_stride), np.int32)
        res["ambrMembership"] = _own(o, "ambrMembership", T * n, np.int32)
    if o.nodeMembership:
        res["nodeMembership"] = _own(o, "nodeMembership", T * n, np.int32)
        res["bootMembership"] = _own(o, "bootMembership", T * n, np.int32)
    if o.perfClas:
        res["perfClas"] = _own(o, "perfClas", (ntree_total or T) * 3, np.float64)
    if o.vimpClas:
        res["importance"] = _own(o, "vimpClas", int(o.vimp_count) * 3, np.float64).reshape(-1, 3)
        res["vimp_ms"] = float(o.vimp_ms)
    return res


class ResidentCpiu:
    """The CPIU table uploaded and rank-coded once, resident in HBM (rfslam_b200_dataset_create)."""

    def __init__(self, x, y, risk_time=None, stratifier=None, *, device: int = -1, xvar_types=None, xvar_levels=None):
        self._xvar = (xvar_types, xvar_levels)                    # 'C' = unordered factor with levels 1 .. xvar_levels[c]
        self._pk = _Packed(x, y, risk_time, stratifier, splitrule=4, ntree=1, mtry=1, nodesize=1, nodedepth=-1, seed=-1,
                           k_for_alpha=2.0, membership=False, samp=None, sampsize=None, split_wt=None, device=device,
                           tree_first=0, tree_step=0, finalize=True, xvar_types=xvar_types, xvar_levels=xvar_levels)
        self.handle = C.c_void_p()
        capi.check(capi.lib().rfslam_b200_dataset_create(C.byref(self._pk.args), C.byref(self.handle)))
        self.n, self.p = self._pk.n, self._pk.p
        self.device = device

    @property
    def hbm_bytes(self) -> int:
        return int(capi.lib().rfslam_b200_dataset_bytes(self.handle))

    def close(self):
        if self.handle:
            capi.lib().rfslam_b200_dataset_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def grow(self, *, splitrule="poisson.split4", ntree=500, mtry=None, nodesize=5, nodedepth=-1, seed=-12345,
             k_for_alpha=2.0, membership=False, samp=None, sampsize=None, split_wt=None, tree_first=0, tree_step=0,
             finalize=True, split_stat=False, member_lists=False, nsplit=0, bootstrap=None, comm=None,
             gather_forest=False, perf_type="misclass", rfq=False, importance=None) -> Dict[str, np.ndarray]:
        k = self._pk.keep
        pk = _Packed(k["x"], k["y"], k["rt"], k["st"], splitrule=splitrule, ntree=ntree, mtry=mtry, nodesize=nodesize,
                     nodedepth=nodedepth, seed=seed, k_for_alpha=k_for_alpha, membership=membership, samp=samp,
                     sampsize=sampsize, split_wt=split_wt, device=self.device, tree_first=tree_first, tree_step=tree_step,
                     finalize=finalize, split_stat=split_stat, member_lists=member_lists, nsplit=nsplit, bootstrap=bootstrap,
                     comm=comm, gather_forest=gather_forest, xvar_types=self._xvar[0], xvar_levels=self._xvar[1],
                     perf_type=perf_type, rfq=rfq, importance=importance)
        out = ForestOut()
        capi.check(capi.lib().rfslam_b200_grow_resident(self.handle, C.byref(pk.args), C.byref(out)))
        res = _unpack_forest(out, finalize, int(ntree))
        capi.lib().rfslam_b200_free_forest(C.byref(out))
        return res


def rfsrc(x, y, risk_time=None, stratifier=None, *, splitrule="poisson.split4", ntree=500, mtry=None, nodesize=5,
          nodedepth=-1, seed=-12345, k_for_alpha=2.0, membership=False, samp=None, sampsize=None, split_wt=None,
          device=-1, tree_first=0, tree_step=0, finalize=True, nsplit=0, xvar_types=None, split_stat=False,
          member_lists=False, var_categories=None, category_weights=None, bootstrap=None, comm=None,
          gather_forest=False, xvar_levels=None, perf_type="misclass", rfq=False, importance=None) -> Dict[str, np.ndarray]:
    """Grow an RF-SLAM forest from HOST arrays through rfslam_b200_grow (upload + grow + download).

    x [p, n] float64 (p contiguous columns), y response codes 1.0 / 2.0, risk_time [n], stratifier [n]
    (interval number >= 1); defaults follow R/rfsrc.R:1026-1067 (k_for_alpha = 2, nodedepth unlimited,
    mtry = ceil(sqrt(p)), bootstrap by.root unless ``samp`` is given (by.user)).  ``importance`` = "permute" (or True:
    per-tree permutation importance) / "permute.ensemble" adds ``importance`` [p][3] = vimpClas (R/rfsrc.R:2389-2392)."""
    pk = _Packed(x, y, risk_time, stratifier, splitrule=splitrule, ntree=ntree, mtry=mtry, nodesize=nodesize,
                 nodedepth=nodedepth, seed=seed, k_for_alpha=k_for_alpha, membership=membership, samp=samp, sampsize=sampsize,
                 split_wt=split_wt, device=device, tree_first=tree_first, tree_step=tree_step, finalize=finalize, nsplit=nsplit,
                 xvar_types=xvar_types, split_stat=split_stat, member_lists=member_lists, var_categories=var_categories,
                 category_weights=category_weights, bootstrap=bootstrap, comm=comm, gather_forest=gather_forest,
                 xvar_levels=xvar_levels, perf_type=perf_type, rfq=rfq, importance=importance)
    out = ForestOut()
    capi.check(capi.lib().rfslam_b200_grow(C.byref(pk.args), C.byref(out)))
    res = _unpack_forest(out, finalize, int(ntree))
    capi.lib().rfslam_b200_free_forest(C.byref(out))
    return res


def predict_rfsrc(forest: Dict[str, np.ndarray], fx=None, *, x=None, y=None, membership=True, device=-1, finalize=True,
                  samp=None, sampsize=None, use_tnclas=True, frow_first=0, frow_count=0, perf=True,
                  xvar_types=None, xvar_levels=None, perf_type="misclass", rfq=False, importance=None,
                  xvar_importance=None, fy=None) -> Dict[str, np.ndarray]:
    """rfsrcPredict through the C-ABI (rfslam_b200_predict), mirroring R/generic.predict.rfsrc.R:437-534.

    ``forest`` is a grow result (the reference's wire format).  ``fx`` [p, m] predicts held-out rows; ``fx=None``
    RESTORES the forest on its training rows ``x`` / ``y`` (OOB + full ensembles, perfClas).  The leaf class counts
    come from ``forest["tnCLAS"]`` when present (and ``use_tnclas``), else they are rebuilt from the training rows
    with the bootstrap replayed from ``forest["seed"]`` (or ``samp``).  ``frow_first`` / ``frow_count`` select this
    process' shard of the held-out rows (forest replicated, no collective).  In restore mode ``importance`` ("permute" /
    "permute.ensemble") adds the permutation importance [len(xvar_importance)][3] of the 1-based covariates in
    ``xvar_importance`` (default: all), as vimp.rfsrc obtains it (R/generic.predict.rfsrc.R:1018-1021).  ``fy`` [m] = the
    held-out rows' responses (newdata that carries the outcome): ``perfClas`` then scores the full ensemble on them."""
    keep = {k: np.ascontiguousarray(forest[k], np.int32) for k in ("treeID", "nodeID", "parmID", "seed")}
    keep["contPT"] = np.ascontiguousarray(forest["contPT"], np.float64)
    keep["mwcpSZ"] = np.ascontiguousarray(forest["mwcpSZ"], np.int32) if "mwcpSZ" in forest else np.zeros(keep["treeID"].size, np.int32)
    ntree = int(keep["seed"].size)
    a = PredictArgs()
    a.trace_flag = 0; a.seed = -1
    a.opt_low = ((1 << 2) if perf else 0) + (1 << 5) + PERF_BITS[perf_type] + ((1 << 15) if rfq else 0)
    a.opt_high = (1 << 17)
    if importance not in IMPORTANCE_BITS:
        raise ValueError(f"importance {importance!r}: only permutation importance is on the accelerated path")
    a.opt_low |= IMPORTANCE_BITS[importance]
    if membership:
        a.opt_high |= (1 << 6)
    a.ntree = ntree; a.y_size = 1
    rtype = b"C"; a.r_type = rtype
    keep["rlev"] = np.array([2], np.int32); a.r_levels = _ip(keep["rlev"])
    if fx is not None:
        fx = np.ascontiguousarray(fx, np.float64)
        p, m = fx.shape
    else:
        if x is None or y is None:
            raise ValueError("restore mode (fx=None) needs the training x and y")
        p, m = np.asarray(x).shape[0], 0
    if x is not None:
        keep["x"] = np.ascontiguousarray(x, np.float64); keep["y"] = np.ascontiguousarray(y, np.float64)
        if keep["x"].shape[0] != p or keep["y"].shape != (keep["x"].shape[1],):
            raise ValueError("training x must be [p, n] and y [n]")
        n = keep["x"].shape[1]
        a.x_data = _dp(keep["x"]); a.r_data = _dp(keep["y"])
    else:
        n = int(forest["allEnsbDen"].size) if "allEnsbDen" in forest else 1
    a.observation_size = n; a.x_size = p
    if IMPORTANCE_BITS[importance]:
        keep["intr"] = np.ascontiguousarray(np.arange(1, p + 1) if xvar_importance is None else xvar_importance, np.int32)
        a.intr_predictor_size = int(keep["intr"].size); a.intr_predictor = _ip(keep["intr"])
    xtype = ("".join(xvar_types) if xvar_types is not None else "R" * p).encode(); a.x_type = xtype
    keep["xlev"] = np.zeros(p, np.int32) if xvar_levels is None else np.ascontiguousarray(xvar_levels, np.int32); a.x_levels = _ip(keep["xlev"])
    a.k_for_alpha = 2.0
    if samp is not None:
        a.opt_low |= (1 << 19) + (1 << 20)
        keep["samp"] = np.ascontiguousarray(samp, np.int32)
        assert keep["samp"].shape == (ntree, n)
        a.bootstrap = _ip(keep["samp"]); keep["bs"] = keep["samp"].sum(axis=1).astype(np.int32)
    else:
        keep["bs"] = np.full(ntree, n if sampsize is None else int(sampsize), np.int32)
    a.max_bootstrap_size = int(keep["bs"].max()); a.bootstrap_size = _ip(keep["bs"])
    a.total_node_count = int(keep["treeID"].size); a.forest_seed = _ip(keep["seed"]); a.htry = 0
    a.treeID = _ip(keep["treeID"]); a.nodeID = _ip(keep["nodeID"]); a.parmID = _ip(keep["parmID"]); a.contPT = _dp(keep["contPT"])
    a.mwcpSZ = _ip(keep["mwcpSZ"])
    if "mwcpPT" in forest and np.asarray(forest["mwcpPT"]).size:  # hc_zero[[4]]: the factor splits' MWCP words in record order
        keep["mwcpPT"] = np.ascontiguousarray(forest["mwcpPT"], np.int32)
        a.mwcpPT = _ip(keep["mwcpPT"]); a.mwcp_total = int(keep["mwcpPT"].size)
    for k, f in (("tnRCNT", "tnRCNT"), ("tnACNT", "tnACNT"), ("rmbrMembership", "tnRMBR"), ("ambrMembership", "tnAMBR")):
        if k in forest and forest[k] is not None:
            keep[f] = np.ascontiguousarray(forest[k], np.int32); setattr(a, f, _ip(keep[f]))
    if use_tnclas and forest.get("tnCLAS") is not None:
        keep["tnCLAS"] = np.ascontiguousarray(forest["tnCLAS"], np.int32)
        a.tnCLAS = _ip(keep["tnCLAS"]); a.tnCLAS_len = int(keep["tnCLAS"].size); a.opt_high |= (1 << 19)
    keep["rtarget"] = np.array([1], np.int32); a.r_target = _ip(keep["rtarget"]); a.r_target_count = 1
    a.fobservation_size = m
    if fx is not None:
        a.fx_data = _dp(fx)
        if fy is not None:
            keep["fy"] = np.ascontiguousarray(fy, np.float64)
            if keep["fy"].shape != (m,):
                raise ValueError(f"fy: expected {m} held-out responses")
            a.fr_size = 1; a.fr_data = _dp(keep["fy"])
    a.perf_block = ntree; a.num_threads = -1
    a.device = int(device); a.finalize_ensembles = 1 if finalize else 0
    a.frow_first = int(frow_first); a.frow_count = int(frow_count)
    out = PredictOut()
    capi.check(capi.lib().rfslam_b200_predict(C.byref(a), C.byref(out)))
    M = int(out.m)
    res = {"allEnsbCLS": _take(out.allEnsbCLS, 2 * M, np.float64), "allEnsbDen": _take(out.allEnsbDen, M, np.uint32),
           "leafCount": _take(out.leafCount, ntree, np.int32), "kernel_ms": float(out.kernel_ms),
           "total_device_ms": float(out.total_device_ms), "kernel_launches": int(out.kernel_launches),
           "row_tree_visits": int(out.row_tree_visits), "node_visits": int(out.node_visits), "finalized": bool(finalize)}
    if out.oobEnsbCLS:
        res["oobEnsbCLS"] = _take(out.oobEnsbCLS, 2 * M, np.float64); res["oobEnsbDen"] = _take(out.oobEnsbDen, M, np.uint32)
    if out.perfClas:
        res["perfClas"] = _take(out.perfClas, 3 * ntree, np.float64)
    if out.vimpClas:
        res["importance"] = _take(out.vimpClas, 3 * int(out.vimp_count), np.float64).reshape(-1, 3)
        res["vimp_ms"] = float(out.vimp_ms)
    if out.nodeMembership:
        res["nodeMembership"] = _take(out.nodeMembership, ntree * M, np.int32)
    if out.bootMembership:
        res["bootMembership"] = _take(out.bootMembership, ntree * n, np.int32)
    capi.lib().rfslam_b200_free_predict(C.byref(out))
    return res


def score_node(rule: int, k_for_alpha: float, x, response, stratum, risk_time, *, device=-1):
    """Delta for every cutpoint of one covariate in one node (rfslam_b200_score_node): what the
    reference obtains from one customFunction call per cut (randomForestSRC.c:13404-13529)."""
    x = np.ascontiguousarray(x, np.float64); r = np.ascontiguousarray(response, np.float64)
    s = np.ascontiguousarray(stratum, np.int32); t = np.ascontiguousarray(risk_time, np.float64)
    m = x.size
    cut = np.zeros(m, np.float64); delta = np.zeros(m, np.float64); nc = C.c_int32(0)
    capi.check(capi.lib().rfslam_b200_score_node(int(rule), float(k_for_alpha), m, _dp(x), _dp(r), _ip(s), _dp(t),
                                                 _dp(cut), _dp(delta), C.byref(nc), device))
    return cut[: nc.value].copy(), delta[: nc.value].copy()


def bootstrap_counts(seed: int, ntree: int, n: int, tree: int, *, device=-1) -> np.ndarray:
    out = np.zeros(n, np.int32)
    capi.check(capi.lib().rfslam_b200_bootstrap_counts(seed, ntree, n, tree, _ip(out), device))
    return out
