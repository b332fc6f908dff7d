"""ctypes driver for the UNMODIFIED reference compiled into oracle/_ref/librfslam_ref.so.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
``--impl reference`` legs may import this module; the product (rfslam_b200/) never does.

It packs the 44 ``.Call("rfsrcGrow", ...)`` arguments exactly in the order of
/root/reference/src/randomForestSRC.c:22110-22155 (the R side is R/rfsrc.R:1590-1650) and the 58
``rfsrcPredict`` arguments in the order of randomForestSRC.c:22219-22279
(R/generic.predict.rfsrc.R:437-532), through the fake-R boxes of oracle/fake_r/, and returns
the named output list as a dict of numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "librfslam_ref.so")
# the PRODUCT's rfsrcGrow / rfsrcPredict (rfslam_b200/csrc/r_shim.c over librfslam_b200.so) compiled against the same
# fake-R boxes (`make -C oracle rshim`): the same packer below drives it when impl="b200"
SHIM_SO = os.path.join(_HERE, "_ref", "librfslam_b200_rshim.so")

_libs = {}


def available(impl: str = "ref") -> bool:
    return os.path.exists(REF_SO if impl == "ref" else SHIM_SO)


def lib(impl: str = "ref"):
    if impl not in _libs:
        if not available(impl):
            raise RuntimeError(
                "oracle/_ref/librfslam_ref.so is missing: run `make -C oracle ref` in a container "
                "that has /root/reference (the GPU box only uses the prebuilt file)" if impl == "ref" else
                "oracle/_ref/librfslam_b200_rshim.so is missing: run `make -C oracle rshim`")
        L = C.CDLL(REF_SO if impl == "ref" else SHIM_SO, mode=C.RTLD_LOCAL)
        vp = C.c_void_p
        L.fr_int.restype = vp; L.fr_int.argtypes = [vp, C.c_long]
        L.fr_real.restype = vp; L.fr_real.argtypes = [vp, C.c_long]
        L.fr_real_nocopy.restype = vp; L.fr_real_nocopy.argtypes = [vp, C.c_long]
        L.fr_int_nocopy.restype = vp; L.fr_int_nocopy.argtypes = [vp, C.c_long]
        L.fr_str.restype = vp; L.fr_str.argtypes = [C.POINTER(C.c_char_p), C.c_long]
        L.fr_list.restype = vp; L.fr_list.argtypes = [C.c_long]
        L.fr_list_set.restype = None; L.fr_list_set.argtypes = [vp, C.c_long, vp]
        L.fr_nil.restype = vp; L.fr_nil.argtypes = []
        L.fr_type.restype = C.c_int; L.fr_type.argtypes = [vp]
        L.fr_len.restype = C.c_long; L.fr_len.argtypes = [vp]
        L.fr_data.restype = vp; L.fr_data.argtypes = [vp]
        L.fr_name.restype = C.c_char_p; L.fr_name.argtypes = [vp, C.c_long]
        L.fr_elt.restype = vp; L.fr_elt.argtypes = [vp, C.c_long]
        L.fr_release_all.restype = None; L.fr_release_all.argtypes = []
        L.fr_guard_call.restype = vp
        L.fr_guard_call.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_char_p, C.c_int]
        _libs[impl] = L
    return _libs[impl]


class _Packer:
    """Keeps the numpy buffers alive for the duration of one call."""

    def __init__(self, impl: str = "ref"):
        self.L = lib(impl)
        self.keep = []

    def i(self, v) -> int:
        a = np.ascontiguousarray(np.atleast_1d(v), dtype=np.int32)
        self.keep.append(a)
        return self.L.fr_int(a.ctypes.data, a.size)

    def d(self, v) -> int:
        a = np.ascontiguousarray(np.atleast_1d(v), dtype=np.float64)
        self.keep.append(a)
        return self.L.fr_real(a.ctypes.data, a.size)

    def d_alias(self, a: np.ndarray) -> int:
        assert a.dtype == np.float64 and a.flags.c_contiguous
        self.keep.append(a)
        return self.L.fr_real_nocopy(a.ctypes.data, a.size)

    def i_alias(self, a: np.ndarray) -> int:
        assert a.dtype == np.int32 and a.flags.c_contiguous
        self.keep.append(a)
        return self.L.fr_int_nocopy(a.ctypes.data, a.size)

    def s(self, v: Sequence[str]) -> int:
        arr = (C.c_char_p * max(len(v), 1))(*[x.encode() for x in v])
        self.keep.append(arr)
        return self.L.fr_str(arr, len(v))

    def lst(self, items) -> int:
        l = self.L.fr_list(len(items))
        for k, it in enumerate(items):
            self.L.fr_list_set(l, k, it)
        return l

    def nil(self) -> int:
        return self.L.fr_nil()


_INTSXP, _REALSXP = 13, 14


def _unpack(L, out) -> Dict[str, np.ndarray]:
    res: Dict[str, np.ndarray] = {}
    n = L.fr_len(out)
    for k in range(n):
        name = L.fr_name(out, k).decode()
        e = L.fr_elt(out, k)
        t, ln, p = L.fr_type(e), L.fr_len(e), L.fr_data(e)
        if t == _INTSXP:
            arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(ln,)).copy() if ln else np.zeros(0, np.int32)
        elif t == _REALSXP:
            arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(ln,)).copy() if ln else np.zeros(0, np.float64)
        else:
            continue
        res[name] = arr
    return res


# R/rfsrc.R:1360-1379 maps "poisson.splitN" -> "custom(N+1)"; data.utilities.R:497-501 -> cust.idx;
# utilities.R:403-418 -> optHigh bits 256*(idx-1).
def split_slot(rule: int) -> int:
    assert 1 <= rule <= 6
    return rule + 1


# importance (R/utilities.R:150-208): OPT_VIMP 2^25, per-tree (Breiman-Cutler) adds OPT_VIMP_LEOB 2^24, "permute" TYP1 2^8,
# "random" TYP2 2^9, "anti" neither
def importance_bits(importance: Optional[str]) -> int:
    if importance in (None, "none"):
        return 0
    kind = {"anti": 0, "permute": 1 << 8, "random": 1 << 9}[importance.split(".")[0]]
    return (1 << 25) + kind + (0 if importance.endswith(".ensemble") else (1 << 24))


def grow(x: np.ndarray, y: np.ndarray, risk_time: np.ndarray, stratifier: np.ndarray, *,
         rule: int = 4, ntree: int = 4, mtry: Optional[int] = None, nodesize: int = 5,
         nodedepth: int = -1, seed: int = -12345, k_for_alpha: float = 2.0,
         membership: bool = True, nthreads: int = -1, samp: Optional[np.ndarray] = None,
         nsplit: int = 0, split_weight: Optional[np.ndarray] = None,
         xtype: Optional[Sequence[str]] = None, xlevels: Optional[np.ndarray] = None,
         var_categories: Optional[np.ndarray] = None, category_weights: Optional[np.ndarray] = None,
         statistics: bool = False, sampsize: Optional[int] = None, impl: str = "ref", quants: bool = False,
         perf: str = "misclass", rfq: bool = False, importance: Optional[str] = None) -> Dict[str, np.ndarray]:
    """Run the reference's rfsrcGrow.  ``x`` is [p, n] float64 (p contiguous columns of n), ``y``
    the response codes in {1.0, 2.0}.  ``samp`` ([ntree, n] int32 in-bag counts) selects
    bootstrap="by.user" (R/utilities.R:43-45 -> bits 2^19+2^20).  ``statistics`` sets OPT_NODE_STAT
    (randomForestSRC.h:133, R statistics = TRUE): per node the covariates tried and the best statistic of
    each come back as ``mtryID`` / ``mtryST`` [node][mtry] (randomForestSRC.c:15553-15575; ``spltST`` itself is
    never filled on the custom-rule path of this fork).  They are appended under their own critical section
    (randomForestSRC.c:20948-20962), so they line up with ``treeID`` only when ``nthreads == 1``;
    trim_forest() turns them into ``splitStat`` = the statistic of the covariate the node split on."""
    L = lib(impl)
    p, n = x.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    if mtry is None:
        mtry = int(np.ceil(np.sqrt(p)))        # R/data.utilities.R:391-405
    P = _Packer(impl)
    opt_low = (1 << 5) + (1 << 2)              # forest + perf
    opt_high = 256 * (split_slot(rule) - 1) + (1 << 16)
    if quants:
        opt_high += (1 << 18)                  # terminal.quants outgoing: tnCLAS (R/utilities.R:462-483)
    if membership:
        opt_high += (1 << 6)
    # perf.type (R/utilities.R:254-272): "brier" -> OPT_PERF_CALB 2^3, "g.mean" -> OPT_PERF_GMN2 2^14; rfq -> OPT_CLAS_RFQ 2^15
    opt_low += {"misclass": 0, "brier": 1 << 3, "gmean": 1 << 14}[perf]
    if rfq:
        opt_low += (1 << 15)
    opt_low += importance_bits(importance)
    if statistics:
        assert nthreads == 1, "spltST is only in tree order with one thread"
        opt_low += 0x08000000
    if samp is not None:
        opt_low += (1 << 19) + (1 << 20)
        samp = np.ascontiguousarray(samp, dtype=np.int32)
        assert samp.shape == (ntree, n)
        boot_sizes = samp.sum(axis=1).astype(np.int32)
        max_boot = int(boot_sizes.max())
        samp_arg = P.i_alias(samp.reshape(-1))
    else:
        boot_sizes = np.full(ntree, n if sampsize is None else int(sampsize), np.int32)
        max_boot = int(boot_sizes.max())
        samp_arg = P.i(np.zeros(0, np.int32))
    if xtype is None:
        xtype = ["R"] * p
    if xlevels is None:
        xlevels = np.zeros(p, np.int32)
    if split_weight is None:
        split_weight = np.ones(p)
    if var_categories is None:
        var_categories = np.ones(p, np.int32); category_weights = np.array([1.0])
    ncat = int(np.unique(var_categories).size)
    args = [
        P.i(0), P.i(seed), P.i(opt_low), P.i(opt_high),
        P.i(16),                    # splitRule = custom (randomForestSRC.h:189)
        P.i(nsplit), P.i(mtry), P.i(0), P.i(1), P.i(nodesize), P.i(nodedepth),
        P.i(0), P.d(np.zeros(0)),   # crWeightSize, crWeight
        P.i(ntree), P.i(n), P.i(1), P.s(["C"]), P.i(2),
        P.d(np.asarray(y, np.float64)),
        P.i(p), P.s(list(xtype)), P.i(xlevels),
        P.d(np.asarray(risk_time, np.float64)), P.i(np.asarray(stratifier, np.int32)),
        P.i(int(np.max(stratifier))),
        P.i(np.asarray(var_categories, np.int32)), P.d(np.asarray(category_weights, np.float64)), P.i(ncat),   # varCategories, categoryWeights, nCategories
        P.i(np.zeros(0, np.int32)), P.i(0),               # toFix, nToFix
        P.d(k_for_alpha), P.i(max_boot), P.i(boot_sizes), samp_arg,
        P.d(np.ones(n)), P.d(np.asarray(split_weight, np.float64)), P.d(1.0), P.d(np.ones(p)),
        P.d_alias(x.reshape(-1)),
        P.i(0), P.d(np.zeros(0)), P.i(1), P.i(ntree), P.i(nthreads),
    ]
    assert len(args) == 44
    arr = (C.c_void_p * 44)(*args)
    msg = C.create_string_buffer(2048)
    fn = C.cast(L.rfsrcGrow, C.c_void_p)
    out = L.fr_guard_call(fn, 44, arr, msg, 2048)
    if not out:
        L.fr_release_all()
        raise RuntimeError(f"rfsrcGrow ({impl}) failed: " + msg.value.decode(errors="replace"))
    res = _unpack(L, out)
    L.fr_release_all()
    res["_maxBootstrapSize"] = np.array([max_boot], np.int64)
    res["_nodesize"] = np.array([nodesize], np.int64)
    return res


def trim_forest(res: Dict[str, np.ndarray], n: int) -> Dict[str, np.ndarray]:
    """What R/rfsrc.R:1766-1861 does with the raw native output: cut the forest arrays to
    sum(2*leafCount-1) records and the per-leaf arrays to leafCount[b] entries per tree (the native
    stride is treeTheoreticalMaximum = maxBootstrapSize - 2*(nodesize-1))."""
    leaf = res["leafCount"].astype(np.int64)
    ntree = leaf.size
    total = int(np.sum(2 * leaf[leaf > 0] - 1))
    out = dict(res)
    for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ"):
        out[k] = res[k][:total]
    if "mtryST" in res and res["mtryST"].size:
        mtry = res["mtryST"].size // res["spltST"].size
        ids = res["mtryID"].reshape(-1, mtry)[:total]
        st = res["mtryST"].reshape(-1, mtry)[:total]
        hit = ids == out["parmID"][:, None]
        stat = np.full(total, np.nan)
        rows = hit.any(axis=1) & (out["parmID"] > 0)
        stat[rows] = st[rows, hit[rows].argmax(axis=1)]
        out["splitStat"] = stat
    stride = int(res["_maxBootstrapSize"][0]) - 2 * (int(res["_nodesize"][0]) - 1)
    for k in ("tnRCNT", "tnACNT"):
        if k in res:
            full = res[k]
            parts = [full[b * stride: b * stride + leaf[b]] for b in range(ntree)]
            out[k] = np.concatenate(parts) if parts else full[:0]
    if "mwcpPT" in res and "mwcpCT" in res:
        # mwcpPT is allocated at its theoretical maximum; the words in use are the first sum(mwcpCT), written in the order
        # trees were saved (thread-completion order, like the records)
        out["mwcpPT"] = res["mwcpPT"][: int(res["mwcpCT"].sum())]
    if "tnCLAS" in res and res["tnCLAS"].size:
        full = res["tnCLAS"]
        parts = [full[b * stride * 2: b * stride * 2 + 2 * leaf[b]] for b in range(ntree)]
        out["tnCLAS"] = np.concatenate(parts)
    return out


def per_tree(res: Dict[str, np.ndarray]) -> Dict[int, Dict[str, np.ndarray]]:
    """Key the (trimmed) pre-order forest records by treeID: the reference appends trees in
    thread-completion order (randomForestSRC.c:20938-20946)."""
    trees: Dict[int, Dict[str, np.ndarray]] = {}
    tid = res["treeID"]
    for b in np.unique(tid):
        sel = tid == b
        trees[int(b)] = {k: res[k][sel] for k in ("nodeID", "parmID", "contPT", "mwcpSZ")}
    return trees


def predict(x: np.ndarray, y: np.ndarray, forest: Dict[str, np.ndarray], fx: Optional[np.ndarray], *,
            k_for_alpha: float = 2.0, membership: bool = True, samp: Optional[np.ndarray] = None,
            nodesize: int = 5, impl: str = "ref", quants: bool = False,
            xtype: Optional[Sequence[str]] = None, xlevels: Optional[np.ndarray] = None,
            importance: Optional[str] = None, xvar_imp: Optional[Sequence[int]] = None, seed: int = -1,
            perf: str = "misclass", rfq: bool = False, nthreads: int = -1, fy: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """Run the reference's rfsrcPredict on a (trimmed) grow output.  ``fx`` [p, m_test] selects
    RF_PRED; ``fx=None`` is restore mode (RF_REST) (randomForestSRC.c:22361).  ``fy`` [m_test] = held-out responses (frSize = 1):
    the reference then scores the full ensemble on the held-out rows (``perfClas``)."""
    L = lib(impl)
    p, n = x.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    ntree = forest["leafCount"].size
    P = _Packer(impl)
    opt_low = (1 << 5) + (1 << 2)
    opt_low += {"misclass": 0, "brier": 1 << 3, "gmean": 1 << 14}[perf] + ((1 << 15) if rfq else 0)
    opt_low += importance_bits(importance)
    intr = np.asarray(xvar_imp if xvar_imp is not None else (np.arange(1, p + 1) if importance_bits(importance) else []), np.int32)   # 1-based
    opt_high = (1 << 17)
    if quants:
        opt_high += (1 << 19)                  # terminal.quants incoming: tnCLAS is passed (R/utilities.R:462-483)
    if membership:
        opt_high += (1 << 6)
    if samp is not None:
        opt_low += (1 << 19) + (1 << 20)
        samp = np.ascontiguousarray(samp, dtype=np.int32)
        boot_sizes = samp.sum(axis=1).astype(np.int32)
        max_boot = int(boot_sizes.max())
        samp_arg = P.i_alias(samp.reshape(-1))
    else:
        boot_sizes = np.full(ntree, n, np.int32)
        max_boot = n
        samp_arg = P.i(np.zeros(0, np.int32))
    total = forest["treeID"].size
    has_mwcp = "mwcpPT" in forest and np.asarray(forest["mwcpPT"]).size > 0
    hc_zero = P.lst([P.i(forest["parmID"]), P.d(forest["contPT"]), P.i(forest["mwcpSZ"]),
                     P.i(np.asarray(forest["mwcpPT"], np.int32)) if has_mwcp else P.nil()])
    if xtype is None:
        xtype = ["R"] * p
    if xlevels is None:
        xlevels = np.zeros(p, np.int32)
    partial = P.lst([P.i(0), P.i(0), P.i(0), P.nil(), P.i(0), P.nil(), P.nil()])
    if fx is not None:
        fx = np.ascontiguousarray(fx, dtype=np.float64)
        m_test = fx.shape[1]
        fx_arg = P.d_alias(fx.reshape(-1))
    else:
        m_test = 0
        fx_arg = P.d(np.zeros(0))
    z_i = lambda: P.i(np.zeros(0, np.int32))
    z_d = lambda: P.d(np.zeros(0))
    args = [
        P.i(0), P.i(seed), P.i(opt_low), P.i(opt_high), P.i(ntree), P.i(n), P.i(1), P.s(["C"]), P.i(2),
        P.d(np.asarray(y, np.float64)), P.i(p), P.s(list(xtype)), P.i(np.asarray(xlevels, np.int32)),
        P.d_alias(x.reshape(-1)),
        P.d(k_for_alpha), P.i(max_boot), P.i(boot_sizes), samp_arg, P.d(np.ones(n)),
        P.i(0), z_d(), P.i(total), P.i(forest["seed"]), P.i(0),
        P.i(forest["treeID"]), P.i(forest["nodeID"]),
        hc_zero, P.nil(), P.nil(), P.nil(), P.nil(), P.nil(), P.nil(),
        P.i(forest["rmbrMembership"]), P.i(forest["ambrMembership"]),
        P.i(forest["tnRCNT"]), P.i(forest["tnACNT"]),
        z_d(), z_d(), z_d(), z_d(), z_d(), z_d(),
        P.i(forest["tnCLAS"]) if (quants and "tnCLAS" in forest) else z_i(),
        P.i(1), P.i(1), P.i(0),
        P.i(int(intr.size)), P.i(intr), partial, P.i(0), z_i(),
        P.i(m_test), P.i(0 if fy is None else 1), z_d() if fy is None else P.d(np.asarray(fy, np.float64)), fx_arg, P.i(ntree), P.i(nthreads),
    ]
    assert len(args) == 58, len(args)
    arr = (C.c_void_p * 58)(*args)
    msg = C.create_string_buffer(2048)
    fn = C.cast(L.rfsrcPredict, C.c_void_p)
    # rfsrcPredict ignores its numThreads argument ("RF_numThreads = 2;//INTEGER(numThreads)[0]", randomForestSRC.c:22318) and its
    # per-tree importance shares buffers between trees (randomForestSRC.c:2265-2280): one thread has to be forced through OpenMP
    gomp = prev = None
    if nthreads == 1 and impl == "ref":
        try:
            gomp = C.CDLL("libgomp.so.1"); prev = gomp.omp_get_max_threads(); gomp.omp_set_num_threads(1)
        except OSError:
            gomp = None
    try:
        out = L.fr_guard_call(fn, 58, arr, msg, 2048)
    finally:
        if gomp is not None:
            gomp.omp_set_num_threads(prev)
    if not out:
        L.fr_release_all()
        raise RuntimeError(f"rfsrcPredict ({impl}) failed: " + msg.value.decode(errors="replace"))
    res = _unpack(L, out)
    L.fr_release_all()
    return res
