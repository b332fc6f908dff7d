"""ctypes wrapper of the CPU restatement (oracle/rfslam_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "librfslam_oracle.so")
_lib = None

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)


class OrcArgs(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n", "p", "ntree", "mtry", "nodesize", "nodedepth", "rule", "seed", "max_boot")] + [
        ("k_for_alpha", C.c_double),
        ("x", _pd), ("y", _pd), ("rt", _pd), ("strat", _pi), ("samp", _pi), ("boot_size", _pi),
        ("split_weight", _pd), ("tree_seeds", _pi), ("nsplit", C.c_int), ("xtype", C.c_char_p), ("xlevels", _pi)]


class OrcOut(C.Structure):
    _fields_ = [("ntree", C.c_int), ("n", C.c_int),
                ("total_nodes", C.c_long), ("total_leaves", C.c_long), ("candidates", C.c_long), ("rmbr_stride", C.c_long),
                ("treeID", _pi), ("nodeID", _pi), ("parmID", _pi), ("contPT", _pd), ("splitStat", _pd),
                ("mwcpSZ", _pi), ("mwcpPT", C.POINTER(C.c_uint)), ("mwcpCT", _pi), ("mwcp_total", C.c_long),
                ("leafCount", _pi), ("seed", _pi), ("bootMembership", _pi), ("nodeMembership", _pi),
                ("tnRCNT", _pi), ("tnACNT", _pi), ("tnCLAS", _pi), ("rmbr", _pi), ("ambr", _pi),
                ("oobNum", _pd), ("allNum", _pd), ("oobDen", _pi), ("allDen", _pi), ("oobEns", _pd), ("allEns", _pd)]


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "rfslam_oracle.c")
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(PORT_SO)
        L.orc_grow.restype = C.POINTER(OrcOut); L.orc_grow.argtypes = [C.POINTER(OrcArgs)]
        L.orc_free.restype = None; L.orc_free.argtypes = [C.POINTER(OrcOut)]
        L.orc_poisson_stat.restype = C.c_double
        L.orc_poisson_stat.argtypes = [C.c_int, C.c_int, C.c_double, C.c_char_p, _pd, _pd, _pd]
        L.orc_chain_seeds.restype = None; L.orc_chain_seeds.argtypes = [C.c_int, C.c_int, _pi, _pi]
        L.orc_chain_seeds_restore.restype = None; L.orc_chain_seeds_restore.argtypes = [C.c_int, _pi]
        L.orc_bootstrap.restype = None; L.orc_bootstrap.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _pi]
        L.orc_ran1_stream.restype = None; L.orc_ran1_stream.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.orc_predict.restype = None
        L.orc_predict.argtypes = [C.c_int, C.c_int, C.c_long, _pi, _pi, _pi, _pd, _pi, _pi, _pi, _pd, C.c_int, _pd, _pi, _pi, C.POINTER(C.c_uint)]
        _lib = L
    return _lib


def _dp(a): return a.ctypes.data_as(_pd)
def _ip(a): return a.ctypes.data_as(_pi)


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def grow(x, y, rt, strat, *, rule=4, ntree=4, mtry=None, nodesize=5, nodedepth=-1, seed=-12345,
         k_for_alpha=2.0, samp=None, split_weight=None, tree_seeds=None, nsplit=0, sampsize=None,
         xtype=None, xlevels=None, perf="misclass", rfq=False) -> Dict[str, np.ndarray]:
    L = lib()
    p, n = x.shape
    perf_mode = {"misclass": 0, "brier": 1, "gmean": 2}[perf]
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    rt = np.ascontiguousarray(rt, np.float64); strat = np.ascontiguousarray(strat, np.int32)
    if mtry is None:
        mtry = int(np.ceil(np.sqrt(p)))
    a = OrcArgs()
    a.n, a.p, a.ntree, a.mtry, a.nodesize, a.nodedepth, a.rule, a.seed = n, p, ntree, mtry, nodesize, nodedepth, rule, seed
    a.k_for_alpha = k_for_alpha
    a.nsplit = int(nsplit)
    a.x, a.y, a.rt, a.strat = _dp(x), _dp(y), _dp(rt), _ip(strat)
    keep = [x, y, rt, strat]
    a.max_boot = n
    if samp is not None:
        samp = np.ascontiguousarray(samp, np.int32); keep.append(samp)
        a.samp = _ip(samp); a.max_boot = int(samp.sum(axis=1).max())
    if sampsize is not None and samp is None:
        bs = np.full(ntree, int(sampsize), np.int32); keep.append(bs)
        a.boot_size = _ip(bs); a.max_boot = int(sampsize)
    if split_weight is not None:
        sw = np.ascontiguousarray(split_weight, np.float64); keep.append(sw); a.split_weight = _dp(sw)
    if tree_seeds is not None:
        ts = np.ascontiguousarray(tree_seeds, np.int32); keep.append(ts); a.tree_seeds = _ip(ts)
    if xtype is not None:
        xt = "".join(xtype).encode(); xl = np.ascontiguousarray(xlevels, np.int32); keep += [xt, xl]
        a.xtype = xt; a.xlevels = _ip(xl)
    op = L.orc_grow(C.byref(a))
    o = op.contents
    tn, tl = o.total_nodes, o.total_leaves
    res = {
        "treeID": _arr(o.treeID, tn, np.int32), "nodeID": _arr(o.nodeID, tn, np.int32),
        "parmID": _arr(o.parmID, tn, np.int32), "contPT": _arr(o.contPT, tn, np.float64),
        "splitStat": _arr(o.splitStat, tn, np.float64), "mwcpSZ": _arr(o.mwcpSZ, tn, np.int32),
        "mwcpPT": _arr(o.mwcpPT, o.mwcp_total, np.uint32).astype(np.int32), "mwcpCT": _arr(o.mwcpCT, ntree, np.int32),
        "leafCount": _arr(o.leafCount, ntree, np.int32), "seed": _arr(o.seed, ntree, np.int32),
        "bootMembership": _arr(o.bootMembership, ntree * n, np.int32),
        "nodeMembership": _arr(o.nodeMembership, ntree * n, np.int32),
        "tnRCNT": _arr(o.tnRCNT, tl, np.int32), "tnACNT": _arr(o.tnACNT, tl, np.int32),
        "tnCLAS": _arr(o.tnCLAS, 2 * tl, np.int32),
        "rmbrMembership": _arr(o.rmbr, ntree * o.rmbr_stride, np.int32),
        "ambrMembership": _arr(o.ambr, ntree * n, np.int32),
        "oobEnsbCLS": _arr(o.oobEns, 2 * n, np.float64), "allEnsbCLS": _arr(o.allEns, 2 * n, np.float64),
        "candidates": np.array([o.candidates], np.int64),
    }
    # perfClas [ntree][3]: NA_REAL rows except the last (perfBlock = ntree, randomForestSRC.c:8155)
    na = np.array([0x7FF00000000007A2], np.uint64).view(np.float64)[0]
    perf = np.full(ntree * 3, na, np.float64)
    last = np.zeros(3, np.float64)
    oob_den = _arr(o.oobDen, n, np.int32)
    L.orc_perf_clas2.restype = None
    L.orc_perf_clas2.argtypes = [C.c_int, _pd, _pd, _pi, C.c_double, C.c_int, C.c_int, _pd]
    L.orc_perf_clas2(n, _dp(y), _dp(res["oobEnsbCLS"]), _ip(oob_den), C.c_double(na), perf_mode, int(bool(rfq)), _dp(last))
    perf[-3:] = last
    res["perfClas"] = perf
    L.orc_free(op)
    return res


def poisson_stat(rule, k_for_alpha, membership, response, strat, rt) -> float:
    """One plugin call, 0-based arrays; membership 1 = LEFT (splitCustom.h:5-6)."""
    L = lib()
    mem = np.ascontiguousarray(membership, np.int8)
    r = np.ascontiguousarray(response, np.float64); s = np.ascontiguousarray(strat, np.float64)
    t = np.ascontiguousarray(rt, np.float64)
    return L.orc_poisson_stat(rule, r.size, k_for_alpha, mem.ctypes.data_as(C.c_char_p), _dp(r), _dp(s), _dp(t))


def chain_seeds(seed: int, ntree: int):
    L = lib()
    a = np.zeros(ntree, np.int32); b = np.zeros(ntree, np.int32)
    L.orc_chain_seeds(seed, ntree, _ip(a), _ip(b))
    return a, b


def bootstrap(seed: int, ntree: int, n: int, tree: int) -> np.ndarray:
    L = lib()
    d = np.zeros(n, np.int32)
    L.orc_bootstrap(seed, ntree, n, tree, _ip(d))
    return d


def ran1_stream(seed: int, count: int) -> np.ndarray:
    L = lib()
    out = np.zeros(count, np.float32)
    L.orc_ran1_stream(seed, count, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def predict(forest: Dict[str, np.ndarray], fx: np.ndarray):
    """forest: trimmed arrays in tree order (treeID ascending) + leafCount/tnRCNT/tnCLAS."""
    L = lib()
    fx = np.ascontiguousarray(fx, np.float64)
    p, m = fx.shape
    ntree = forest["leafCount"].size
    tid = np.ascontiguousarray(forest["treeID"], np.int32); nid = np.ascontiguousarray(forest["nodeID"], np.int32)
    pid = np.ascontiguousarray(forest["parmID"], np.int32); cpt = np.ascontiguousarray(forest["contPT"], np.float64)
    lc = np.ascontiguousarray(forest["leafCount"], np.int32); rc = np.ascontiguousarray(forest["tnRCNT"], np.int32)
    te = np.ascontiguousarray(forest["tnCLAS"].reshape(-1, 2)[:, 1], np.int32)
    ens = np.zeros(2 * m, np.float64); mem = np.zeros(ntree * m, np.int32)
    msz = np.ascontiguousarray(forest.get("mwcpSZ", np.zeros(tid.size, np.int32)), np.int32)
    mpt = np.ascontiguousarray(forest.get("mwcpPT", np.zeros(0, np.int32)), np.int32).view(np.uint32)
    if mpt.size == 0: mpt = np.zeros(1, np.uint32)
    L.orc_predict(ntree, p, tid.size, _ip(tid), _ip(nid), _ip(pid), _dp(cpt), _ip(lc), _ip(rc), _ip(te), _dp(fx), m, _dp(ens), _ip(mem),
                  _ip(msz), mpt.ctypes.data_as(C.POINTER(C.c_uint)))
    return ens, mem


def chain_seed_c(seed: int, ntree: int, restore: bool = False) -> np.ndarray:
    """Chain-C seeds (the VIMP chain): third block of the seed LCG in grow mode, second block from 0 in restore mode."""
    L = lib()
    c = np.zeros(ntree, np.int32)
    L.orc_chain_seed_c(int(seed), ntree, 1 if restore else 0, _ip(c))
    return c


def vimp_permute(forest: Dict[str, np.ndarray], x, y, boot, seed_c, *, intr=None, per_tree: bool = True,
                 perf: str = "misclass", rfq: bool = False) -> np.ndarray:
    """Permutation importance [len(intr)][3] of a trimmed forest in tree order (treeID ascending); ``boot`` [ntree][n]
    in-bag counts, ``forest['tnCLAS']`` the leaves' class counts."""
    L = lib()
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    p, n = x.shape
    ntree = forest["leafCount"].size
    tid = np.ascontiguousarray(forest["treeID"], np.int32); nid = np.ascontiguousarray(forest["nodeID"], np.int32)
    pid = np.ascontiguousarray(forest["parmID"], np.int32); cpt = np.ascontiguousarray(forest["contPT"], np.float64)
    lc = np.ascontiguousarray(forest["leafCount"], np.int32); tc = np.ascontiguousarray(forest["tnCLAS"], np.int32)
    msz = np.ascontiguousarray(forest.get("mwcpSZ", np.zeros(tid.size, np.int32)), np.int32)
    mpt = np.ascontiguousarray(forest.get("mwcpPT", np.zeros(0, np.int32)), np.int32).view(np.uint32)
    if mpt.size == 0: mpt = np.zeros(1, np.uint32)
    boot = np.ascontiguousarray(boot, np.int32).reshape(ntree, n)
    sc = np.ascontiguousarray(seed_c, np.int32)
    intr = np.ascontiguousarray(np.arange(1, p + 1) if intr is None else intr, np.int32)
    out = np.zeros(intr.size * 3, np.float64)
    na = np.frombuffer(np.uint64(0x7FF00000000007A2).tobytes(), np.float64)[0]
    L.orc_vimp_permute.restype = None
    L.orc_vimp_permute.argtypes = [C.c_int, C.c_int, C.c_int, C.c_long, _pi, _pi, _pi, _pd, _pi, _pi, _pi, C.POINTER(C.c_uint), _pd, _pd, _pi, _pi,
                                   C.c_int, _pi, C.c_int, C.c_int, C.c_int, C.c_double, _pd]
    L.orc_vimp_permute(ntree, p, n, tid.size, _ip(tid), _ip(nid), _ip(pid), _dp(cpt), _ip(lc), _ip(tc), _ip(msz),
                       mpt.ctypes.data_as(C.POINTER(C.c_uint)), _dp(x), _dp(y), _ip(boot), _ip(sc), int(intr.size), _ip(intr),
                       1 if per_tree else 0, {"misclass": 0, "brier": 1, "gmean": 2}[perf], 1 if rfq else 0, na, _dp(out))
    return out.reshape(-1, 3)


def perf_clas(y, ens, den=None, *, perf: str = "misclass", rfq: bool = False) -> np.ndarray:
    """One performance row [overall, class 1, class 2] of a two-class ensemble ``ens`` [2][n] against ``y`` (getPerformance's
    class branch).  With ``rfq`` the minority class / threshold are taken from ``y`` itself (true for the training rows)."""
    L = lib()
    y = np.ascontiguousarray(y, np.float64); ens = np.ascontiguousarray(ens, np.float64).reshape(-1)
    n = y.size
    den = np.ones(n, np.int32) if den is None else np.ascontiguousarray(den, np.int32)
    na = np.frombuffer(np.uint64(0x7FF00000000007A2).tobytes(), np.float64)[0]
    out = np.zeros(3, np.float64)
    L.orc_perf_clas2.restype = None
    L.orc_perf_clas2.argtypes = [C.c_int, _pd, _pd, _pi, C.c_double, C.c_int, C.c_int, _pd]
    L.orc_perf_clas2(n, _dp(y), _dp(ens), _ip(den), C.c_double(na), {"misclass": 0, "brier": 1, "gmean": 2}[perf], int(bool(rfq)), _dp(out))
    return out
