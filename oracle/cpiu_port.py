"""CPU restatement of the reference's CPIU construction (utilities/cpiu.R:455-668 `event`, utilities/cpiu.cpp:27-249).

TEST INFRASTRUCTURE ONLY: imported by tests/ and bench.py's cpu_baseline leg -- never by the product package.

The four per-person helpers of cpiu.cpp are restated in plain Python and PINNED against the unmodified reference file
compiled over a stand-in Rcpp.h (oracle/_ref/libcpiu_ref.so, `make -C oracle cpiu`; tests/test_cpiu.py).  The R driver
around them (interval table: int.n, t.start, t.end, rt, i.death, cpiu.R:567-577; NaN / -1 handling cpiu.R:583-584, :642)
is restated from the R source; R is not in the image, so that part is pinned by reading only ("parity unpinned" for
the R arithmetic, which is five ifelse lines).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libcpiu_ref.so")
INDICATOR, COUNT_PREV, COUNT_CUR, ELAPSED, CONTINUOUS = range(5)
_ref = None


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def _lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.cpiu_ref_indicator.argtypes = [pd, C.c_int, pd, pd, C.c_int, pi]
        L.cpiu_ref_accumulator.argtypes = [pd, C.c_int, pd, pd, C.c_int, C.c_int, pi]
        L.cpiu_ref_time_elapsed.argtypes = [pd, C.c_int, pd, pd, C.c_int, pd]
        L.cpiu_ref_continuous.argtypes = [pd, pd, C.c_int, pd, pd, C.c_int, C.c_int, pd]
        _ref = L
    return _ref


def _dp(a): return a.ctypes.data_as(C.POINTER(C.c_double))


# ---- the reference helpers run from the compiled, unmodified cpiu.cpp -------------------------------------------------
def ref_helper(kind: int, times, values, ts, te, k_to_persist: int = 0) -> np.ndarray:
    L = _lib()
    ev = np.ascontiguousarray(times, np.float64); ts = np.ascontiguousarray(ts, np.float64); te = np.ascontiguousarray(te, np.float64)
    ni = ts.size
    if kind in (INDICATOR, COUNT_PREV, COUNT_CUR):
        out = np.zeros(ni, np.int32); po = out.ctypes.data_as(C.POINTER(C.c_int))
        if kind == INDICATOR:
            L.cpiu_ref_indicator(_dp(ev), ev.size, _dp(ts), _dp(te), ni, po)
        else:
            L.cpiu_ref_accumulator(_dp(ev), ev.size, _dp(ts), _dp(te), ni, 1 if kind == COUNT_PREV else 0, po)
        return out.astype(np.float64)
    out = np.zeros(ni, np.float64)
    if kind == ELAPSED:
        L.cpiu_ref_time_elapsed(_dp(ev), ev.size, _dp(ts), _dp(te), ni, _dp(out))
    else:
        mv = np.ascontiguousarray(values, np.float64)
        L.cpiu_ref_continuous(_dp(ev), _dp(mv), ev.size, _dp(ts), _dp(te), ni, int(k_to_persist), _dp(out))
    return out


# ---- plain restatements (cpiu.cpp:27-249) ---------------------------------------------------------------------------
def port_helper(kind: int, times, values, ts, te, k_to_persist: int = 0) -> np.ndarray:
    ni = len(ts)
    out = np.zeros(ni, np.float64)
    if kind == INDICATOR:                                   # cpiu.cpp:27-50
        for i in range(ni):
            for e in times:
                if e > ts[i] and e <= te[i]:
                    out[i] = 1
    elif kind in (COUNT_PREV, COUNT_CUR):                   # cpiu.cpp:75-113 (flag 1 / flag 0)
        count = 0
        for i in range(ni):
            if kind == COUNT_CUR:
                count = 0
            else:
                out[i] = count
            for e in times:
                if e > ts[i] and e <= te[i]:
                    count += 1
            if kind == COUNT_CUR:
                out[i] = count
    elif kind == ELAPSED:                                   # cpiu.cpp:129-165
        last = 0.0
        for i in range(ni):
            toggle = False
            for e in times:
                if e > ts[i] and e <= te[i]:
                    last = e; toggle = True
            if not toggle:
                out[i] = te[i] - last
    else:                                                   # cpiu.cpp:186-249
        k = ni if k_to_persist == -1 else k_to_persist
        j_start = 0; age = 0
        out[0] = values[0]
        for i in range(ni - 1):
            changed = False
            for j in range(j_start, len(times)):
                if times[j] > ts[i] and times[j] <= te[i]:
                    out[i + 1] = values[j]; changed = True; j_start = j; age = 0
            if not changed:
                out[i + 1] = -1 if age + 1 >= k else out[i]
                age += 1
    return out


NA_REAL = np.array([0x7FF00000000007A2], np.uint64).view(np.float64)[0]


def cpiu_fc(t_outcome, t_death, death, *, interval: float = 0.5, unit_norm: float = 365.0,
            var_kind: Sequence[int] = (), var_times: Sequence[np.ndarray] = (), var_values: Sequence[Optional[np.ndarray]] = (),
            k_to_persist: int = 0, use_ref: bool = True) -> Dict[str, np.ndarray]:
    """`event` of cpiu.R:455-668 for every person, rows of person 0 first.  var_times[v] is [persons, width] (NaN =
    missing), var_values[v] likewise for CONTINUOUS variables (measurements are sorted by time first, cpiu.R:488-494;
    missing measurements must trail)."""
    helper = ref_helper if (use_ref and ref_available()) else port_helper
    P = len(t_outcome)
    imod = interval * unit_norm                                              # cpiu.R:472
    cols: Dict[str, List] = {k: [] for k in ("person", "int_n", "t_start", "t_end", "rt", "i_death")}
    vcols: List[List[np.ndarray]] = [[] for _ in var_kind]
    for q in range(P):
        ni = int(math.ceil(t_outcome[q] / imod))                             # cpiu.R:476
        n = np.arange(1, ni + 1)
        ts = imod * (n - 1)                                                  # cpiu.R:568
        inside = ts + imod < t_outcome[q]
        te = np.where(inside, ts + imod, t_outcome[q])                       # cpiu.R:569-571
        rt = np.where(inside, interval, (t_outcome[q] - ts) / unit_norm)     # cpiu.R:572-574
        idth = np.where(ts + imod < t_death[q], 0, death[q])                 # cpiu.R:575-577
        cols["person"].append(np.full(ni, q)); cols["int_n"].append(n); cols["t_start"].append(ts); cols["t_end"].append(te)
        cols["rt"].append(rt); cols["i_death"].append(idth)
        for v, kind in enumerate(var_kind):
            tm = np.asarray(var_times[v][q], np.float64)
            if kind == CONTINUOUS:
                vals = np.asarray(var_values[v][q], np.float64)
                keep = ~np.isnan(tm)
                order = np.argsort(tm[keep], kind="stable")                  # cvars.sort, cpiu.R:488-494
                out = helper(kind, tm[keep][order], vals[keep][order], ts, te, k_to_persist)
            else:
                out = helper(kind, np.where(np.isnan(tm), -1.0, tm), None, ts, te)   # cpiu.R:583-584
            out = np.where(out == -1, NA_REAL, out)                          # cpiu.R:642
            vcols[v].append(out)
    res = {k: np.concatenate(v) if v else np.zeros(0) for k, v in cols.items()}
    res["person"] = res["person"].astype(np.int32); res["int_n"] = res["int_n"].astype(np.int32); res["i_death"] = res["i_death"].astype(np.int32)
    res["vars"] = np.stack([np.concatenate(c) for c in vcols]) if vcols else np.zeros((0, res["rt"].size))
    return res
