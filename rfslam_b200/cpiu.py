"""CPIU construction on the device (rfslam_b200_cpiu_build): the host-side mirror of `cpiu.fc` (utilities/cpiu.R:361-668).

``to_create`` follows the reference's naming convention (cpiu.R:504-516): ``i.x`` indicator, ``ni.x`` count in the
current interval, ``n.x`` count over previous intervals, ``t.x`` time elapsed since the event, ``c.x`` carried-forward
measurement.  Values are ``[persons, width]`` arrays of event times (NaN = no event in that slot); ``c.x`` takes a dict
``{"times": ..., "values": ...}``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping

import numpy as np

from . import capi

KINDS = {"i": 0, "n": 1, "ni": 2, "t": 3, "c": 4}
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class CpiuArgs(C.Structure):
    _fields_ = [("persons", C.c_int32), ("t_outcome", _pd), ("t_death", _pd), ("death", _pi), ("interval", C.c_double),
                ("unit_norm", C.c_double), ("n_vars", C.c_int32), ("var_kind", _pi), ("var_width", _pi),
                ("var_times", C.POINTER(_pd)), ("var_values", C.POINTER(_pd)), ("k_to_persist", C.c_int32), ("device", C.c_int32)]


class CpiuOut(C.Structure):
    _fields_ = [("rows", C.c_int64), ("n_vars", C.c_int32), ("person", _pi), ("int_n", _pi), ("t_start", _pd), ("t_end", _pd),
                ("rt", _pd), ("i_death", _pi), ("vars", _pd), ("kernel_ms", C.c_double), ("kernel_launches", C.c_int64)]


def _take(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype)


def cpiu_fc(t_outcome, t_death, death, *, interval: float = 0.5, unit_norm: float = 365.0,
            to_create: Mapping[str, object] = (), k_to_persist: int = 0, device: int = -1) -> Dict[str, np.ndarray]:
    L = capi.lib()
    L.rfslam_b200_cpiu_build.restype = C.c_int; L.rfslam_b200_cpiu_build.argtypes = [C.POINTER(CpiuArgs), C.POINTER(CpiuOut)]
    L.rfslam_b200_free_cpiu.restype = None; L.rfslam_b200_free_cpiu.argtypes = [C.POINTER(CpiuOut)]
    to = np.ascontiguousarray(t_outcome, np.float64); td = np.ascontiguousarray(t_death, np.float64)
    dth = np.ascontiguousarray(death, np.int32)
    P = to.size
    if td.shape != (P,) or dth.shape != (P,):
        raise ValueError("t_death and death must hold one value per person")
    names, kinds, widths, times, values = [], [], [], [], []
    for name, spec in dict(to_create).items():
        prefix = name.split(".", 1)[0]
        if prefix not in KINDS or "." not in name:
            raise ValueError("Variables do not conform to naming convention (i. / ni. / n. / t. / c.): " + name)   # cpiu.R:504-528
        kind = KINDS[prefix]
        if kind == 4:
            tm = np.ascontiguousarray(np.asarray(spec["times"], np.float64).reshape(P, -1))
            vl = np.ascontiguousarray(np.asarray(spec["values"], np.float64).reshape(P, -1))
            if vl.shape != tm.shape:
                raise ValueError(name + ": times and values must have one shape")
        else:
            tm = np.ascontiguousarray(np.asarray(spec, np.float64).reshape(P, -1)); vl = None
        names.append(name); kinds.append(kind); widths.append(tm.shape[1]); times.append(tm); values.append(vl)
    V = len(names)
    a = CpiuArgs()
    a.persons = P; a.t_outcome = to.ctypes.data_as(_pd); a.t_death = td.ctypes.data_as(_pd); a.death = dth.ctypes.data_as(_pi)
    a.interval = float(interval); a.unit_norm = float(unit_norm); a.n_vars = V; a.k_to_persist = int(k_to_persist); a.device = int(device)
    kk = np.asarray(kinds, np.int32); ww = np.asarray(widths, np.int32)
    tp = (_pd * max(V, 1))(*[t.ctypes.data_as(_pd) for t in times])
    vp = (_pd * max(V, 1))(*[(v.ctypes.data_as(_pd) if v is not None else None) for v in values])
    a.var_kind = kk.ctypes.data_as(_pi); a.var_width = ww.ctypes.data_as(_pi); a.var_times = tp; a.var_values = vp
    out = CpiuOut()
    capi.check(L.rfslam_b200_cpiu_build(C.byref(a), C.byref(out)))
    R = int(out.rows)
    res = {"person": _take(out.person, R, np.int32), "int_n": _take(out.int_n, R, np.int32), "t_start": _take(out.t_start, R, np.float64),
           "t_end": _take(out.t_end, R, np.float64), "rt": _take(out.rt, R, np.float64), "i_death": _take(out.i_death, R, np.int32),
           "vars": _take(out.vars, R * V, np.float64).reshape(V, R), "names": names, "kernel_ms": float(out.kernel_ms)}
    L.rfslam_b200_free_cpiu(C.byref(out))
    return res
