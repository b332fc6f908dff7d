"""CPU suite for the host side: the C-ABI library loads, exports every symbol the header declares,
mirrors the reference's slot registry, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import rfslam_b200
from rfslam_b200 import capi
from rfslam_b200.synth import make_cpiu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rfslam_b200.h")).read()
    declared = set(re.findall(r"\b(rfslam_b200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS)
    L = ctypes.CDLL(capi.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    # the reference's plugin symbols are exported under their own names (splitCustom.h:24-44, splitCustom.c:23-67)
    for s in capi.REFERENCE_SYMBOLS:
        assert re.search(r"\b" + s + r"\b", hdr), s
        assert hasattr(L, s), s


def test_register_this_maps_function_pointers_like_the_reference():
    # registerThis(&poissonSplitN, CLAS_FAM, slot) (splitCustom.c:47-62): the six known pointers select device rules,
    # any other function pointer is a CPU rule -> the slot is foreign and a grow that selects it is refused
    L = capi.lib()
    L.rfslam_b200_register_defaults()
    addr = lambda k: ctypes.cast(getattr(L, f"poissonSplit{k}"), ctypes.c_void_p).value
    L.registerThis(addr(4), 0, 9)
    assert L.rfslam_b200_rule_of_slot(0, 9) == 4
    L.registerThis(addr(2), 0, 9)
    assert L.rfslam_b200_rule_of_slot(0, 9) == 2

    @capi.CUSTOM_FUNCTION
    def user_rule(*a):
        return 0.0
    L.registerThis(ctypes.cast(user_rule, ctypes.c_void_p).value, 0, 10)
    assert L.rfslam_b200_rule_of_slot(0, 10) == 0
    d = make_cpiu(200, 3, levels=8, seed=1)
    pk = rfslam_b200.forest._Packed(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=2, mtry=None, nodesize=5, nodedepth=-1, seed=-3,
                                    k_for_alpha=2.0, membership=False, samp=None, sampsize=None, split_wt=None, device=-1,
                                    tree_first=0, tree_step=0, finalize=True)
    pk.args.opt_high = (pk.args.opt_high & ~0x0F00) | (256 * (10 - 1))
    out = capi.ForestOut()
    rc = L.rfslam_b200_grow(ctypes.byref(pk.args), ctypes.byref(out))
    assert rc == 3 and b"user-supplied" in L.rfslam_b200_last_error()
    L.registerThis(addr(4), 1, 9)                    # a Poisson rule in the regression family: no device rule there
    assert L.rfslam_b200_rule_of_slot(1, 9) == 0
    L.rfslam_b200_register_defaults()
    L.registerCustomFunctions()
    assert [L.rfslam_b200_rule_of_slot(0, s) for s in range(1, 9)] == [0, 1, 2, 3, 4, 5, 6, 0]


def test_slot_registry_mirrors_reference():
    # splitCustom.c:47-62: poissonSplitN sits in CLAS_FAM slot N+1; slot 1 is the stock example rule
    L = capi.lib()
    L.rfslam_b200_register_defaults()
    assert [L.rfslam_b200_rule_of_slot(0, s) for s in range(1, 9)] == [0, 1, 2, 3, 4, 5, 6, 0]
    assert L.rfslam_b200_rule_of_slot(1, 2) == 0          # regression slots abort in the reference too
    assert L.rfslam_b200_register_this(4, 0, 9) == 0
    assert L.rfslam_b200_rule_of_slot(0, 9) == 4
    assert L.rfslam_b200_register_this(4, 1, 9) != 0
    L.rfslam_b200_register_defaults()
    assert L.rfslam_b200_rule_of_slot(0, 9) == 0


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors against the header itself: a C program compiled from include/rfslam_b200.h prints the
    size of every struct and the offset of every field; the ctypes classes must agree field by field."""
    import subprocess
    structs = {"rfslam_grow_args": capi.GrowArgs, "rfslam_forest_out": capi.ForestOut,
               "rfslam_predict_args": capi.PredictArgs, "rfslam_predict_out": capi.PredictOut}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rfslam_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    # and no field of the header is missing from the mirror (sizes equal + last field at the same offset covers it)


@pytest.mark.skipif(capi.lib().rfslam_b200_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback():
    d = make_cpiu(200, 3, levels=8, seed=1)
    with pytest.raises(rfslam_b200.RfslamError) as e:
        rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=2)
    assert e.value.code == 2


def test_argument_validation_precedes_device_work():
    d = make_cpiu(200, 3, levels=8, seed=1)
    with pytest.raises(rfslam_b200.RfslamError) as e:
        rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=2, seed=7)
    assert e.value.code == 1 and "seed" in str(e.value).lower()
    with pytest.raises(ValueError):
        rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=2, splitrule="gini")


def test_synth_is_deterministic_and_shaped():
    a = make_cpiu(5000, 9, levels=64)
    b = make_cpiu(5000, 9, levels=64)
    assert a.x.shape == (9, 5000) and np.array_equal(a.x, b.x) and np.array_equal(a.y, b.y)
    assert set(np.unique(a.y)) <= {1.0, 2.0} and a.strat.min() >= 1 and a.strat.max() <= 16 and (a.rt > 0).all()
    assert 0.01 < (a.y == 2).mean() < 0.15
    assert max(np.unique(c).size for c in a.x) <= 64


def test_importance_requests_are_validated_before_device_work(monkeypatch):
    """Only permutation importance is on the accelerated path: the other types of R/utilities.R:150-208 (anti = no type
    bit, random = 2^9, joint = 2^10) and importance on held-out rows are refused by the argument check, GPU or not."""
    from rfslam_b200 import forest as fmod
    d = make_cpiu(300, 4, levels=8, seed=2)
    stump = dict(treeID=np.ones(1, np.int32), nodeID=np.ones(1, np.int32), parmID=np.zeros(1, np.int32), contPT=np.zeros(1),
                 mwcpSZ=np.zeros(1, np.int32), seed=np.array([-5], np.int32), tnCLAS=np.array([3, 1], np.int32))
    for name, bits, what in (("anti", (1 << 25) + (1 << 24), "permutation"), ("random", (1 << 25) + (1 << 24) + (1 << 9), "permutation"),
                             ("permute.joint", (1 << 25) + (1 << 24) + (1 << 10) + (1 << 8), "joint")):
        monkeypatch.setitem(fmod.IMPORTANCE_BITS, name, bits)
        with pytest.raises(rfslam_b200.RfslamError, match=what) as e:
            rfslam_b200.predict_rfsrc(stump, None, x=d.x, y=d.y, importance=name)
        assert e.value.code == 3                                      # RFSLAM_ENOSUP
    with pytest.raises(rfslam_b200.RfslamError, match="held-out"):
        rfslam_b200.predict_rfsrc(stump, d.x[:, :10].copy(), x=d.x, y=d.y, importance="permute")
    with pytest.raises(rfslam_b200.RfslamError, match="intrPredictor"):
        rfslam_b200.predict_rfsrc(stump, None, x=d.x, y=d.y, importance="permute", xvar_importance=[0, 9])
