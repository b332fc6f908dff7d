"""The drop-in boundary under test: the product's `.Call` entry points
(rfslam_b200/csrc/r_shim.c: SEXP rfsrcGrow(44) / SEXP rfsrcPredict(58) over librfslam_b200.so) and the
UNMODIFIED reference, both compiled against the same stand-in R API (oracle/fake_r) and driven by ONE
argument packer (oracle/refdriver.py) -- what comes back must be the same named list."""
import ctypes
import os

import numpy as np
import pytest

from oracle import refdriver
from rfslam_b200 import capi
from rfslam_b200.synth import make_cpiu
from tests import parity

needs_shim = pytest.mark.skipif(not refdriver.available("b200"), reason="oracle/_ref/librfslam_b200_rshim.so not built (make -C oracle rshim)")
needs_ref = pytest.mark.skipif(not refdriver.available("ref"), reason="oracle/_ref/librfslam_ref.so not present")


@needs_shim
def test_shim_exports_the_reference_call_symbols():
    L = ctypes.CDLL(refdriver.SHIM_SO, mode=ctypes.RTLD_LOCAL)
    for s in ("rfsrcGrow", "rfsrcPredict"):                  # src/R_init_randomForestSRC.c:26-31
        assert hasattr(L, s), s


@needs_shim
@pytest.mark.skipif(capi.lib().rfslam_b200_device_count() > 0, reason="a GPU is present")
def test_shim_turns_library_errors_into_r_errors():
    d = make_cpiu(300, 4, levels=8, seed=2)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        refdriver.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=2, impl="b200")


def _sorted_by_tree(t):
    order = np.argsort(t["treeID"], kind="stable")
    out = dict(t)
    for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ"):
        out[k] = t[k][order]
    return out


CASES = [dict(rule=4, levels=32, kw=dict()), dict(rule=1, levels=0, kw=dict()), dict(rule=5, levels=16, kw=dict(byuser=True)),
         dict(rule=4, levels=64, kw=dict(quants=True))]


@pytest.mark.gpu
@needs_shim
@needs_ref
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"r{c['rule']}L{c['levels']}" + "".join(c["kw"]))
def test_both_libraries_return_the_same_named_lists(case):
    d = make_cpiu(3000, 9, levels=case["levels"], seed=77)
    ntree, nodesize = 5, 5
    kw = dict(rule=case["rule"], ntree=ntree, nodesize=nodesize, seed=-4242, membership=True, quants=bool(case["kw"].get("quants")))
    if case["kw"].get("byuser"):
        kw["samp"] = np.random.default_rng(3).multinomial(d.n, np.full(d.n, 1.0 / d.n), size=ntree).astype(np.int32)
    ref = refdriver.grow(d.x, d.y, d.rt, d.strat, impl="ref", **kw)
    got = refdriver.grow(d.x, d.y, d.rt, d.strat, impl="b200", **kw)
    # same names in the same order (what R indexes by name and, around "treeID", by position: R/rfsrc.R:1766)
    assert [k for k in got if not k.startswith("_")] == [k for k in ref if not k.startswith("_")]
    # per-leaf arrays come at the reference's stride, so ONE trimming routine (R/rfsrc.R:1771-1861) serves both
    assert got["tnRCNT"].size == ref["tnRCNT"].size and got["tnACNT"].size == ref["tnACNT"].size
    tg, tr = _sorted_by_tree(refdriver.trim_forest(got, d.n)), _sorted_by_tree(refdriver.trim_forest(ref, d.n))
    parity.assert_forest_equal(tg, tr, d.n, ntree, check_membr_lists=True, label="rshim")
    np.testing.assert_array_equal(got["mwcpCT"], ref["mwcpCT"])
    assert got["mwcpPT"].size == ref["mwcpPT"].size == 0
    if kw["quants"]:
        np.testing.assert_array_equal(tg["tnCLAS"], tr["tnCLAS"])

    # predict and restore through the 58-argument entry, forest = what R would keep from EITHER grow call
    fx = make_cpiu(700, 9, levels=case["levels"], seed=78).x
    for forest in (tr, tg):
        for mode_fx in (fx, None):
            pkw = dict(nodesize=nodesize, samp=kw.get("samp"), quants=kw["quants"])
            pr = refdriver.predict(d.x, d.y, forest, mode_fx, impl="ref", **pkw)
            pg = refdriver.predict(d.x, d.y, forest, mode_fx, impl="b200", **pkw)
            assert list(pg) == list(pr)
            for k in pr:
                if pr[k].dtype.kind == "i":
                    np.testing.assert_array_equal(pg[k], pr[k], err_msg=k)
                else:
                    np.testing.assert_array_equal(np.isnan(pg[k]), np.isnan(pr[k]), err_msg=k)
                    ok = ~np.isnan(pr[k])
                    np.testing.assert_allclose(pg[k][ok], pr[k][ok], rtol=parity.HAZARD_RTOL, atol=0, err_msg=k)


@pytest.mark.gpu
@needs_shim
@needs_ref
def test_both_libraries_agree_on_a_mixed_factor_table():
    """xType "C" covariates through the 44 / 58-argument entries: parmID / contPT = NA / mwcpSZ / mwcpPT[1:sum(mwcpCT)] /
    mwcpCT as R reads them (R/rfsrc.R:1772-1800), and predictions of either library's forest by either library."""
    d = make_cpiu(3000, 8, levels=16, seed=5)
    x = d.x.copy(); rng = np.random.default_rng(9)
    xt = ["R"] * 8; xl = np.zeros(8, np.int32)
    for c, lv in {1: 5, 4: 8, 6: 3}.items():
        x[c] = rng.integers(1, lv + 1, size=d.n).astype(np.float64); xt[c] = "C"; xl[c] = lv
    ntree = 4
    kw = dict(rule=4, ntree=ntree, nodesize=10, seed=-77, membership=True, xtype=xt, xlevels=xl)
    ref = refdriver.grow(x, d.y, d.rt, d.strat, impl="ref", **kw)
    got = refdriver.grow(x, d.y, d.rt, d.strat, impl="b200", **kw)
    assert [k for k in got if not k.startswith("_")] == [k for k in ref if not k.startswith("_")]

    def canon(t):
        t = refdriver.trim_forest(t, d.n)
        order = np.argsort(t["treeID"], kind="stable")
        w = np.zeros(t["treeID"].size, np.int64); w[t["mwcpSZ"] > 0] = t["mwcpPT"][: int((t["mwcpSZ"] > 0).sum())]
        t["mwcpPT"] = w[order][t["mwcpSZ"][order] > 0].astype(np.int32)
        for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ"):
            t[k] = t[k][order]
        return t
    tg, tr = canon(got), canon(ref)
    assert int(tr["mwcpSZ"].sum()) > 20
    parity.assert_forest_equal(tg, tr, d.n, ntree, label="rshim factors")
    fx = x[:, ::3].copy()
    for forest in (tr, tg):
        pr = refdriver.predict(x, d.y, forest, fx, impl="ref", nodesize=10, xtype=xt, xlevels=xl)
        pg = refdriver.predict(x, d.y, forest, fx, impl="b200", nodesize=10, xtype=xt, xlevels=xl)
        np.testing.assert_array_equal(pg["nodeMembership"], pr["nodeMembership"])
        np.testing.assert_allclose(pg["allEnsbCLS"], pr["allEnsbCLS"], rtol=parity.HAZARD_RTOL, atol=0)


@pytest.mark.gpu
@needs_shim
@needs_ref
@pytest.mark.parametrize("imp", ["permute", "permute.ensemble"])
def test_both_libraries_return_the_same_importance(imp):
    """importance = TRUE through the .Call entries: "vimpClas" [xSize][3] from rfsrcGrow and, for a covariate subset, from
    rfsrcPredict in restore mode (what vimp.rfsrc calls), in the reference's place in the named list.  One reference thread
    (its per-tree importance shares buffers between trees)."""
    d = make_cpiu(2500, 7, levels=16, seed=6)
    kw = dict(rule=4, ntree=5, nodesize=10, seed=-31, membership=True, importance=imp, nthreads=1, quants=True)
    ref = refdriver.grow(d.x, d.y, d.rt, d.strat, impl="ref", **kw)
    got = refdriver.grow(d.x, d.y, d.rt, d.strat, impl="b200", **kw)
    assert [k for k in got if not k.startswith("_")] == [k for k in ref if not k.startswith("_")]
    assert got["vimpClas"].shape == ref["vimpClas"].shape == (21,)
    np.testing.assert_allclose(got["vimpClas"], ref["vimpClas"], rtol=0, atol=0 if imp == "permute" else 1e-12)
    forest = _sorted_by_tree(refdriver.trim_forest(ref, d.n))
    pkw = dict(nodesize=10, importance=imp, xvar_imp=[2, 7, 4], quants=True, nthreads=1)
    pr = refdriver.predict(d.x, d.y, forest, None, impl="ref", **pkw)
    pg = refdriver.predict(d.x, d.y, forest, None, impl="b200", **pkw)
    assert pg["vimpClas"].shape == pr["vimpClas"].shape == (9,)
    np.testing.assert_allclose(pg["vimpClas"], pr["vimpClas"], rtol=0, atol=0 if imp == "permute" else 1e-12)


@pytest.mark.gpu
@needs_shim
@needs_ref
@pytest.mark.parametrize("rfq", [False, True], ids=["vote", "rfq"])
@pytest.mark.parametrize("perf", ["misclass", "brier", "gmean"])
def test_both_libraries_score_held_out_rows_alike(perf, rfq):
    """rfsrcPredict on newdata with the outcome (frSize = 1, frData): the same "perfClas" row from both libraries; the RFQ
    vote takes its minority class and threshold from the TRAINING response (randomForestSRC.c:16859-16895)."""
    d = make_cpiu(4000, 7, levels=16, seed=8); t = make_cpiu(1500, 7, levels=16, seed=9)
    ref = refdriver.grow(d.x, d.y, d.rt, d.strat, impl="ref", rule=4, ntree=6, nodesize=10, seed=-3, quants=True)
    forest = _sorted_by_tree(refdriver.trim_forest(ref, d.n))
    pkw = dict(nodesize=10, quants=True, fy=t.y, perf=perf, rfq=rfq)
    pr = refdriver.predict(d.x, d.y, forest, t.x, impl="ref", **pkw)
    pg = refdriver.predict(d.x, d.y, forest, t.x, impl="b200", **pkw)
    assert [k for k in pg] == [k for k in pr]
    a, b = pg["perfClas"], pr["perfClas"]
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a[~np.isnan(b)], b[~np.isnan(b)], rtol=1e-12 if perf == "brier" else 0, atol=0)
    # and the permutation importance of three covariates over the held-out rows (one reference thread, see above)
    ikw = dict(pkw, importance="permute", xvar_imp=[3, 6, 1], nthreads=1)
    vr = refdriver.predict(d.x, d.y, forest, t.x, impl="ref", **ikw)["vimpClas"]
    vg = refdriver.predict(d.x, d.y, forest, t.x, impl="b200", **ikw)["vimpClas"]
    np.testing.assert_array_equal(np.isnan(vg), np.isnan(vr))
    np.testing.assert_allclose(vg[~np.isnan(vr)], vr[~np.isnan(vr)], rtol=0, atol=1e-12 if perf == "brier" else 0)
