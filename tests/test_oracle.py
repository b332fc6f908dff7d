"""CPU suite: pins the oracle restatement (oracle/rfslam_oracle.c) against the reference's own
outputs -- the committed golden fixtures (generated from the unmodified reference by
tests/golden/make_golden.py) and, when oracle/_ref/librfslam_ref.so is present, the reference
itself run live on fresh seeds."""
import os

import numpy as np
import pytest

from oracle import port, refdriver
from rfslam_b200.synth import make_cpiu
from tests import parity


def _golden_names():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return [os.path.basename(f) for f in parity.golden_files(here)]


@pytest.mark.parametrize("name", _golden_names())
def test_port_matches_golden_grow(golden_dir, name):
    g = parity.load_golden(os.path.join(golden_dir, name))
    q = parity.golden_params(g)
    got = port.grow(g["x"], g["y"], g["rt"], g["strat"], rule=q["rule"], ntree=q["ntree"], nodesize=q["nodesize"], nsplit=q["nsplit"],
                    mtry=q["mtry"], seed=q["seed"], samp=g.get("samp"))
    parity.assert_forest_equal(got, parity.golden_ref_forest(g), q["n"], q["ntree"], check_membr_lists=True, label=name)


@pytest.mark.parametrize("name", _golden_names())
def test_port_matches_golden_predict(golden_dir, name):
    g = parity.load_golden(os.path.join(golden_dir, name))
    ref = parity.golden_ref_forest(g)
    tn_e = np.zeros_like(ref["tnRCNT"])
    # leaf event counts are not a reference output; rebuild them from rmbr (in-bag rows by leaf)
    q = parity.golden_params(g)
    n, ntree = q["n"], q["ntree"]
    stride = ref["rmbrMembership"].size // ntree
    off = 0
    for b in range(ntree):
        rm = ref["rmbrMembership"][b * stride:(b + 1) * stride]
        t = parity.split_tree(ref, b + 1)
        pos = 0
        for leaf in t["nodeID"][t["parmID"] == 0]:          # rmbr is grouped by leaf in DFS order
            c = ref["tnRCNT"][off + leaf - 1]
            rows = rm[pos:pos + c] - 1
            tn_e[off + leaf - 1] = int(np.sum(g["y"][rows] == 2.0))
            pos += c
        off += ref["leafCount"][b]
    f = dict(ref)
    f["tnCLAS"] = np.stack([ref["tnRCNT"] - tn_e, tn_e], axis=1).reshape(-1)
    ens, mem = port.predict(f, g["fx"])
    np.testing.assert_array_equal(mem, g["pred_nodeMembership"])
    np.testing.assert_allclose(ens, g["pred_allEnsbCLS"], rtol=parity.HAZARD_RTOL)


def test_plugin_kat(golden_dir):
    z = np.load(os.path.join(golden_dir, "plugin_kat.npz"))
    for i in range(int(z["count"][0])):
        for rule in range(1, 7):
            got = port.poisson_stat(rule, float(z[f"kfa_{i}"][0]), z[f"memb_{i}"], z[f"resp_{i}"], z[f"strat_{i}"], z[f"rt_{i}"])
            want = z[f"stat_{i}"][rule - 1]
            assert got == pytest.approx(want, rel=1e-12, abs=1e-12), (i, rule, got, want)


def test_chain_seeds_match_reference_seed_output(golden_dir):
    g = parity.load_golden(os.path.join(golden_dir, "grow_r4_q64.npz"))
    q = parity.golden_params(g)
    a, _ = port.chain_seeds(q["seed"], q["ntree"])
    np.testing.assert_array_equal(a, g["ref_seed"])


def test_bootstrap_counts_match_reference(golden_dir):
    g = parity.load_golden(os.path.join(golden_dir, "grow_r4_q64.npz"))
    q = parity.golden_params(g)
    for b in range(q["ntree"]):
        d = port.bootstrap(q["seed"], q["ntree"], q["n"], b)
        cnt = np.bincount(d - 1, minlength=q["n"])
        np.testing.assert_array_equal(cnt, g["ref_bootMembership"][b * q["n"]:(b + 1) * q["n"]])


def test_ran1_range_and_float_semantics():
    s = port.ran1_stream(-4711, 5000)
    assert s.dtype == np.float32
    assert (s > 0).all() and (s <= np.float32(1.0 - 1.2e-7)).all()


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("rule,n,p,levels,nodesize,seed,dseed", [
    (4, 2200, 11, 32, 5, -31337, 501), (1, 1700, 8, 0, 10, -2, 502), (5, 1300, 6, 8, 1, -987654, 503),
    (3, 1900, 13, 64, 4, -800000, 504)])
def test_port_matches_live_reference(rule, n, p, levels, nodesize, seed, dseed):
    d = make_cpiu(n, p, levels=levels, seed=dseed)
    ntree = 5
    r = refdriver.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=seed)
    t = refdriver.trim_forest(r, n)
    order = np.argsort(t["treeID"], kind="stable")
    for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ"):
        t[k] = t[k][order]
    got = port.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=seed)
    parity.assert_forest_equal(got, t, n, ntree, check_membr_lists=True, label="live")


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("rule,n,p,levels,nsplit,nodesize,seed,dseed", [
    (4, 1500, 9, 64, 5, 5, -12345, 105), (1, 2000, 8, 32, 3, 5, -7, 103), (4, 1200, 6, 0, 10, 5, -31, 110),
    (6, 1800, 10, 16, 1, 5, -99, 101)])
def test_port_nsplit_matches_live_reference(rule, n, p, levels, nsplit, nodesize, seed, dseed):
    # random cut subsets (randomForestSRC.c:14533-14564), continuous covariates included
    d = make_cpiu(n, p, levels=levels, seed=dseed)
    ntree = 5
    r = refdriver.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=seed, nsplit=nsplit)
    t = refdriver.trim_forest(r, n)
    order = np.argsort(t["treeID"], kind="stable")
    for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ"):
        t[k] = t[k][order]
    got = port.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=seed, nsplit=nsplit)
    parity.assert_forest_equal(got, t, n, ntree, check_membr_lists=True, label=f"live nsplit={nsplit}")


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_variable_categories_do_not_change_the_draws():
    """Pins the claim behind accepting nCategories > 1 on the device path (randomForestSRC.c:14887-15014):
    under uniform xvar.wt the category loop never changes which covariates are drawn."""
    d = make_cpiu(1200, 9, levels=32, seed=71)
    base = refdriver.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=4, nodesize=5, seed=-19)
    cats = np.array([1, 2, 3, 1, 2, 3, 1, 2, 3], np.int32)
    for w in ([0.5, 0.25, 0.25], [1 / 3, 1 / 3, 1 / 3], [0.8, 0.1, 0.1]):
        other = refdriver.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=4, nodesize=5, seed=-19, var_categories=cats,
                               category_weights=np.array(w))
        a, b = refdriver.trim_forest(base, d.n), refdriver.trim_forest(other, d.n)
        ta, tb = refdriver.per_tree(a), refdriver.per_tree(b)
        for k in ta:
            np.testing.assert_array_equal(ta[k]["parmID"], tb[k]["parmID"])
            np.testing.assert_array_equal(ta[k]["nodeID"], tb[k]["nodeID"])
        np.testing.assert_array_equal(base["nodeMembership"], other["nodeMembership"])


# ---- arguments the first round left untested: each pinned against the live reference -----------------
def _wide_strata(d, K):
    """Re-stratify a synthetic table into K strata (the reference allocates per-stratum arrays of K + 1)."""
    rng = np.random.default_rng(5)
    return rng.integers(1, K + 1, d.n).astype(np.int32)


UNTESTED_ARG_CASES = [
    dict(tag="split_weight", rule=4, kw=dict(), split_weight=True),
    dict(tag="nodedepth3", rule=4, kw=dict(nodedepth=3)),
    dict(tag="nodedepth0", rule=1, kw=dict(nodedepth=0)),
    dict(tag="sampsize", rule=4, kw=dict(sampsize=900)),
    dict(tag="sampsize_bayes", rule=2, kw=dict(sampsize=1400)),
    dict(tag="strata40", rule=4, kw=dict(), strata=40),
    dict(tag="strata40_bayes", rule=1, kw=dict(), strata=40),
    dict(tag="byuser_r1", rule=1, kw=dict(), byuser=True),
    dict(tag="byuser_r2", rule=2, kw=dict(), byuser=True),
    dict(tag="byuser_r3", rule=3, kw=dict(), byuser=True),
]


def untested_arg_inputs(case, n=1600, p=9):
    d = make_cpiu(n, p, levels=32, seed=900 + len(case["tag"]))
    kw = dict(case["kw"])
    if case.get("strata"):
        d.strat = _wide_strata(d, case["strata"])
    if case.get("split_weight"):
        kw["split_weight"] = np.linspace(0.25, 2.0, p)
    if case.get("byuser"):
        rng = np.random.default_rng(17)
        samp = rng.multinomial(n, np.full(n, 1.0 / n), size=4).astype(np.int32)
        samp[:, 0] = 70                                           # one row drawn more often than an entry's multiplicity field holds (64)
        kw["samp"] = samp
    return d, kw


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("case", UNTESTED_ARG_CASES, ids=[c["tag"] for c in UNTESTED_ARG_CASES])
def test_port_matches_live_reference_other_arguments(case):
    d, kw = untested_arg_inputs(case)
    t = parity.live_reference_forest(d, rule=case["rule"], ntree=4, nodesize=5, seed=-77, **kw)
    got = port.grow(d.x, d.y, d.rt, d.strat, rule=case["rule"], ntree=4, nodesize=5, seed=-77, **kw)
    parity.assert_forest_equal(got, t, d.n, 4, check_membr_lists=not case.get("byuser") and "sampsize" not in case["kw"], label=case["tag"])


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_port_split_statistic_matches_reference_node_statistics():
    """The winning statistic per node (what the tie-aware comparison of tests/parity.py rests on):
    the port's splitStat against the reference's own mtryST of the chosen covariate (OPT_NODE_STAT)."""
    d = make_cpiu(4000, 10, levels=64, seed=31)
    for rule in (1, 4):
        t = parity.live_reference_forest(d, statistics=True, rule=rule, ntree=3, nodesize=5, seed=-7)
        w = port.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=3, nodesize=5, seed=-7)
        parity.assert_forest_equal(w, t, d.n, 3, label="stat")
        m = w["parmID"] != 0
        np.testing.assert_allclose(w["splitStat"][m], t["splitStat"][m], rtol=1e-9, atol=1e-9)


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("rule", [1, 4])
def test_port_matches_live_reference_zero_risk_time(rule):
    """Degenerate exposure: every risk time 0.  The Bayes rules then produce inf - inf = NaN deltas; the
    reference's tracker takes the FIRST delta of a node whatever it is (deltaMax starts as NA,
    randomForestSRC.c:15497-15499) and nothing ever beats a NaN."""
    d = make_cpiu(900, 5, levels=16, seed=41)
    d.rt = np.zeros(d.n)
    t = parity.live_reference_forest(d, rule=rule, ntree=3, nodesize=5, seed=-3)
    got = port.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=3, nodesize=5, seed=-3)
    parity.assert_forest_equal(got, t, d.n, 3, label="zero-rt")


def test_tie_aware_comparison_accepts_ties_only():
    d = make_cpiu(1500, 6, levels=16, seed=51)
    a = port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=3, nodesize=5, seed=-5)
    b = {k: v.copy() for k, v in a.items()}
    assert parity.assert_forest_equal(a, b, d.n, 3, tie_aware=True) == {"ties": 0, "trees_compared": 3}
    # move the cutpoint of the second tree's root: a tie when the statistics agree, a failure when they do not
    k = int(np.nonzero(b["treeID"] == 2)[0][0])
    b["contPT"][k] = b["contPT"][k] + 1.0
    res = parity.assert_forest_equal(a, b, d.n, 3, tie_aware=True)
    assert res == {"ties": 1, "trees_compared": 2}
    with pytest.raises(AssertionError):
        parity.assert_forest_equal(a, b, d.n, 3)
    b["splitStat"][k] = b["splitStat"][k] * (1.0 + 1e-6)
    with pytest.raises(AssertionError, match="not a tie"):
        parity.assert_forest_equal(a, b, d.n, 3, tie_aware=True)


# ---- unordered factors (MWCP splits, randomForestSRC.c:14452-14518) --------------------------------------------------
FACTOR_GOLDENS = ["fac_r4_det.npz", "fac_r1_rand.npz", "fac_r6_nsplit.npz", "fac_r4_big.npz"]


@pytest.mark.parametrize("name", FACTOR_GOLDENS)
def test_port_factor_splits_match_golden(golden_dir, name):
    """Deterministic and random complementary pairs, nsplit > 0 on factors: forests (incl. mwcpSZ / mwcpPT / mwcpCT) and the
    held-out predictions of the reference's own rfsrcPredict."""
    g = parity.load_golden(os.path.join(golden_dir, name))
    q = parity.golden_params(g)
    xt = list(str(g["xtype"]))
    got = port.grow(g["x"], g["y"], g["rt"], g["strat"], rule=q["rule"], ntree=q["ntree"], nodesize=q["nodesize"], nsplit=q["nsplit"],
                    mtry=q["mtry"], seed=q["seed"], xtype=xt, xlevels=g["xlevels"])
    ref = parity.golden_ref_forest(g)
    assert int(ref["mwcpSZ"].sum()) > 50
    parity.assert_forest_equal(got, ref, q["n"], q["ntree"], label=name)
    ens, memb = port.predict(got, g["fx"])
    np.testing.assert_array_equal(memb, g["pred_nodeMembership"])
    np.testing.assert_allclose(ens, g["pred_allEnsbCLS"], rtol=parity.HAZARD_RTOL, atol=0)


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_factor_pair_enumeration_matches_reference_tables():
    """bookFactor's complementary pairs (randomForestSRC.c:1459-1530) for r = 2..12 levels: by LEFT cardinality, lexicographic
    inside a cardinality, the middle cardinality of an even r halved -- the order the restatement and the kernel unrank."""
    import ctypes as C
    import itertools

    class Factor(C.Structure):
        _fields_ = [("r", C.c_uint), ("cgc", C.c_uint), ("cpc", C.c_void_p), ("cgs", C.c_void_p),
                    ("cgb", C.POINTER(C.POINTER(C.c_uint))), ("mwcp", C.c_uint)]
    L = C.CDLL(refdriver.REF_SO)
    L.makeFactor.restype = C.POINTER(Factor); L.makeFactor.argtypes = [C.c_uint, C.c_char]
    for r in range(2, 13):
        f = L.makeFactor(r, b"\x01").contents
        sizes = C.cast(f.cgs, C.POINTER(C.c_uint))
        for gsz in range(1, f.cgc + 1):
            ref = [f.cgb[gsz][k] for k in range(1, sizes[gsz] + 1)]
            combos = list(itertools.combinations(range(1, r + 1), gsz))
            if r % 2 == 0 and gsz == r // 2:
                combos = combos[: len(combos) // 2]
            assert ref == [sum(1 << (l - 1) for l in c) for c in combos], (r, gsz)


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("n,p,fac,rule,nodesize,seed,nsplit,mtry", [
    (3000, 6, {1: 5, 4: 8}, 4, 5, -7, 0, None), (2000, 5, {0: 3, 2: 12}, 4, 3, -11, 0, None),
    (3000, 6, {1: 6, 4: 20}, 2, 5, -21, 10, None), (1500, 4, {0: 32, 1: 7}, 5, 2, -5, 0, 4)])
def test_port_factor_splits_match_live_reference(n, p, fac, rule, nodesize, seed, nsplit, mtry):
    d = make_cpiu(n, p, levels=16, seed=5)
    x = d.x.copy(); rng = np.random.default_rng(1)
    xt = ["R"] * p; xl = np.zeros(p, np.int32)
    for c, lv in fac.items():
        x[c] = rng.integers(1, lv + 1, size=d.n).astype(float); xt[c] = "C"; xl[c] = lv
    kw = dict(rule=rule, ntree=3, nodesize=nodesize, seed=seed, nsplit=nsplit, mtry=mtry, xtype=xt, xlevels=xl)
    import dataclasses
    ref = parity.live_reference_forest(dataclasses.replace(d, x=x), **kw)
    got = port.grow(x, d.y, d.rt, d.strat, **kw)
    parity.assert_forest_equal(got, ref, d.n, 3, label="factor live")


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("rfq", [False, True])
@pytest.mark.parametrize("perf", ["misclass", "brier", "gmean"])
def test_port_perf_types_match_live_reference(perf, rfq):
    """perf.type = misclass / brier / g.mean and the RFQ vote (randomForestSRC.c:974-1218, :7954-8010): bit-exact."""
    d = make_cpiu(3000, 6, levels=16, seed=5)
    kw = dict(rule=4, ntree=5, nodesize=5, seed=-7, perf=perf, rfq=rfq)
    r = refdriver.grow(d.x, d.y, d.rt, d.strat, **kw)
    o = port.grow(d.x, d.y, d.rt, d.strat, **kw)
    np.testing.assert_array_equal(r["perfClas"].view(np.uint64), o["perfClas"].view(np.uint64))


# ---- permutation importance (randomForestSRC.c:2106-2218, 2416-2697) -----------------------------------------
VIMP_GOLDENS = ["vimp_q16.npz", "vimp_fac.npz", "vimp_byuser.npz"]


def vimp_golden_forest(g):
    f = {k[4:]: g[k] for k in g.files if k.startswith("ref_") and not k.startswith(("ref_vimp", "ref_perf", "ref_restore"))}
    return f


def vimp_modes(g):
    for k in g.files:
        if k.startswith("ref_vimp_"):
            imp, perf, rfq = k[len("ref_vimp_"):].rsplit("_", 2)
            yield k, imp, perf, bool(int(rfq))


def assert_vimp_close(got, want, exact):
    got = np.asarray(got).reshape(-1, 3); want = np.asarray(want).reshape(-1, 3)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    if exact:
        np.testing.assert_array_equal(got[ok], want[ok])
    else:
        np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", VIMP_GOLDENS)
def test_port_permutation_importance_matches_golden(golden_dir, name):
    """The restatement of permute() / getPermuteMembership / summarize + finalizeVimpPerformance against the unmodified
    reference: per-tree and ensemble importance, every perf.type, RFQ, numeric + factor tables, by.user bootstraps, in
    grow mode (chain C = third seed block) and restore mode (seed LCG from 0, a covariate subset)."""
    g = np.load(os.path.join(golden_dir, name))
    n, p, levels, rule, ntree, nodesize, seed, dseed, byuser = [int(v) for v in g["params"]]
    f = vimp_golden_forest(g)
    boot = g["ref_bootMembership"]
    for key, imp, perf, rfq in vimp_modes(g):
        got = port.vimp_permute(f, g["x"], g["y"], boot, port.chain_seed_c(seed, ntree), per_tree=(imp == "permute"), perf=perf, rfq=rfq)
        assert_vimp_close(got, g[key], exact=(imp == "permute" and perf != "brier"))
    for imp in ("permute", "permute.ensemble"):
        got = port.vimp_permute(f, g["x"], g["y"], boot, port.chain_seed_c(0, ntree, restore=True), intr=g["intr"], per_tree=(imp == "permute"))
        assert_vimp_close(got, g["ref_restore_vimp_" + imp], exact=(imp == "permute"))


def test_permutation_is_the_reference_one():
    """permute() (randomForestSRC.c:1974-2007): slot 'k-th still empty' receives i for i = m..1 -- every value once."""
    f = dict(treeID=np.ones(1, np.int32), nodeID=np.ones(1, np.int32), parmID=np.zeros(1, np.int32), contPT=np.zeros(1),
             leafCount=np.ones(1, np.int32), tnCLAS=np.array([3, 1], np.int32))
    x = np.arange(10, dtype=np.float64)[None, :]; y = np.ones(10); y[::3] = 2.0
    v = port.vimp_permute(f, x, y, np.zeros((1, 10), np.int32), port.chain_seed_c(-3, 1), per_tree=True)
    assert v.shape == (1, 3) and v[0, 0] == 0.0                       # a stump: permuting changes nothing


@pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("imp,perf,rfq", [("permute", "misclass", False), ("permute.ensemble", "brier", False), ("permute", "gmean", True)])
def test_port_importance_matches_live_reference(imp, perf, rfq):
    """Fresh seeds, three run modes: inside the grow call, restore mode (covariate subset) and held-out rows with responses.
    One reference thread: its per-tree importance shares buffers between trees, and rfsrcPredict hard-codes two threads."""
    d = make_cpiu(2500, 6, levels=16, seed=41); t = make_cpiu(800, 6, levels=16, seed=42)
    ntree, seed = 5, -57
    raw = refdriver.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=ntree, nodesize=8, seed=seed, importance=imp, perf=perf, rfq=rfq,
                         quants=True, nthreads=1)
    f = refdriver.trim_forest(raw, d.n)
    order = np.argsort(f["treeID"], kind="stable")
    for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ"):
        f[k] = f[k][order]
    per_tree = imp == "permute"
    exact = per_tree and perf != "brier"
    got = port.vimp_permute(f, d.x, d.y, raw["bootMembership"], port.chain_seed_c(seed, ntree), per_tree=per_tree, perf=perf, rfq=rfq)
    assert_vimp_close(got, raw["vimpClas"], exact=exact)
    intr = [5, 2]
    pkw = dict(nodesize=8, importance=imp, xvar_imp=intr, quants=True, nthreads=1, perf=perf, rfq=rfq)
    rest = refdriver.predict(d.x, d.y, f, None, **pkw)["vimpClas"]
    got = port.vimp_permute(f, d.x, d.y, raw["bootMembership"], port.chain_seed_c(0, ntree, restore=True), intr=intr, per_tree=per_tree, perf=perf, rfq=rfq)
    assert_vimp_close(got, rest, exact=exact)
    if not rfq:                                                       # (the restatement takes the RFQ threshold from the rows it scores)
        held = refdriver.predict(d.x, d.y, f, t.x, fy=t.y, **pkw)["vimpClas"]
        got = port.vimp_permute(f, t.x, t.y, np.zeros((ntree, t.n), np.int32), port.chain_seed_c(0, ntree, restore=True), intr=intr,
                                per_tree=per_tree, perf=perf)
        assert_vimp_close(got, held, exact=exact)
