"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C-ABI,
against the committed golden outputs of the unmodified reference and against the oracle port on
fresh seeded inputs.  Bars (north_star): in-bag counts / node membership / topology bit-exact;
split variable and cutpoint identical; hazards (ensembles) within 1e-6 relative."""
import os

import numpy as np
import pytest

import rfslam_b200
from oracle import port
from rfslam_b200.synth import make_cpiu
from tests import parity

pytestmark = pytest.mark.gpu


def _golden_names():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return [os.path.basename(f) for f in parity.golden_files(here)]


def test_library_sees_a_gpu():
    from rfslam_b200 import capi
    assert capi.lib().rfslam_b200_device_count() >= 1


def test_bootstrap_counts_bit_exact(golden_dir):
    g = parity.load_golden(os.path.join(golden_dir, "grow_r4_q64.npz"))
    q = parity.golden_params(g)
    for b in range(q["ntree"]):
        got = rfslam_b200.bootstrap_counts(q["seed"], q["ntree"], q["n"], b + 1)
        np.testing.assert_array_equal(got, g["ref_bootMembership"][b * q["n"]:(b + 1) * q["n"]])


def test_score_node_matches_plugin_kat(golden_dir):
    """One-cut nodes built from the reference plugin's own known answers (x = 0 on the LEFT rows)."""
    z = np.load(os.path.join(golden_dir, "plugin_kat.npz"))
    for i in range(int(z["count"][0])):
        memb = z[f"memb_{i}"]
        x = np.where(memb == 1, 0.0, 1.0)
        for rule in range(1, 7):
            cut, delta = rfslam_b200.score_node(rule, float(z[f"kfa_{i}"][0]), x, z[f"resp_{i}"],
                                                z[f"strat_{i}"].astype(np.int32), z[f"rt_{i}"])
            assert cut.size == 1 and cut[0] == 0.0
            want = z[f"stat_{i}"][rule - 1]
            assert delta[0] == pytest.approx(want, rel=1e-9, abs=1e-9), (i, rule, delta[0], want)


@pytest.mark.parametrize("levels", [8, 200, 0])
@pytest.mark.parametrize("rule", [1, 2, 3, 4, 5, 6])
def test_score_node_every_cut_matches_oracle(rule, levels):
    """Histogram path (<= 256 distinct values) and sorted-walk path (continuous) against one oracle
    plugin evaluation per cut."""
    d = make_cpiu(700, 3, levels=levels, seed=77 + rule)
    x = d.x[2] if levels else d.x[0] + 1e-3 * np.arange(d.n)
    cut, delta = rfslam_b200.score_node(rule, 2.0, x, d.y, d.strat, d.rt)
    vals = np.unique(x)
    np.testing.assert_array_equal(cut, vals[:-1])
    for k in range(0, cut.size, max(1, cut.size // 60)):
        memb = (x <= cut[k]).astype(np.int8)
        want = port.poisson_stat(rule, 2.0, memb, d.y, d.strat.astype(np.float64), d.rt)
        assert delta[k] == pytest.approx(want, rel=1e-9, abs=1e-9), (rule, levels, k)


@pytest.mark.parametrize("name", _golden_names())
def test_grow_matches_reference_golden(golden_dir, name):
    g = parity.load_golden(os.path.join(golden_dir, name))
    q = parity.golden_params(g)
    got = rfslam_b200.rfsrc(g["x"], g["y"], g["rt"], g["strat"], splitrule=q["rule"], ntree=q["ntree"],
                            nodesize=q["nodesize"], mtry=q["mtry"], seed=q["seed"], samp=g.get("samp"), membership=True,
                            member_lists=True, nsplit=q["nsplit"])
    parity.assert_forest_equal(got, parity.golden_ref_forest(g), q["n"], q["ntree"], check_membr_lists=True, label=name)


@pytest.mark.parametrize("name", ["grow_r4_q64.npz", "grow_r1_cont.npz", "grow_r4_byuser.npz"])
def test_predict_matches_reference_golden(golden_dir, name):
    g = parity.load_golden(os.path.join(golden_dir, name))
    q = parity.golden_params(g)
    f = rfslam_b200.rfsrc(g["x"], g["y"], g["rt"], g["strat"], splitrule=q["rule"], ntree=q["ntree"],
                          nodesize=q["nodesize"], mtry=q["mtry"], seed=q["seed"], samp=g.get("samp"))
    pr = rfslam_b200.predict_rfsrc(f, g["fx"])
    np.testing.assert_array_equal(pr["nodeMembership"], g["pred_nodeMembership"])
    np.testing.assert_allclose(pr["allEnsbCLS"], g["pred_allEnsbCLS"], rtol=parity.HAZARD_RTOL)


def test_restore_and_sharded_predict():
    """Restore mode (fx = None: the forest on its own training rows, rf.c:22361) reproduces the grow call's ensembles
    and perfClas; leaf class counts rebuilt from forest$seed equal the tnCLAS the grow call returned; held-out rows
    sharded over processes (forest replicated, no collective) concatenate to the unsharded result."""
    d = make_cpiu(5000, 10, levels=64, seed=91)
    f = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=7, nodesize=5, seed=-17, membership=True)
    r = rfslam_b200.predict_rfsrc(f, None, x=d.x, y=d.y, use_tnclas=False)
    for k in ("oobEnsbCLS", "allEnsbCLS"):
        np.testing.assert_allclose(r[k], f[k], rtol=1e-12, atol=0, err_msg=k)
    np.testing.assert_array_equal(r["perfClas"].view(np.uint64), f["perfClas"].view(np.uint64))
    np.testing.assert_array_equal(r["nodeMembership"], f["nodeMembership"])
    np.testing.assert_array_equal(r["bootMembership"], f["bootMembership"])
    np.testing.assert_array_equal(r["leafCount"], f["leafCount"])
    fx = make_cpiu(3001, 10, levels=64, seed=92).x
    a = rfslam_b200.predict_rfsrc(f, fx)                                         # tnCLAS passed
    b = rfslam_b200.predict_rfsrc(f, fx, x=d.x, y=d.y, use_tnclas=False)          # rebuilt from the training rows
    np.testing.assert_array_equal(a["allEnsbCLS"], b["allEnsbCLS"])
    np.testing.assert_array_equal(a["nodeMembership"], b["nodeMembership"])
    ens, memb = port.predict(f, fx)
    np.testing.assert_array_equal(a["nodeMembership"], memb)
    np.testing.assert_allclose(a["allEnsbCLS"], ens, rtol=parity.HAZARD_RTOL)
    parts = [rfslam_b200.predict_rfsrc(f, fx, frow_first=lo, frow_count=cnt, membership=False) for lo, cnt in ((0, 1000), (1000, 1500), (2500, 501))]
    cat = np.concatenate([q["allEnsbCLS"].reshape(2, -1) for q in parts], axis=1).reshape(-1)
    np.testing.assert_array_equal(cat, a["allEnsbCLS"])
    assert a["node_visits"] > 0 and a["row_tree_visits"] == 7 * 3001
    # small tiles: several uploads per sweep, both buffers in use
    os.environ["RFSLAM_PREDICT_TILE_ROWS"] = "512"
    try:
        c = rfslam_b200.predict_rfsrc(f, fx)
    finally:
        del os.environ["RFSLAM_PREDICT_TILE_ROWS"]
    np.testing.assert_array_equal(c["allEnsbCLS"], a["allEnsbCLS"])
    np.testing.assert_array_equal(c["nodeMembership"], a["nodeMembership"])


def test_plugin_symbols_match_reference_known_answers(golden_dir):
    """poissonSplit1..6 with the reference's own signature (splitCustom.h:24-44), 1-based arrays, against the
    statistics the reference's plugin returned for the same memberships (tests/golden/plugin_kat.npz)."""
    import ctypes as C
    from rfslam_b200 import capi
    L = capi.lib()
    z = np.load(os.path.join(golden_dir, "plugin_kat.npz"))
    for i in range(int(z["count"][0])):
        memb = np.concatenate([[0], z[f"memb_{i}"]]).astype(np.int8)
        resp = np.concatenate([[0.0], z[f"resp_{i}"]]); strat = np.concatenate([[0.0], z[f"strat_{i}"].astype(np.float64)])
        rt = np.concatenate([[0.0], z[f"rt_{i}"]])
        feat = (C.c_void_p * 3)(None, strat.ctypes.data, rt.ctypes.data)
        for rule in range(1, 7):
            got = getattr(L, f"poissonSplit{rule}")(memb.size - 1, float(z[f"kfa_{i}"][0]), memb.ctypes.data, None, None, 0, 0, None,
                                                     resp.ctypes.data, 0.0, 0.0, 2, C.addressof(feat), 2)
            want = z[f"stat_{i}"][rule - 1]
            assert got == pytest.approx(want, rel=1e-9, abs=1e-9), (i, rule, got, want)


@pytest.mark.parametrize("rule,n,p,levels,nodesize,seed,dseed", [
    (4, 6000, 14, 64, 5, -31337, 601), (1, 5000, 10, 0, 10, -2, 602), (5, 3000, 6, 8, 1, -987654, 603),
    (3, 4000, 13, 256, 4, -800000, 604), (2, 3500, 9, 0, 20, -55, 605), (6, 4500, 20, 32, 7, -66, 606)])
def test_grow_matches_oracle_fresh_seeds(rule, n, p, levels, nodesize, seed, dseed):
    d = make_cpiu(n, p, levels=levels, seed=dseed)
    ntree = 6
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=seed)
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=rule, ntree=ntree, nodesize=nodesize, seed=seed, membership=True)
    parity.assert_forest_equal(got, want, n, ntree, label="fresh")
    assert got["split_candidates"] == int(want["candidates"][0])


def test_empty_and_degenerate_nodes():
    """No events at all -> every tree is a stump; a constant covariate is never chosen."""
    d = make_cpiu(600, 4, levels=16, seed=9)
    y = np.ones(d.n)
    got = rfslam_b200.rfsrc(d.x, y, d.rt, d.strat, ntree=3, nodesize=2)
    np.testing.assert_array_equal(got["leafCount"], [1, 1, 1])
    x = d.x.copy(); x[1] = 3.25
    want = port.grow(x, d.y, d.rt, d.strat, rule=4, ntree=4, nodesize=3, seed=-9)
    got = rfslam_b200.rfsrc(x, d.y, d.rt, d.strat, splitrule=4, ntree=4, nodesize=3, seed=-9, membership=True)
    parity.assert_forest_equal(got, want, d.n, 4, label="constant-column")
    assert not np.any(got["parmID"] == 2)


def test_tree_sharding_is_invariant():
    """Tree b is identical whichever shard grows it, and partial ensembles add up."""
    d = make_cpiu(3000, 8, levels=32, seed=10)
    full = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=6, seed=-4, finalize=False)
    parts = [rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=6, seed=-4, tree_first=r + 1, tree_step=2, finalize=False) for r in range(2)]
    for r, pt in enumerate(parts):
        np.testing.assert_array_equal(pt["tree_ids"], np.arange(r + 1, 7, 2))
        for b in pt["tree_ids"]:
            for k in ("nodeID", "parmID"):
                np.testing.assert_array_equal(parity.split_tree(pt, b)[k], parity.split_tree(full, b)[k])
    np.testing.assert_allclose(parts[0]["allEnsbCLS"] + parts[1]["allEnsbCLS"], full["allEnsbCLS"], rtol=1e-12)
    np.testing.assert_array_equal(parts[0]["oobEnsbDen"] + parts[1]["oobEnsbDen"], full["oobEnsbDen"])


def test_resident_dataset_and_determinism():
    d = make_cpiu(20000, 12, levels=64, seed=11)
    ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
    a = ds.grow(ntree=8, nodesize=20, seed=-3)
    b = ds.grow(ntree=8, nodesize=20, seed=-3)
    for k in ("treeID", "nodeID", "parmID", "contPT", "tnRCNT", "tnACNT"):
        np.testing.assert_array_equal(a[k], b[k])
    c = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=8, nodesize=20, seed=-3)
    np.testing.assert_array_equal(a["parmID"], c["parmID"])
    # structural invariants that hold at any size
    assert a["total_device_ms"] > 0 and a["split_candidates"] > 0
    for t in range(8):
        tr = parity.split_tree(a, t + 1)
        assert tr["nodeID"].size == 2 * a["leafCount"][t] - 1
    assert a["tnRCNT"].sum() == 8 * d.n           # every in-bag draw lands in exactly one leaf
    assert a["tnACNT"].sum() == 8 * d.n           # every row reaches exactly one leaf per tree
    ds.close()


def test_unsupported_inputs_fail_loudly():
    d = make_cpiu(300, 3, levels=8, seed=12)
    with pytest.raises(rfslam_b200.RfslamError):
        rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=2, seed=5)            # seed must be < 0
    with pytest.raises(rfslam_b200.RfslamError):
        rfslam_b200.rfsrc(d.x, d.y + 1.0, d.rt, d.strat, ntree=2)              # response codes outside {1, 2}
    dc = make_cpiu(2000, 3, levels=0, seed=12)                                   # continuous: > 256 distinct values
    with pytest.raises(rfslam_b200.RfslamError):
        rfslam_b200.rfsrc(dc.x, dc.y, dc.rt, dc.strat, ntree=2, nsplit=300)    # more than 256 drawn cuts of such a covariate: not kept


def test_variable_categories_accepted():
    """rfSLAM's structured-covariate arguments (R/rfsrc.R:1529-1585): inert for the draws (see
    tests/test_oracle.py::test_reference_variable_categories_do_not_change_the_draws), so the forest is unchanged."""
    d = make_cpiu(1500, 9, levels=32, seed=72)
    base = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=3, nodesize=5, seed=-21)
    cats = np.array([1, 2, 3, 1, 2, 3, 1, 2, 3], np.int32)
    other = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=3, nodesize=5, seed=-21, var_categories=cats,
                              category_weights=np.array([0.5, 0.25, 0.25]))
    np.testing.assert_array_equal(base["parmID"], other["parmID"])
    with pytest.raises(rfslam_b200.RfslamError):
        rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=3, nodesize=5, seed=-21, mtry=9, var_categories=cats,
                          category_weights=np.array([0.1, 0.1, 0.1]))


def test_time_slicing_does_not_change_the_forest(monkeypatch):
    # more trees than resident CTAs: trees are parked and resumed by other CTAs (grow_kernel.cuh);
    # the forest must be the one grown with one CTA per tree from start to finish, and match the oracle
    d = make_cpiu(3000, 8, levels=32, seed=21)
    ntree = 700
    monkeypatch.setenv("RFSLAM_QUANTUM_MS", "0.05")              # many switches per tree
    a = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=ntree, nodesize=5, seed=-9, membership=True)
    monkeypatch.setenv("RFSLAM_NO_SLICING", "1")
    b = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=ntree, nodesize=5, seed=-9, membership=True)
    for k in ("treeID", "nodeID", "parmID", "leafCount", "tnRCNT", "tnACNT", "nodeMembership", "bootMembership"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    np.testing.assert_array_equal(a["contPT"].view(np.uint64), b["contPT"].view(np.uint64))
    np.testing.assert_array_equal(a["oobEnsbCLS"], b["oobEnsbCLS"])
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=ntree, nodesize=5, seed=-9)
    parity.assert_forest_equal(a, want, d.n, ntree, label="sliced")


def test_full_size_properties():
    # BASELINE configs[1] rows x covariates (1M x 50, 64 levels), more trees than resident CTAs:
    # size-independent properties -- every draw and every row lands in exactly one leaf, records are
    # a full binary tree in pre-order, the time-sliced schedule and tree sharding do not change a tree
    d = make_cpiu(1_000_000, 50, levels=64)
    ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat)
    ntree = 304
    a = ds.grow(ntree=ntree, nodesize=20, seed=-12345)
    assert a["tnRCNT"].sum() == ntree * d.n and a["tnACNT"].sum() == ntree * d.n
    assert a["treeID"].size == int((2 * a["leafCount"].astype(np.int64) - 1).sum())
    assert int((a["parmID"] == 0).sum()) == int(a["leafCount"].sum())
    assert np.all(a["tnRCNT"] >= 1) and np.all(a["oobEnsbDen"] <= ntree)
    ens = a["allEnsbCLS"].reshape(2, -1)
    np.testing.assert_allclose(ens.sum(axis=0), 1.0, rtol=1e-12)             # class proportions of a row sum to one
    assert 0.0 <= a["perfClas"][-3] <= 1.0 and np.isnan(a["perfClas"][:-3]).all()
    # the same trees from a shard that holds every 38th tree (8 trees: one CTA each, no slicing)
    part = ds.grow(ntree=ntree, nodesize=20, seed=-12345, tree_first=5, tree_step=38, finalize=False)
    for j, b in enumerate(range(5, ntree + 1, 38)):
        full_t, part_t = parity.split_tree(a, b), parity.split_tree(part, b)
        for k in ("nodeID", "parmID"):
            np.testing.assert_array_equal(full_t[k], part_t[k], err_msg=f"tree {b} {k}")
        np.testing.assert_array_equal(full_t["contPT"].view(np.uint64), part_t["contPT"].view(np.uint64))
    ds.close()


@pytest.mark.parametrize("rule,n,p,levels,nsplit,nodesize,seed,dseed", [
    (4, 4000, 9, 64, 5, 5, -12345, 701), (1, 3000, 8, 32, 3, 3, -7, 702), (6, 2500, 12, 256, 10, 5, -99, 703),
    (2, 3000, 6, 16, 1, 2, -31, 704), (4, 20000, 10, 64, 2, 20, -5, 705),
    # covariates with more than 256 distinct values (continuous, levels = 0): the node's codes are sorted to count them
    (4, 5000, 8, 0, 4, 5, -41, 706), (1, 3000, 6, 0, 1, 3, -43, 707), (5, 12000, 9, 0, 10, 20, -47, 708), (6, 2000, 5, 0, 200, 2, -49, 709)])
def test_nsplit_random_cut_subsets_match_oracle(rule, n, p, levels, nsplit, nodesize, seed, dseed):
    # nsplit > 0 (rf.c:14533-14564): cut draws interleave with covariate draws in chain B and depend on the
    # number of distinct values in the node; the oracle's nsplit is pinned against the live reference (test_oracle)
    d = make_cpiu(n, p, levels=levels, seed=dseed)
    ntree = 5
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=rule, ntree=ntree, nodesize=nodesize, seed=seed, nsplit=nsplit, membership=True)
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=seed, nsplit=nsplit)
    parity.assert_forest_equal(got, want, d.n, ntree, label=f"nsplit={nsplit}")
    assert got["split_candidates"] == int(want["candidates"][0])


def test_time_slicing_is_stable_over_many_calls():
    # regression: resuming a parked tree raced with the pick (a thread could take the "fresh tree" branch
    # while its CTA restored) -- about one illegal address per 200 sliced calls.  300 calls with a tiny
    # quantum against the build with device-side checks (-DRF_CHECK: entry lists, DFS state), in a process
    # of its own so that a fault cannot poison the rest of the suite
    import subprocess
    import sys
    from rfslam_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RFSLAM_LIB=build.CHECK_LIB, RFSLAM_QUANTUM_MS="0.02")
    assert os.path.exists(build.CHECK_LIB), "librfslam_b200_check.so missing: __graft_entry__.build() makes it"
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "slice_stress.py"), "300", "small"], env=env, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "ok 300 calls" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


# ---- parity at the sizes that are benchmarked ----------------------------------
def _tie_report(tag, res, f):
    print(f"[parity] {tag}: {res['trees_compared']} trees compared in full, {res['ties']} tie(s) within 1e-9; "
          f"{f['treeID'].size} nodes, {f['split_candidates']} split candidates")


@pytest.mark.skipif(not __import__("oracle.refdriver", fromlist=["x"]).available(), reason="oracle/_ref not present")
def test_grow_matches_live_reference_at_bench_sample():
    """The exact sample bench.py's cpu_baseline grows with the unmodified reference (n = 100k of configs[1]:
    50 covariates, 64 levels, nodesize 20, poisson.split4), tree by tree against the CUDA path."""
    d = make_cpiu(100_000, 50, levels=64)
    ntree = 16
    want = parity.live_reference_forest(d, rule=4, ntree=ntree, nodesize=20, seed=-12345)
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=ntree, nodesize=20, seed=-12345, membership=True,
                            split_stat=True, member_lists=True)
    res = parity.assert_forest_equal(got, want, d.n, ntree, check_membr_lists=True, label="ref100k", tie_aware=True,
                                     ref_stat=lambda: port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=ntree, nodesize=20, seed=-12345))
    _tie_report("live reference, 100k x 50 x 16 trees", res, got)


def test_grow_matches_oracle_at_config1_rows():
    """BASELINE configs[1] at its full row count (1M x 50, 64 levels, nodesize 20): exposure sums are 64-bit fixed
    point at 2^-43 resolution there and summation order differs from the reference's, so this is where a
    hysteresis flip of the tracker (1e-9 absolute) would first show."""
    d = make_cpiu(1_000_000, 50, levels=64)
    ntree = 4
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=ntree, nodesize=20, seed=-12345)
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=ntree, nodesize=20, seed=-12345, membership=True, split_stat=True)
    res = parity.assert_forest_equal(got, want, d.n, ntree, label="1M", tie_aware=True)
    m = want["parmID"] != 0
    if res["ties"] == 0:
        np.testing.assert_allclose(got["splitStat"][m], want["splitStat"][m], rtol=1e-9, atol=1e-9)
    _tie_report("oracle, 1M x 50 x 4 trees", res, got)


@pytest.mark.parametrize("levels", [64, 0], ids=["modeQ", "modeC"])
@pytest.mark.parametrize("rule", [1, 2, 3, 4, 5, 6])
def test_grow_matches_oracle_at_config0_shape(rule, levels):
    """BASELINE configs[0] at its real shape (examples/example.R: ~5k people x 16 half-year CPIUs = 80k rows x 30
    covariates), quantised and continuous, every rule."""
    d = make_cpiu(80_000, 30, levels=levels, seed=20260101 + rule)
    ntree = 2
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=5, seed=-12345)
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=rule, ntree=ntree, nodesize=5, seed=-12345, membership=True, split_stat=True)
    res = parity.assert_forest_equal(got, want, d.n, ntree, label=f"c0 rule {rule} L{levels}", tie_aware=True)
    _tie_report(f"oracle, 80k x 30, rule {rule}, levels {levels}", res, got)


from tests.test_oracle import UNTESTED_ARG_CASES, untested_arg_inputs  # noqa: E402


@pytest.mark.parametrize("case", UNTESTED_ARG_CASES, ids=[c["tag"] for c in UNTESTED_ARG_CASES])
def test_other_arguments_match_oracle(case):
    """xSplitStatWeight != 1, nodedepth >= 0, sampsize < n, more than 16 strata, the Bayes rules with by.user
    counts (one of them above the 64 an entry holds): the oracle is pinned to the live reference on the
    same cases (tests/test_oracle.py)."""
    d, kw = untested_arg_inputs(case, n=6000, p=12)
    ntree = 4
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=case["rule"], ntree=ntree, nodesize=5, seed=-77, **kw)
    gkw = dict(kw)
    if "split_weight" in gkw:
        gkw["split_wt"] = gkw.pop("split_weight")
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=case["rule"], ntree=ntree, nodesize=5, seed=-77, membership=True, **gkw)
    parity.assert_forest_equal(got, want, d.n, ntree, label=case["tag"])


@pytest.mark.parametrize("rule", [1, 4])
def test_zero_risk_time_nan_statistics(rule):
    d = make_cpiu(900, 5, levels=16, seed=41)
    rt = np.zeros(d.n)
    want = port.grow(d.x, d.y, rt, d.strat, rule=rule, ntree=3, nodesize=5, seed=-3)
    got = rfslam_b200.rfsrc(d.x, d.y, rt, d.strat, splitrule=rule, ntree=3, nodesize=5, seed=-3, membership=True)
    parity.assert_forest_equal(got, want, d.n, 3, label="zero-rt")


def test_bootstrap_larger_than_the_table_is_sized_or_refused():
    """Entry lists are sized by the number of entries a tree can have: a by.user row drawn 1000 times
    folds into ceil(1000 / 64) entries, and a by.root sample far above n is refused rather than overrun."""
    d = make_cpiu(500, 4, levels=8, seed=61)
    samp = np.ones((2, d.n), np.int32)
    samp[:, :40] = 1000
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=2, nodesize=5, seed=-3, samp=samp)
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=2, nodesize=5, seed=-3, samp=samp, membership=True)
    parity.assert_forest_equal(got, want, d.n, 2, label="heavy by.user")
    big = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=2, nodesize=5, seed=-3, sampsize=40 * d.n, membership=True)
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=2, nodesize=5, seed=-3, sampsize=40 * d.n)
    parity.assert_forest_equal(big, want, d.n, 2, label="sampsize 40n")


@pytest.mark.parametrize("threads", [256, 320, 512])
def test_every_cta_width_grows_the_same_forest(threads, monkeypatch):
    """The tree kernel is built at three CTA widths (csrc/grow_variant.h); which one a launch takes depends on mtry and on
    how many trees it holds.  Forced one by one here -- with enough trees to be time-sliced and mtry = 10, the case the
    ten-warp build exists for -- they must all grow the oracle's forest."""
    d = make_cpiu(2500, 30, levels=32, seed=33)
    ntree = 320
    monkeypatch.setenv("RFSLAM_CTA_THREADS", str(threads))
    monkeypatch.setenv("RFSLAM_QUANTUM_MS", "0.2")
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=1, ntree=ntree, nodesize=5, mtry=10, seed=-19, membership=True)
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=1, ntree=ntree, nodesize=5, mtry=10, seed=-19)
    parity.assert_forest_equal(got, want, d.n, ntree, label=f"cta{threads}")


def test_row_major_mirror_and_staged_subtrees_match_oracle(monkeypatch):
    """Tables beyond L2 get a row-major mirror and grow small subtrees out of a shared-memory tile (csrc/grow_kernel.cuh
    stage_subtree, opt-in: RFSLAM_SUBTREE_MAX); both forced on here for a table that would not take those paths by size."""
    monkeypatch.setenv("RFSLAM_ROWMAJOR", "1")
    monkeypatch.setenv("RFSLAM_SUBTREE_MAX", "512")
    for rule, levels, nodesize, p in ((4, 64, 5, 40), (1, 200, 3, 17), (5, 16, 10, 9)):
        d = make_cpiu(30000, p, levels=levels, seed=44 + rule)
        ntree = 3
        got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=rule, ntree=ntree, nodesize=nodesize, seed=-23, membership=True)
        want = port.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=-23)
        parity.assert_forest_equal(got, want, d.n, ntree, label=f"rowmajor rule {rule}")
    d = make_cpiu(6000, 12, levels=32, seed=50)                       # and time-sliced (staged subtrees are never parked)
    monkeypatch.setenv("RFSLAM_QUANTUM_MS", "0.05")
    monkeypatch.setenv("RFSLAM_CTA_THREADS", "256")
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=400, nodesize=5, seed=-29, membership=True)
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=4, ntree=400, nodesize=5, seed=-29)
    parity.assert_forest_equal(got, want, d.n, 400, label="rowmajor sliced")


# ---- unordered factors (MWCP splits, randomForestSRC.c:14452-14518, :1607-1616) ----------------------------------------
FACTOR_GOLDENS = ["fac_r4_det.npz", "fac_r1_rand.npz", "fac_r6_nsplit.npz", "fac_r4_big.npz"]


@pytest.mark.parametrize("name", FACTOR_GOLDENS)
def test_factor_splits_match_reference_golden(golden_dir, name):
    """Mixed numeric / unordered-factor tables (config 4's covariate kinds) against the unmodified reference: forests
    with their MWCP words (deterministic pairs, random pairs drawn from chain B, nsplit > 0), node membership, ensembles,
    and held-out predictions routed by level subset."""
    g = parity.load_golden(os.path.join(golden_dir, name))
    q = parity.golden_params(g)
    xt = list(str(g["xtype"]))
    got = rfslam_b200.rfsrc(g["x"], g["y"], g["rt"], g["strat"], splitrule=q["rule"], ntree=q["ntree"], nodesize=q["nodesize"],
                            mtry=q["mtry"], seed=q["seed"], membership=True, nsplit=q["nsplit"], xvar_types=xt, xvar_levels=g["xlevels"])
    ref = parity.golden_ref_forest(g)
    parity.assert_forest_equal(got, ref, q["n"], q["ntree"], label=name)
    assert int(got["mwcpSZ"].sum()) == int(ref["mwcpSZ"].sum()) > 50
    pr = rfslam_b200.predict_rfsrc(got, g["fx"], xvar_types=xt, xvar_levels=g["xlevels"])
    np.testing.assert_array_equal(pr["nodeMembership"], g["pred_nodeMembership"])
    np.testing.assert_allclose(pr["allEnsbCLS"], g["pred_allEnsbCLS"], rtol=parity.HAZARD_RTOL, atol=0)


@pytest.mark.parametrize("n,p,fac,rule,nodesize,seed,nsplit,levels", [
    (30000, 16, {1: 5, 4: 8, 9: 2, 12: 9}, 4, 20, -101, 0, 64),       # large nodes: histogram path + deterministic pairs
    (8000, 10, {0: 9, 3: 7, 8: 6}, 1, 2, -103, 0, 32),                # tiny nodes: random pairs nearly everywhere
    (12000, 9, {2: 4, 6: 8}, 5, 10, -105, 3, 16),                     # nsplit > 0 on factors and numerics
    (9000, 8, {0: 3, 1: 3, 2: 4, 3: 5, 4: 6, 5: 7, 6: 8, 7: 9}, 6, 5, -107, 0, 16)])   # factors only
def test_factor_splits_match_oracle_fresh_seeds(n, p, fac, rule, nodesize, seed, nsplit, levels):
    d = make_cpiu(n, p, levels=levels, seed=-seed)
    x = d.x.copy(); rng = np.random.default_rng(-seed)
    xt = ["R"] * p; xl = np.zeros(p, np.int32)
    for c, lv in fac.items():
        base = rng.integers(1, lv + 1, size=d.n)
        x[c] = ((base - 1 + ((d.y == 2.0) & (rng.random(d.n) < 0.4))) % lv + 1).astype(np.float64); xt[c] = "C"; xl[c] = lv
    ntree = 4
    want = port.grow(x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=seed, nsplit=nsplit, xtype=xt, xlevels=xl)
    got = rfslam_b200.rfsrc(x, d.y, d.rt, d.strat, splitrule=rule, ntree=ntree, nodesize=nodesize, seed=seed, membership=True,
                            nsplit=nsplit, xvar_types=xt, xvar_levels=xl, split_stat=True)
    res = parity.assert_forest_equal(got, want, d.n, ntree, label="factor fresh", tie_aware=True)
    assert res["ties"] == 0, res
    assert got["split_candidates"] == int(want["candidates"][0])
    assert int(got["mwcpSZ"].sum()) > 20


def test_factor_limits_fail_loudly():
    d = make_cpiu(600, 3, levels=8, seed=1)
    x = d.x.copy(); x[1] = (np.arange(d.n) % 12 + 1).astype(np.float64)
    with pytest.raises(rfslam_b200.RfslamError, match="levels"):       # 12 levels present: beyond the accelerated factor width
        rfslam_b200.rfsrc(x, d.y, d.rt, d.strat, ntree=1, xvar_types=["R", "C", "R"], xvar_levels=np.array([0, 12, 0], np.int32))
    x[1] = (np.arange(d.n) % 4 + 1).astype(np.float64) + 0.5
    with pytest.raises(rfslam_b200.RfslamError, match="level"):        # factor values must be the integers 1..xLevels
        rfslam_b200.rfsrc(x, d.y, d.rt, d.strat, ntree=1, xvar_types=["R", "C", "R"], xvar_levels=np.array([0, 4, 0], np.int32))


# ---- the callers of the ensemble: perf.type and the RFQ vote -----------------------------------------------
@pytest.mark.parametrize("rfq", [False, True])
@pytest.mark.parametrize("perf", ["misclass", "brier", "gmean"])
def test_perf_types_match_oracle(perf, rfq):
    """Misclassification / Brier / g-mean of the OOB ensemble, plain and RFQ vote (randomForestSRC.c:974-1218): counts are
    exact, Brier sums differ by summation order only.  The same table comes back from restore-mode predict."""
    d = make_cpiu(20000, 8, levels=32, seed=77)
    kw = dict(ntree=12, nodesize=10, seed=-19)
    want = port.grow(d.x, d.y, d.rt, d.strat, rule=4, perf=perf, rfq=rfq, **kw)["perfClas"]
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, perf_type=perf, rfq=rfq, **kw)
    for table in (got["perfClas"], rfslam_b200.predict_rfsrc(got, None, x=d.x, y=d.y, perf_type=perf, rfq=rfq)["perfClas"]):
        np.testing.assert_array_equal(np.isnan(table), np.isnan(want))
        ok = ~np.isnan(want)
        if perf == "brier":
            np.testing.assert_allclose(table[ok], want[ok], rtol=1e-12, atol=0)
        else:
            np.testing.assert_array_equal(table[ok], want[ok])


# ---- permutation importance (randomForestSRC.c:2106-2218, 2416-2697) -----------------------------------------
from tests.test_oracle import VIMP_GOLDENS, assert_vimp_close, vimp_golden_forest, vimp_modes   # noqa: E402


@pytest.mark.parametrize("name", VIMP_GOLDENS)
def test_permutation_importance_matches_reference_golden(golden_dir, name):
    """importance = "permute" (per tree) / "permute.ensemble" inside the grow call and in restore mode against the
    unmodified reference (one thread: its per-tree importance races otherwise): vote counts make the misclassification
    and g-mean importances exact, Brier sums and ensemble sums differ by summation order only."""
    g = np.load(os.path.join(golden_dir, name))
    n, p, levels, rule, ntree, nodesize, seed, dseed, byuser = [int(v) for v in g["params"]]
    xt = list(str(g["xtype"])); samp = g["samp"] if byuser else None
    forest = None
    for key, imp, perf, rfq in vimp_modes(g):
        got = rfslam_b200.rfsrc(g["x"], g["y"], g["rt"], g["strat"], splitrule=rule, ntree=ntree, nodesize=nodesize, seed=seed,
                                xvar_types=xt, xvar_levels=g["xlevels"], samp=samp, importance=imp, perf_type=perf, rfq=rfq)
        np.testing.assert_array_equal(got["parmID"], g["ref_parmID"])
        assert got["importance"].shape == (p, 3)
        assert_vimp_close(got["importance"], g[key], exact=(imp == "permute" and perf != "brier"))
        pk = "ref_perf_" + key[len("ref_vimp_"):]
        np.testing.assert_allclose(got["perfClas"][-3:][~np.isnan(g[pk])], g[pk][~np.isnan(g[pk])], rtol=1e-12, atol=0)
        forest = forest or got
    for imp in ("permute", "permute.ensemble"):
        for use in (True, False):                                     # leaf class counts passed, or rebuilt by the restore
            pr = rfslam_b200.predict_rfsrc(forest, None, x=g["x"], y=g["y"], samp=samp, xvar_types=xt, xvar_levels=g["xlevels"],
                                           importance=imp, xvar_importance=g["intr"], use_tnclas=use)
            assert_vimp_close(pr["importance"], g["ref_restore_vimp_" + imp], exact=(imp == "permute"))


@pytest.mark.parametrize("per_tree", [True, False], ids=["per-tree", "ensemble"])
def test_permutation_importance_matches_oracle_at_scale(per_tree):
    """More OOB rows than one bitmap word group (three index levels in use), trees time-sliced over the CTAs."""
    d = make_cpiu(100000, 4, levels=32, seed=91)                       # ~36.8 k OOB rows per tree: two top-level groups
    ntree, seed = 3, -41
    got = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=ntree, nodesize=20, seed=seed, membership=True,
                            importance="permute" if per_tree else "permute.ensemble")
    want = port.vimp_permute(got, d.x, d.y, got["bootMembership"], port.chain_seed_c(seed, ntree), per_tree=per_tree)
    assert_vimp_close(got["importance"], want, exact=per_tree)
    assert np.abs(want[:, 0]).max() > 0


def test_importance_limits_fail_loudly():
    d = make_cpiu(800, 4, levels=8, seed=2)
    with pytest.raises(ValueError, match="permutation"):
        rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=2, importance="anti")
    with pytest.raises(rfslam_b200.RfslamError, match="whole forest"):
        rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, ntree=4, importance="permute", tree_first=1, tree_step=2, finalize=False)


@pytest.mark.parametrize("perf", ["misclass", "brier", "gmean"])
def test_held_out_performance_matches_oracle(perf):
    """newdata that carries the outcome (frSize = 1): perfClas scores the full ensemble on the held-out rows
    (randomForestSRC.c:1323-1326); a row shard returns no table (its counts are partial)."""
    d = make_cpiu(20000, 8, levels=32, seed=78); t = make_cpiu(6000, 8, levels=32, seed=79)
    f = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=10, nodesize=10, seed=-29)
    pr = rfslam_b200.predict_rfsrc(f, t.x, x=d.x, y=d.y, fy=t.y, perf_type=perf)
    want = port.perf_clas(t.y, pr["allEnsbCLS"], perf=perf)
    got = pr["perfClas"][-3:]
    assert np.isnan(pr["perfClas"][:-3]).all()
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12 if perf == "brier" else 0, atol=0)
    assert "perfClas" not in rfslam_b200.predict_rfsrc(f, t.x, x=d.x, y=d.y, fy=t.y, frow_first=0, frow_count=3000)
    assert "perfClas" not in rfslam_b200.predict_rfsrc(f, t.x, fy=None)


@pytest.mark.parametrize("per_tree", [True, False], ids=["per-tree", "ensemble"])
def test_importance_on_held_out_rows_matches_oracle(per_tree):
    """predict(newdata with the outcome, importance = ...): every held-out row is permuted and dropped down every tree
    (membershipIndex = identity, randomForestSRC.c:2125-2131), chain C seeded as in restore mode."""
    d = make_cpiu(12000, 7, levels=32, seed=80); t = make_cpiu(5000, 7, levels=32, seed=81)
    ntree = 6
    f = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=4, ntree=ntree, nodesize=10, seed=-33)
    intr = [4, 1, 7]
    pr = rfslam_b200.predict_rfsrc(f, t.x, x=d.x, y=d.y, fy=t.y, importance="permute" if per_tree else "permute.ensemble", xvar_importance=intr)
    want = port.vimp_permute(f, t.x, t.y, np.zeros((ntree, t.n), np.int32), port.chain_seed_c(0, ntree, restore=True), intr=intr, per_tree=per_tree)
    assert_vimp_close(pr["importance"], want, exact=per_tree)
    with pytest.raises(rfslam_b200.RfslamError, match="whole row set"):
        rfslam_b200.predict_rfsrc(f, t.x, x=d.x, y=d.y, fy=t.y, importance="permute", frow_first=0, frow_count=1000)
