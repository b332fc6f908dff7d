"""CPIU construction: the restatement of utilities/cpiu.cpp against the UNMODIFIED reference file compiled
over a stand-in Rcpp.h (CPU), and the device kernels against both (GPU)."""
import numpy as np
import pytest

from oracle import cpiu_port as cp


def _cohort(persons, seed, width=6):
    rng = np.random.default_rng(seed)
    t_out = np.round(rng.uniform(30, 2900, persons), 1)            # days of follow-up
    t_out[rng.random(persons) < 0.1] = 182.5 * rng.integers(1, 9, size=int((rng.random(persons) < 0.1).sum()) or 1)[0]   # exact interval multiples
    death = (rng.random(persons) < 0.3).astype(np.int32)
    t_death = np.where(death == 1, t_out, t_out + rng.uniform(1, 400, persons))
    ev = rng.uniform(-50, 3000, (persons, width))
    ev[rng.random((persons, width)) < 0.5] = np.nan                 # missing slots anywhere
    ev[:, 0] = np.where(rng.random(persons) < 0.2, 182.5 * rng.integers(1, 5, persons), ev[:, 0])   # events exactly on a boundary
    mt = np.sort(rng.uniform(0, 2500, (persons, width)), axis=1)[:, ::-1].copy()   # measurements out of order
    mt[:, 0] = 0.0
    nmiss = rng.integers(0, width - 1, persons)
    for q in range(persons):
        if nmiss[q]:
            mt[q, -nmiss[q]:] = np.nan                              # missing measurements trail
    mv = np.round(rng.normal(50, 10, (persons, width)), 2)
    mv[rng.random((persons, width)) < 0.05] = -1.0                  # a value of -1 reads as NA (cpiu.R:642)
    return t_out, t_death, death, ev, mt, mv


@pytest.mark.skipif(not cp.ref_available(), reason="oracle/_ref/libcpiu_ref.so not built")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_cpiu_helpers_restatement_matches_reference_cpp(seed):
    """The four helpers of cpiu.cpp (indicator, accumulator flag 0 / 1, time elapsed, continuous with k.to.persist
    0 / 2 / -1) on random persons, including events on interval boundaries and before time 0."""
    t_out, _, _, ev, mt, mv = _cohort(200, seed)
    for q in range(200):
        ni = int(np.ceil(t_out[q] / 182.5))
        ts = 182.5 * np.arange(ni); te = np.minimum(ts + 182.5, t_out[q])
        e = np.where(np.isnan(ev[q]), -1.0, ev[q])
        for kind in (cp.INDICATOR, cp.COUNT_PREV, cp.COUNT_CUR, cp.ELAPSED):
            np.testing.assert_array_equal(cp.port_helper(kind, e, None, ts, te), cp.ref_helper(kind, e, None, ts, te))
        keep = ~np.isnan(mt[q]); order = np.argsort(mt[q][keep], kind="stable")
        for k in (0, 2, -1):
            np.testing.assert_array_equal(cp.port_helper(cp.CONTINUOUS, mt[q][keep][order], mv[q][keep][order], ts, te, k),
                                          cp.ref_helper(cp.CONTINUOUS, mt[q][keep][order], mv[q][keep][order], ts, te, k))


def test_cpiu_port_interval_table():
    """cpiu.R:567-577 on a hand-checked person: 400 days of follow-up at half-year intervals."""
    r = cp.cpiu_fc(np.array([400.0]), np.array([400.0]), np.array([1], np.int32), use_ref=False)
    np.testing.assert_array_equal(r["int_n"], [1, 2, 3])
    np.testing.assert_array_equal(r["t_start"], [0.0, 182.5, 365.0])
    np.testing.assert_array_equal(r["t_end"], [182.5, 365.0, 400.0])
    np.testing.assert_allclose(r["rt"], [0.5, 0.5, 35.0 / 365.0], rtol=0, atol=0)
    np.testing.assert_array_equal(r["i_death"], [0, 0, 1])


@pytest.mark.gpu
@pytest.mark.parametrize("persons,seed,k", [(3000, 11, 0), (5000, 12, 2), (2000, 13, -1)])
def test_cpiu_kernels_match_reference(persons, seed, k):
    import rfslam_b200
    t_out, t_death, death, ev, mt, mv = _cohort(persons, seed)
    kinds = [cp.INDICATOR, cp.COUNT_PREV, cp.COUNT_CUR, cp.ELAPSED, cp.CONTINUOUS, cp.INDICATOR]
    want = cp.cpiu_fc(t_out, t_death, death, var_kind=kinds, var_times=[ev, ev, ev, ev, mt, ev[:, :1]], var_values=[None] * 4 + [mv, None], k_to_persist=k)
    got = rfslam_b200.cpiu_fc(t_out, t_death, death, k_to_persist=k,
                              to_create={"i.hf": ev, "n.hf": ev, "ni.hf": ev, "t.hf": ev, "c.lab": {"times": mt, "values": mv}, "i.first": ev[:, :1]})
    for key in ("person", "int_n", "t_start", "t_end", "rt", "i_death"):
        np.testing.assert_array_equal(got[key], want[key], err_msg=key)
    np.testing.assert_array_equal(got["vars"].view(np.uint64), want["vars"].view(np.uint64))     # incl. NA_REAL where a value is -1
    assert got["rt"].size == int(np.ceil(t_out / 182.5).sum())


@pytest.mark.gpu
def test_cpiu_bad_inputs_fail_loudly():
    import rfslam_b200
    with pytest.raises(rfslam_b200.RfslamError):
        rfslam_b200.cpiu_fc(np.array([0.0]), np.array([1.0]), np.array([0], np.int32))
    with pytest.raises(ValueError, match="naming convention"):
        rfslam_b200.cpiu_fc(np.array([10.0]), np.array([11.0]), np.array([0], np.int32), to_create={"x.hf": np.zeros((1, 2))})
