"""Regenerate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/librfslam_ref.so).

Run in the build container only (needs /root/reference to have been compiled by
`make -C oracle ref`):   python tests/golden/make_golden.py
Each file holds the seeded synthetic inputs and the reference's own outputs for them, with the
forest records re-ordered by treeID (the reference appends trees in thread-completion order,
randomForestSRC.c:20938-20946) and trimmed the way R/rfsrc.R:1766-1861 trims them.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refdriver as rd          # noqa: E402
from rfslam_b200.synth import make_cpiu     # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def canon(r, n):
    t = rd.trim_forest(r, n)
    order = np.argsort(t["treeID"], kind="stable")
    if "mwcpPT" in t and t["mwcpPT"].size:                 # the MWCP words travel with their records
        w = np.zeros(t["treeID"].size, np.int64); w[t["mwcpSZ"] > 0] = t["mwcpPT"]
        t["mwcpPT"] = w[order][t["mwcpSZ"][order] > 0].astype(np.int32)
    for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ"):
        t[k] = t[k][order]
    return t


# unordered-factor covariates (xType "C", MWCP splits rf.c:14452-14518): name, n, p, {column: levels}, rule, ntree, nodesize,
# mtry, forest seed, data seed, nsplit.  Small nodesize -> random pairs (pairs >= node size), large -> deterministic.
FACTOR_CASES = [
    ("fac_r4_det", 3000, 8, {1: 5, 4: 8, 6: 3}, 4, 5, 20, None, -7, 5, 0),
    ("fac_r1_rand", 2500, 6, {0: 9, 3: 6}, 1, 4, 3, None, -11, 6, 0),
    ("fac_r6_nsplit", 3000, 7, {2: 7, 5: 4}, 6, 4, 5, None, -13, 7, 4),
    ("fac_r4_big", 20000, 12, {0: 2, 3: 9, 7: 6, 10: 8}, 4, 3, 10, None, -17, 8, 0),
]


def factor_table(n, p, fac, dseed):
    d = make_cpiu(n, p, levels=16, seed=dseed)
    x = d.x.copy()
    rng = np.random.default_rng(dseed + 77)
    xt = ["R"] * p; xl = np.zeros(p, np.int32)
    for c, L in fac.items():
        # levels tied to the outcome so that factor splits win often; every level occurs
        base = rng.integers(1, L + 1, size=d.n)
        shift = ((d.y == 2.0) & (rng.random(d.n) < 0.5)).astype(np.int64)
        x[c] = ((base - 1 + shift) % L + 1).astype(np.float64)
        xt[c] = "C"; xl[c] = L
    return d, x, xt, xl


def main_factors():
    for (name, n, p, fac, rule, ntree, nodesize, mtry, seed, dseed, nsplit) in FACTOR_CASES:
        d, x, xt, xl = factor_table(n, p, fac, dseed)
        r = rd.grow(x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, mtry=mtry, seed=seed, nsplit=nsplit,
                    xtype=xt, xlevels=xl)
        t = canon(r, d.n)
        dt, fx, _, _ = factor_table(max(300, n // 4), p, fac, dseed + 1000)
        pr = rd.predict(x, d.y, t, fx, nodesize=nodesize, xtype=xt, xlevels=xl)
        save = dict(x=x, y=d.y, rt=d.rt, strat=d.strat, xtype=np.array("".join(xt)), xlevels=xl,
                    params=np.array([d.n, p, 16, rule, ntree, nodesize, -1 if mtry is None else mtry, seed, dseed, 0, nsplit], np.int64),
                    fx=fx, pred_allEnsbCLS=pr["allEnsbCLS"], pred_nodeMembership=pr["nodeMembership"])
        for k in ("oobEnsbCLS", "allEnsbCLS", "leafCount", "seed", "nodeMembership", "bootMembership",
                  "treeID", "nodeID", "parmID", "contPT", "mwcpSZ", "mwcpPT", "mwcpCT", "tnRCNT", "tnACNT", "perfClas"):
            save["ref_" + k] = t[k]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print(name, "nodes", t["treeID"].size, "factor splits", int(t["mwcpSZ"].sum()), "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


GROW_CASES = [
    # name, n, p, levels, rule, ntree, nodesize, mtry, forest seed, data seed, by.user
    ("grow_r1_q64", 1500, 9, 64, 1, 6, 5, None, -12345, 20260101, False),
    ("grow_r2_q64", 1500, 9, 64, 2, 6, 5, None, -12345, 20260101, False),
    ("grow_r3_q64", 1500, 9, 64, 3, 6, 5, None, -12345, 20260101, False),
    ("grow_r4_q64", 1500, 9, 64, 4, 6, 5, None, -12345, 20260101, False),
    ("grow_r5_q64", 1500, 9, 64, 5, 6, 5, None, -12345, 20260101, False),
    ("grow_r6_q64", 1500, 9, 64, 6, 6, 5, None, -12345, 20260101, False),
    ("grow_r4_cont", 2500, 12, 0, 4, 5, 20, None, -777, 11, False),
    ("grow_r1_cont", 1200, 7, 0, 1, 5, 3, 7, -31, 12, False),
    ("grow_r4_byuser", 1800, 10, 32, 4, 5, 5, None, -99, 13, True),
    ("grow_r4_q256_big", 12000, 20, 256, 4, 4, 20, None, -4242, 14, False),
    ("grow_r1_mtry1", 900, 6, 16, 1, 8, 2, 1, -5, 15, False),
    # nsplit > 0 (12th field): random cut subsets, rf.c:14533-14564
    ("grow_r4_q64_nsplit5", 1500, 9, 64, 4, 6, 5, None, -12345, 20260101, False, 5),
    ("grow_r1_q256_nsplit10", 6000, 12, 256, 1, 4, 10, None, -4242, 16, False, 10),
]


def main():
    for case in GROW_CASES:
        (name, n, p, levels, rule, ntree, nodesize, mtry, seed, dseed, byuser) = case[:11]
        nsplit = case[11] if len(case) > 11 else 0
        d = make_cpiu(n, p, levels=levels, seed=dseed)
        samp = None
        if byuser:
            rng = np.random.default_rng(dseed + 1)
            samp = rng.multinomial(int(0.8 * n), np.full(n, 1.0 / n), size=ntree).astype(np.int32)
        r = rd.grow(d.x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, mtry=mtry,
                    seed=seed, samp=samp, nsplit=nsplit)
        t = canon(r, n)
        dt = make_cpiu(max(200, n // 3), p, levels=levels, seed=dseed + 1000)
        pr = rd.predict(d.x, d.y, t, dt.x, samp=samp, nodesize=nodesize)
        save = dict(
            x=d.x, y=d.y, rt=d.rt, strat=d.strat,
            params=np.array([n, p, levels, rule, ntree, nodesize, -1 if mtry is None else mtry, seed, dseed, int(byuser)] + ([nsplit] if nsplit else []), np.int64),
            fx=dt.x, pred_allEnsbCLS=pr["allEnsbCLS"], pred_nodeMembership=pr["nodeMembership"])
        if samp is not None:
            save["samp"] = samp
        for k in ("oobEnsbCLS", "allEnsbCLS", "leafCount", "seed", "nodeMembership", "bootMembership",
                  "treeID", "nodeID", "parmID", "contPT", "mwcpSZ", "tnRCNT", "tnACNT",
                  "rmbrMembership", "ambrMembership", "perfClas"):
            save["ref_" + k] = t[k]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print(name, "nodes", t["treeID"].size, "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))

    # known-answer vectors for the six plugin functions themselves (splitCustom.c:639-1687),
    # called the way randomForestSRC.c:13469-13482 calls them (1-based arrays, featureCount = 2)
    L = rd.lib()
    rng = np.random.default_rng(7)
    kat = {}
    proto = C.CFUNCTYPE(C.c_double, C.c_uint, C.c_double, C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint,
                        C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_uint,
                        C.POINTER(C.POINTER(C.c_double)), C.c_uint)
    cases = []
    for m in (2, 3, 17, 64, 333, 2000):
        for K in (1, 4, 16):
            cases.append((m, K))
    idx = 0
    for (m, K) in cases:
        resp = 1.0 + (rng.random(m) < 0.2)
        strat = rng.integers(1, K + 1, m).astype(np.float64)
        rtv = np.where(rng.random(m) < 0.8, 0.5, rng.uniform(0.01, 0.5, m))
        left = max(1, min(m - 1, int(rng.integers(1, m))))
        memb = np.zeros(m, np.int8); memb[:left] = 1
        rng.shuffle(memb)
        kfa = float(rng.choice([0.5, 1.0, 2.0, 3.0]))
        stats = np.zeros(6)
        r1 = np.concatenate([[0.0], resp]); s1 = np.concatenate([[0.0], strat]); t1 = np.concatenate([[0.0], rtv])
        m1 = np.concatenate([[0], memb]).astype(np.int8)
        feat = (C.POINTER(C.c_double) * 3)(None, s1.ctypes.data_as(C.POINTER(C.c_double)), t1.ctypes.data_as(C.POINTER(C.c_double)))
        for rule in range(1, 7):
            fn = proto(("poissonSplit%d" % rule, L))
            stats[rule - 1] = fn(m, kfa, m1.ctypes.data_as(C.c_char_p), None, None, 0, 0, None,
                                 r1.ctypes.data_as(C.POINTER(C.c_double)), 0.0, 0.0, 2, feat, 2)
        kat["resp_%d" % idx] = resp; kat["strat_%d" % idx] = strat; kat["rt_%d" % idx] = rtv
        kat["memb_%d" % idx] = memb; kat["kfa_%d" % idx] = np.array([kfa]); kat["stat_%d" % idx] = stats
        idx += 1
    kat["count"] = np.array([idx])
    np.savez_compressed(os.path.join(OUT, "plugin_kat.npz"), **kat)
    print("plugin_kat", idx, "cases")

    # RNG known answers: chain seeds and the first draws of the bootstrap, read back from the
    # reference's own outputs (seed[] and bootMembership are enough to pin chain A; chain B is
    # pinned by the grown trees).
    print("done")


# permutation importance (rf.c:2106-2218, 2416-2697): name, n, p, levels, factor columns, rule, ntree, nodesize, forest seed, data seed,
# by.user.  One forest per case (chain C does not touch the trees), importance in every mode / perf.type, in grow and in restore.
# ONE reference thread: with OPT_VIMP_LEOB every tree resets and refills the shared RF_vimpEnsembleDen[p] / RF_vimpEnsembleCLS[p]
# (rf.c:2265-2280) without a lock, so the reference's per-tree importance is only reproducible single-threaded.
VIMP_CASES = [
    ("vimp_q16", 1500, 8, 16, {}, 4, 6, 10, -21, 31, False),
    ("vimp_fac", 2000, 7, 16, {1: 5, 4: 7}, 4, 5, 8, -23, 32, False),
    ("vimp_byuser", 1200, 6, 32, {}, 1, 5, 5, -25, 33, True),
]
VIMP_MODES = [("permute", "misclass", False), ("permute.ensemble", "misclass", False), ("permute", "brier", False),
              ("permute.ensemble", "brier", False), ("permute", "gmean", False), ("permute", "misclass", True),
              ("permute.ensemble", "gmean", True)]


def main_vimp():
    for (name, n, p, levels, fac, rule, ntree, nodesize, seed, dseed, byuser) in VIMP_CASES:
        if fac:
            d, x, xt, xl = factor_table(n, p, fac, dseed)
        else:
            d = make_cpiu(n, p, levels=levels, seed=dseed); x = d.x; xt = ["R"] * p; xl = np.zeros(p, np.int32)
        samp = None
        if byuser:
            rng = np.random.default_rng(dseed + 5)
            samp = np.stack([np.bincount(rng.integers(0, d.n, d.n), minlength=d.n) for _ in range(ntree)]).astype(np.int32)
        save = dict(x=x, y=d.y, rt=d.rt, strat=d.strat, xtype=np.array("".join(xt)), xlevels=xl,
                    params=np.array([d.n, p, levels, rule, ntree, nodesize, seed, dseed, int(byuser)], np.int64))
        if samp is not None:
            save["samp"] = samp
        t = None
        for (imp, perf, rfq) in VIMP_MODES:
            r = rd.grow(x, d.y, d.rt, d.strat, rule=rule, ntree=ntree, nodesize=nodesize, seed=seed, xtype=xt, xlevels=xl, samp=samp,
                        importance=imp, perf=perf, rfq=rfq, quants=True, nthreads=1)
            key = f"{imp}_{perf}_{int(rfq)}"
            save["ref_vimp_" + key] = r["vimpClas"]
            save["ref_perf_" + key] = r["perfClas"][-3:]
            if t is None:
                t = canon(r, d.n)
                for k in ("leafCount", "seed", "bootMembership", "treeID", "nodeID", "parmID", "contPT", "mwcpSZ", "tnRCNT", "tnACNT", "tnCLAS") + (("mwcpPT",) if fac else ()):
                    save["ref_" + k] = t[k]
        # restore mode (vimp.rfsrc -> rfsrcPredict on the training rows): chain C comes from the seed LCG started at 0
        intr = np.array([3, 1, p], np.int32)
        save["intr"] = intr
        for imp in ("permute", "permute.ensemble"):
            pr = rd.predict(x, d.y, t, None, nodesize=nodesize, xtype=xt, xlevels=xl, samp=samp, importance=imp, xvar_imp=intr, quants=True, nthreads=1)
            save["ref_restore_vimp_" + imp] = pr["vimpClas"]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print(name, "nodes", t["treeID"].size, "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "vimp":
    main_vimp()
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "factors":
    main_factors()
    sys.exit(0)

if __name__ == "__main__":
    main()
