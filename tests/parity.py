"""Shared comparison logic for parity tests (oracle port, reference, CUDA path).

A forest here is a dict of numpy arrays in the reference's output vocabulary
(randomForestSRC.c:9-79): treeID/nodeID/parmID/contPT in pre-order and tree order, leafCount, seed,
bootMembership/nodeMembership [ntree*n], tnRCNT/tnACNT concatenated per tree, oobEnsbCLS/allEnsbCLS
[2*n] class-major.
"""
from __future__ import annotations

import glob
import os
from typing import Dict, List

import numpy as np

HAZARD_RTOL = 1.0e-6       # north_star: terminal / OOB / predicted hazards within 1e-6 relative
TIE_RTOL = 1.0e-9          # north_star: split statistics that tie within 1e-9 relative may differ


def load_golden(path: str) -> Dict[str, np.ndarray]:
    z = np.load(path)
    return {k: z[k] for k in z.files}


def golden_files(golden_dir: str) -> List[str]:
    return sorted(glob.glob(os.path.join(golden_dir, "grow_*.npz")))


def golden_params(g) -> dict:
    vals = [int(v) for v in g["params"]]
    n, p, levels, rule, ntree, nodesize, mtry, seed, dseed, byuser = vals[:10]
    nsplit = vals[10] if len(vals) > 10 else 0
    return dict(n=n, p=p, levels=levels, rule=rule, ntree=ntree, nodesize=nodesize,
                mtry=None if mtry < 0 else mtry, seed=seed, dseed=dseed, byuser=bool(byuser), nsplit=nsplit)


def split_tree(f: Dict[str, np.ndarray], b: int):
    sel = f["treeID"] == b
    return {k: f[k][sel] for k in ("nodeID", "parmID", "contPT")}


def mwcp_by_record(f: Dict[str, np.ndarray]) -> np.ndarray:
    """The MWCP word of every record (0 = numeric split or leaf): mwcpPT holds one word per record with mwcpSZ = 1,
    in record order (saveTree, randomForestSRC.c:10773-10779)."""
    sz = np.asarray(f["mwcpSZ"])
    out = np.zeros(sz.size, np.int64)
    assert sz.max(initial=0) <= 1
    out[sz > 0] = np.asarray(f["mwcpPT"]).astype(np.int64)[: int(sz.sum())] & 0xFFFFFFFF
    return out


def _is_tie(a: float, b: float) -> bool:
    """Two winning split statistics "tie" when they agree within TIE_RTOL relative (north_star) or within
    the reference's own absolute hysteresis EPSILON = 1e-9 (randomForestSRC.c:15502 replaces the running
    best only when delta - best > 1e-9, so two cuts closer than that are interchangeable by construction)."""
    if np.isnan(a) or np.isnan(b):
        return bool(np.isnan(a) and np.isnan(b))
    return abs(a - b) <= max(TIE_RTOL * max(abs(a), abs(b)), 1.0e-9 * (1.0 + 1.0e-3))


def assert_forest_equal(got: Dict[str, np.ndarray], ref: Dict[str, np.ndarray], n: int, ntree: int,
                        *, check_membr_lists: bool = False, label: str = "", tie_aware: bool = False,
                        ref_stat=None) -> Dict[str, int]:
    """Bit-exact on everything integer (in-bag counts, node membership, topology, split variables,
    leaf counts), exact on cutpoints (they are data values), HAZARD_RTOL on the ensembles.

    ``tie_aware`` is north_star's rule: where the chosen split variable / cutpoint of a node differ, the two
    winning split statistics must tie (see _is_tie) -- then the rest of THAT tree is not compared (chain B is
    consumed in DFS order, randomForestSRC.c:21666-21691, so everything after a different subtree may draw
    different covariates) and the tree is left out of the per-row comparisons.  It needs ``got["splitStat"]``
    and the reference's statistic per pre-order record: ``ref["splitStat"]`` or, lazily, ``ref_stat()``
    returning such a forest (only called when a divergence is met).  Returns {"ties": k, "trees_compared": m};
    ensembles are compared only when no tree was dropped."""
    np.testing.assert_array_equal(got["seed"], ref["seed"], err_msg=label + " seed")
    np.testing.assert_array_equal(got["bootMembership"], ref["bootMembership"], err_msg=label + " bootMembership")
    dropped = []
    stat_src = ref if "splitStat" in ref else None
    for b in range(1, ntree + 1):
        tg, tr = split_tree(got, b), split_tree(ref, b)
        same = (tg["parmID"].size == tr["parmID"].size and np.array_equal(tg["parmID"], tr["parmID"])
                and np.array_equal(tg["nodeID"], tr["nodeID"])
                and np.array_equal(np.isnan(tg["contPT"]), np.isnan(tr["contPT"]))
                and np.array_equal(tg["contPT"][~np.isnan(tr["contPT"])], tr["contPT"][~np.isnan(tr["contPT"])]))
        if same:
            continue
        if not tie_aware:
            np.testing.assert_array_equal(tg["parmID"], tr["parmID"], err_msg=f"{label} tree {b} parmID")
            np.testing.assert_array_equal(tg["nodeID"], tr["nodeID"], err_msg=f"{label} tree {b} nodeID")
            np.testing.assert_array_equal(np.isnan(tg["contPT"]), np.isnan(tr["contPT"]), err_msg=f"{label} tree {b} contPT NA")
            m = ~np.isnan(tr["contPT"])
            np.testing.assert_array_equal(tg["contPT"][m], tr["contPT"][m], err_msg=f"{label} tree {b} contPT")
        # first pre-order record at which the two trees differ
        k = 0
        lim = min(tg["parmID"].size, tr["parmID"].size)
        while k < lim and tg["parmID"][k] == tr["parmID"][k] and tg["nodeID"][k] == tr["nodeID"][k] and (
                (np.isnan(tg["contPT"][k]) and np.isnan(tr["contPT"][k])) or tg["contPT"][k] == tr["contPT"][k]):
            k += 1
        assert k < lim, f"{label} tree {b}: one tree is a prefix of the other"
        assert tg["parmID"][k] != 0 and tr["parmID"][k] != 0 and tg["nodeID"][k] == tr["nodeID"][k], \
            f"{label} tree {b} record {k}: split / leaf disagreement (parmID {tg['parmID'][k]} vs {tr['parmID'][k]}) is not a tie"
        if stat_src is None:
            assert ref_stat is not None, f"{label} tree {b} record {k}: trees differ and no reference split statistic is available"
            stat_src = ref_stat()
        sg = got["splitStat"][got["treeID"] == b][k]
        sr = stat_src["splitStat"][stat_src["treeID"] == b][k]
        assert _is_tie(float(sg), float(sr)), \
            f"{label} tree {b} record {k}: different split ({tg['parmID'][k]}, {tg['contPT'][k]!r}) vs ({tr['parmID'][k]}, {tr['contPT'][k]!r}) " \
            f"with statistics {sg!r} vs {sr!r}: not a tie"
        dropped.append(b)
    keep = np.ones(ntree, bool)
    for b in dropped:
        keep[b - 1] = False
    if "mwcpPT" in ref and np.asarray(ref["mwcpSZ"]).sum() > 0 and not dropped:
        # factor splits: the LEFT level sets, record by record, and the per-tree word counts
        np.testing.assert_array_equal(got["mwcpSZ"], ref["mwcpSZ"], err_msg=label + " mwcpSZ")
        np.testing.assert_array_equal(mwcp_by_record(got), mwcp_by_record(ref), err_msg=label + " mwcpPT")
        np.testing.assert_array_equal(got["mwcpCT"], ref["mwcpCT"], err_msg=label + " mwcpCT")
    if not dropped:
        np.testing.assert_array_equal(got["leafCount"], ref["leafCount"], err_msg=label + " leafCount")
        np.testing.assert_array_equal(got["nodeMembership"], ref["nodeMembership"], err_msg=label + " nodeMembership")
        np.testing.assert_array_equal(got["tnRCNT"], ref["tnRCNT"], err_msg=label + " tnRCNT")
        np.testing.assert_array_equal(got["tnACNT"], ref["tnACNT"], err_msg=label + " tnACNT")
        for k in ("oobEnsbCLS", "allEnsbCLS"):
            a, r = got[k], ref[k]
            np.testing.assert_array_equal(np.isnan(a), np.isnan(r), err_msg=f"{label} {k} NA pattern")
            ok = ~np.isnan(r)
            np.testing.assert_allclose(a[ok], r[ok], rtol=HAZARD_RTOL, atol=0, err_msg=f"{label} {k}")
        if "perfClas" in ref and "perfClas" in got:
            # OOB misclassification rates: ratios of integer counts, exact; NA_REAL (payload 1954) elsewhere
            np.testing.assert_array_equal(got["perfClas"].view(np.uint64), ref["perfClas"].view(np.uint64), err_msg=label + " perfClas")
        if check_membr_lists:
            np.testing.assert_array_equal(got["rmbrMembership"][: ref["rmbrMembership"].size], ref["rmbrMembership"], err_msg=label + " rmbr")
            np.testing.assert_array_equal(got["ambrMembership"], ref["ambrMembership"], err_msg=label + " ambr")
    else:
        # trees that tied are compared up to the tie only; every other tree still has to match row by row
        np.testing.assert_array_equal(got["leafCount"][keep], ref["leafCount"][keep], err_msg=label + " leafCount")
        gm, rm = got["nodeMembership"].reshape(ntree, n), ref["nodeMembership"].reshape(ntree, n)
        np.testing.assert_array_equal(gm[keep], rm[keep], err_msg=label + " nodeMembership")
        lg = np.concatenate([[0], np.cumsum(got["leafCount"].astype(np.int64))])
        lr = np.concatenate([[0], np.cumsum(ref["leafCount"].astype(np.int64))])
        for b in np.nonzero(keep)[0]:
            for kk in ("tnRCNT", "tnACNT"):
                np.testing.assert_array_equal(got[kk][lg[b]:lg[b + 1]], ref[kk][lr[b]:lr[b + 1]], err_msg=f"{label} tree {b + 1} {kk}")
    return {"ties": len(dropped), "trees_compared": int(keep.sum())}


def golden_ref_forest(g) -> Dict[str, np.ndarray]:
    return {k[4:]: v for k, v in g.items() if k.startswith("ref_")}


def live_reference_forest(d, *, statistics: bool = False, **kw) -> Dict[str, np.ndarray]:
    """The unmodified reference (oracle/_ref) run on a synthetic table, trimmed as R trims it and put in
    tree order (it appends trees in thread-completion order, randomForestSRC.c:20938-20946)."""
    from oracle import refdriver
    if statistics:
        kw["nthreads"] = 1
    r = refdriver.grow(d.x, d.y, d.rt, d.strat, statistics=statistics, **kw)
    t = refdriver.trim_forest(r, d.n)
    order = np.argsort(t["treeID"], kind="stable")
    if "mwcpPT" in t and t["mwcpPT"].size:
        w = np.zeros(t["treeID"].size, np.int64); w[t["mwcpSZ"] > 0] = t["mwcpPT"]      # words travel with their records
        t["mwcpPT"] = w[order][t["mwcpSZ"][order] > 0].astype(np.int32)
    for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ") + (("splitStat",) if "splitStat" in t else ()):
        t[k] = t[k][order]
    return t
