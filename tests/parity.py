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


def assert_forest_equal(got: Dict[str, np.ndarray], ref: Dict[str, np.ndarray], n: int, ntree: int,
                        *, check_membr_lists: bool = False, label: str = "") -> None:
    """Bit-exact on everything integer (in-bag counts, node membership, topology, split variables,
    leaf counts), exact on cutpoints (they are data values), HAZARD_RTOL on the ensembles."""
    np.testing.assert_array_equal(got["seed"], ref["seed"], err_msg=label + " seed")
    np.testing.assert_array_equal(got["bootMembership"], ref["bootMembership"], err_msg=label + " bootMembership")
    np.testing.assert_array_equal(got["leafCount"], ref["leafCount"], err_msg=label + " leafCount")
    for b in range(1, ntree + 1):
        tg, tr = split_tree(got, b), split_tree(ref, b)
        np.testing.assert_array_equal(tg["parmID"], tr["parmID"], err_msg=f"{label} tree {b} parmID")
        np.testing.assert_array_equal(tg["nodeID"], tr["nodeID"], err_msg=f"{label} tree {b} nodeID")
        np.testing.assert_array_equal(np.isnan(tg["contPT"]), np.isnan(tr["contPT"]), err_msg=f"{label} tree {b} contPT NA")
        m = ~np.isnan(tr["contPT"])
        np.testing.assert_array_equal(tg["contPT"][m], tr["contPT"][m], err_msg=f"{label} tree {b} contPT")
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


def golden_ref_forest(g) -> Dict[str, np.ndarray]:
    return {k[4:]: v for k, v in g.items() if k.startswith("ref_")}
