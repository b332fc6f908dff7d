#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (trees/sec; split candidates scored/sec; % of HBM peak) on
BASELINE.json's configuration, for the B200-native path and, with --impl reference, for the
reference's own CPU implementation on the box's host cores.

  python bench.py --gpus N --steps K --warmup W            (N = 1: plain python)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

One step = one pass of the hot path over the workload: grow the whole forest (bootstrap replay, split
scoring, partition, leaf estimates, OOB / full ensembles) from the CPIU table resident in HBM.
Workload at N = 1 is BASELINE.json configs[1]: 1M CPIUs x 50 covariates, ntree = 500, mtry = ceil(sqrt(p)) = 8,
nodesize = 20, rule poisson.split4, cardinality mode Q (64 levels).  At N > 1 every rank holds a replica of
the table and grows 500 trees of a 500*N-tree forest (weak scaling; trees are independent, no collective
during growth); the step ends with the path's one exchange, inside the library and on device buffers:
an NCCL all-reduce of the ensemble sums and an NCCL gather of the forest records to rank 0.

Two more records ride on the same JSON line:
  c3      BASELINE configs[2], the north-star target: ONE 1000-tree forest on 10M CPIUs x 100 covariates
          (mtry 10), its trees sharded over the N ranks (strong scaling: total work fixed), one timed step
  predict BASELINE configs[4] scaled to the box: rfsrcPredict of the step's forest on held-out rows sharded
          over the N ranks (forest replicated, no collective), row.trees/s and the B_pred roofline
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# torchrun exports OMP_NUM_THREADS=1; the reference arm / CPU baseline (OpenMP over trees,
# randomForestSRC.c:7082-7087) must see every host core, so set it before libgomp is loaded
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

WORKLOAD = dict(n=1_000_000, p=50, levels=64, ntree=500, nodesize=20, rule=4, seed=-12345, k_for_alpha=2.0)
C3 = dict(n=10_000_000, p=100, levels=64, ntree=1000, nodesize=20, rule=4, seed=-12345, k_for_alpha=2.0)
REF_SAMPLE_N = 200_000           # rows of the bounded sample the reference arm / cpu_baseline grow (25 calls must fit a few minutes)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index                       # one GPU index or a comma-separated list (one sampler for the whole job)
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "interval_ms": 500}


def make_data(cfg):
    from rfslam_b200.synth import make_cpiu
    return make_cpiu(cfg["n"], cfg["p"], levels=cfg["levels"])


def cpu_reference_sample(cfg, *, sample_n: int, ntree: int, threads: int, membership: bool = False):
    """The reference's own OpenMP CPU path (oracle/_ref, the unmodified reference compiled against
    the fake-R shim) on a bounded sample of the same synthetic workload."""
    from oracle import refdriver
    from rfslam_b200.synth import make_cpiu
    d = make_cpiu(sample_n, cfg["p"], levels=cfg["levels"])
    t0 = time.perf_counter()
    r = refdriver.grow(d.x, d.y, d.rt, d.strat, rule=cfg["rule"], ntree=ntree, nodesize=cfg["nodesize"], seed=cfg["seed"],
                       k_for_alpha=cfg["k_for_alpha"], membership=membership, nthreads=threads)
    dt = time.perf_counter() - t0
    return d, r, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refdriver
    cfg = dict(WORKLOAD)
    cores = os.cpu_count() or 1
    if not refdriver.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/librfslam_ref.so not built"}))
        return 0
    sample_n = args.ref_sample_n
    ntree = max(cores, 8) if args.ref_ntree <= 0 else args.ref_ntree
    times = []
    for it in range(args.warmup + args.steps):
        _, _, dt = cpu_reference_sample(cfg, sample_n=sample_n, ntree=ntree, threads=-1)
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = ntree / (ms / 1e3)
    sample = (f"unmodified reference (randomForestSRC.c + splitCustom.c, gcc -O2 -fopenmp) on a bounded sample of the workload: "
              f"n={sample_n} of {cfg['n']} CPIUs x p={cfg['p']}, L={cfg['levels']} levels, ntree={ntree} (>= 1 tree per core), "
              f"nodesize={cfg['nodesize']}, poisson.split{cfg['rule']}; the reference allocates ~64*n*ntree bytes and is "
              f"O(m * #distinct) per covariate, so the full n=1e6 x ntree=500 point is out of its reach (BASELINE.md 3); "
              f"the GPU arm's cpu_baseline.same_job times the device path on exactly this job")
    line = {
        "impl": "reference", "metric": "trees_per_sec", "value": value, "unit": "trees/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(cfg), "sample_n": sample_n, "sample_ntree": ntree},
        "cpu_baseline": {"value": value, "unit": "trees/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def ncu_traffic(name):
    """DRAM bytes of one tree-kernel launch of a workload from the committed `ncu --set full` capture of it
    (profiles/ncu_traffic_r02.json); nothing is measured under a profiler here."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")))
        return float(t[name]["dram_bytes_per_launch"])
    except Exception:
        return None


def workload_name(cfg):
    return (f"BASELINE configs[1]: {cfg['n']} CPIUs x {cfg['p']} covariates (mode Q, {cfg['levels']} levels), ntree={cfg['ntree']} per GPU, "
            f"mtry=ceil(sqrt(p)), nodesize={cfg['nodesize']}, poisson.split{cfg['rule']}, bootstrap by.root, seed {cfg['seed']}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="override rows (debug; not --n: torchrun abbreviates it to --nnodes)")
    ap.add_argument("--trees", type=int, default=0, help="override trees per GPU (debug)")
    ap.add_argument("--ref-sample-n", type=int, default=REF_SAMPLE_N)
    ap.add_argument("--ref-ntree", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the configs[2] record (10M x 100, 1000 trees)")
    ap.add_argument("--c3-rows", type=int, default=0)
    ap.add_argument("--c3-trees", type=int, default=0)
    ap.add_argument("--no-predict", action="store_true", help="skip the rfsrcPredict record")
    ap.add_argument("--no-extras", action="store_true", help="skip the mixed-factor and CPIU-construction records")
    ap.add_argument("--predict-rows", type=int, default=2_500_000, help="held-out rows PER GPU for the predict record (configs[4]: 20 M rows over 8 GPUs)")
    ap.add_argument("--predict-trees", type=int, default=1000, help="trees of the forest the predict record routes (configs[4]: 1000)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import rfslam_b200
    from rfslam_b200 import capi, shard
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(WORKLOAD)
    if args.rows:
        cfg["n"] = args.rows
    if args.trees:
        cfg["ntree"] = args.trees
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rfslam_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = shard.Comm.from_torch(local)                      # the library's own NCCL communicator (csrc/comm.cu)

    d = make_data(cfg)
    n, p = d.n, d.p
    forest_ntree = cfg["ntree"] * world                      # weak scaling: every rank grows cfg.ntree trees
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat, device=local)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(vals):
        if dist is None:
            return [float(v) for v in vals]
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(vals):
        if dist is None:
            return [float(v) for v in vals]
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def one_step():
        # N > 1: the all-reduce of the ensemble sums and the gather of the forest records to rank 0 happen inside
        # the call, on device buffers (comm); every rank returns the whole forest's ensembles, rank 0 its records
        return ds.grow(splitrule=cfg["rule"], ntree=forest_ntree, nodesize=cfg["nodesize"], seed=cfg["seed"], k_for_alpha=cfg["k_for_alpha"],
                       tree_first=rank + 1, tree_step=world, finalize=True, comm=comm, gather_forest=True)

    def timed_resident(count, sample_clocks=False):
        dev_ms, kern_ms, walls = [], [], []
        last = None
        # one sampler for the whole job (rank 0, all of the job's GPUs): N polling nvidia-smi processes contend on the driver
        sampler = ClockSampler(",".join(str(i) for i in range(world))) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        for _ in range(count):
            last = None                                      # drop the previous forest: its page-locked blocks are reused
            flush.fill_(1)
            barrier()
            t0 = time.perf_counter()
            last = one_step()
            barrier()
            walls.append(time.perf_counter() - t0)
            dev_ms.append(last["total_device_ms"]); kern_ms.append(last["grow_kernel_ms"])
        clocks = sampler.stop() if sampler else None
        return walls, dev_ms, kern_ms, last, clocks

    timed_resident(args.warmup)
    walls, dev_ms, kern_ms, f, clocks = timed_resident(args.steps, sample_clocks=True)
    step_s, kern_mean, dev_mean, memb_ms, boot_ms, coll_ms = max_over_ranks(
        [float(np.mean(walls)), float(np.mean(kern_ms)), float(np.mean(dev_ms)), f["membership_kernel_ms"], f["bootstrap_ms"], f["collective_ms"]])
    cand_total, algo_grow, algo_memb, algo_boot = sum_over_ranks(
        [f["split_candidates"], f["grow_kernel_algorithmic_bytes"], f["membership_algorithmic_bytes"], f["bootstrap_algorithmic_bytes"]])
    nodes_total = float(f["treeID"].size) if world == 1 else sum_over_ranks([float(f["treeID"].size)])[0]     # rank 0 holds the gathered forest
    launches_step = int(f["kernel_launches"])

    # ---- e2e: the same forest through the public host-buffer call (upload + rank-code + grow + exchange + download).
    # A device fault here is a failed benchmark: nothing is caught.
    def e2e_step():
        return rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=cfg["rule"], ntree=forest_ntree, nodesize=cfg["nodesize"], seed=cfg["seed"],
                                 k_for_alpha=cfg["k_for_alpha"], device=local, tree_first=rank + 1, tree_step=world, finalize=True,
                                 comm=comm, gather_forest=True)
    h2d = d.x.nbytes + d.y.nbytes + d.rt.nbytes + d.strat.nbytes
    e2e_times = []
    fe = None
    for it in range(2 + max(5, min(args.steps, 10))):        # two untimed calls: page-locked output blocks and device scratch get recycled from the third on
        fe = None                                            # the caller drops the previous forest before asking for the next
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        fe = e2e_step()
        barrier()
        if it > 1:
            e2e_times.append(time.perf_counter() - t0)
    e2e_s = max_over_ranks([float(np.mean(e2e_times))])[0]
    d2h = max_over_ranks([float(sum(int(v.nbytes) for v in fe.values() if isinstance(v, np.ndarray)))])[0]
    fe = None

    trees_total = cfg["ntree"] * world
    peak, peak_src = measured_peak_gbs()
    achieved = (algo_grow / world) / 1e9 / (kern_mean / 1e3)        # per GPU, the tree kernel of one launch: its own bytes / its own time
    line = {
        "metric": "trees_per_sec", "value": trees_total / step_s, "unit": "trees/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "n": n, "p": p, "ntree_total": trees_total, "l2": "256 MB flush buffer written between timed steps",
                   "parallelism": f"tree-sharded x{world}" + ("; NCCL all-reduce of the ensemble sums + gather of the forest records to rank 0 inside the step" if world > 1 else "")},
        "step_ms": [1e3 * w for w in walls],                        # rank 0's wall clock of every timed step
        "split_candidates_per_sec": cand_total / step_s,
        "device_ms_per_step": dev_mean, "grow_kernel_ms_per_step": kern_mean, "forest_nodes": nodes_total,
        "clocks": clocks,
        "e2e": {"value": trees_total / e2e_s, "unit": "trees/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s, "timed_calls": len(e2e_times)},
        "gpu_launches": launches_step * args.steps,
        "roofline": {"bound": "hbm", "kernel": "rfslam_grow_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_grow / world, "kernel_ms": kern_mean,
                     "traffic": ncu_traffic("configs1"), "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                     "other_kernels": [
                         {"kernel": "membership_kernel", "algorithmic_bytes_per_step": algo_memb / world, "ms": memb_ms,
                          "achieved": (algo_memb / world) / 1e9 / max(memb_ms / 1e3, 1e-9)},
                         {"kernel": "bootstrap_draw_kernel + bag_count_kernel", "algorithmic_bytes_per_step": algo_boot / world, "ms": boot_ms,
                          "achieved": (algo_boot / world) / 1e9 / max(boot_ms / 1e3, 1e-9)}]},
    }
    if comm is not None:
        line["comm"] = dict(comm.collectives(), collective_ms_per_step=coll_ms, nccl_version=int(capi.lib().rfslam_b200_nccl_version()))

    # ---- predict record: rfsrcPredict of the step's forest, held-out rows sharded over the ranks ----
    if not args.no_predict:
        from rfslam_b200.synth import make_cpiu_torch
        # forest replicated: at N > 1 every rank grows the same cfg.ntree-tree forest for this record (trees are
        # deterministic, so the replicas are identical); the held-out rows are what is sharded
        # BASELINE configs[4]: rfsrcPredict + aggregation of a 1000-tree forest on 20 M held-out CPIUs across 8 GPUs = 2.5 M
        # rows per GPU; the record keeps that per-GPU share at every N (at N = 8 it is the configuration itself)
        fp = ds.grow(splitrule=cfg["rule"], ntree=args.predict_trees, nodesize=cfg["nodesize"], seed=cfg["seed"], k_for_alpha=cfg["k_for_alpha"])
        m_total = args.predict_rows * world
        fx = make_cpiu_torch(args.predict_rows, cfg["p"], levels=cfg["levels"], seed=20260102 + rank, device=f"cuda:{local}").x
        lo, cnt = 0, args.predict_rows                           # every rank holds (only) its own shard of the held-out rows
        pr = None
        ptimes = []
        for it in range(3):
            barrier()
            t0 = time.perf_counter()
            pr = rfslam_b200.predict_rfsrc(fp, fx, membership=False, device=local, frow_first=lo, frow_count=cnt)
            barrier()
            ptimes.append(time.perf_counter() - t0)
        p_s, p_kern = max_over_ranks([float(np.mean(ptimes[1:])), pr["kernel_ms"]])
        visits, rowtrees = sum_over_ranks([pr["node_visits"], pr["row_tree_visits"]])
        pbytes = 24.0 * visits + 8.0 * rowtrees                  # B_pred = d * (8 + 16) + 8 per (row, tree)
        line["predict"] = {"metric": "row_trees_per_sec", "value": rowtrees / p_s, "rows": m_total, "trees": int(fp["leafCount"].size),
                           "ms_per_call": 1e3 * p_s, "route_kernel_ms": p_kern, "mean_depth": visits / max(rowtrees, 1.0),
                           "h2d_bytes_per_call": int(fx.nbytes + sum(int(fp[k].nbytes) for k in ("treeID", "nodeID", "parmID", "contPT", "tnCLAS"))),
                           "roofline": {"bound": "hbm", "kernel": "route_kernel", "achieved": pbytes / world / 1e9 / max(p_kern / 1e3, 1e-9), "peak": peak,
                                        "unit": "GB/s", "frac": pbytes / world / 1e9 / max(p_kern / 1e3, 1e-9) / peak, "algorithmic_bytes_per_call": pbytes / world},
                           "sharding": f"{args.predict_rows} held-out rows per rank over {world} rank(s) (configs[4]'s per-GPU share), forest replicated, no collective"}
        del fx, pr, fp

    # ---- the rows SURVEY 8 marks "next", measured beside the path (rank 0, N = 1; not part of the timed step) ----
    if rank == 0 and world == 1 and not args.no_extras:
        # (f2) mixed numeric / unordered-factor table -- configs[3]'s covariate kinds on configs[1]'s shape: every fifth
        # covariate becomes a factor with 3..8 levels (MWCP splits, level-subset cuts)
        rng = np.random.default_rng(4)
        xm = d.x.copy(); xt = ["R"] * p; xl = np.zeros(p, np.int32)
        for c in range(2, p, 5):
            lv = 3 + (c // 5) % 6
            xm[c] = rng.integers(1, lv + 1, size=n).astype(np.float64); xt[c] = "C"; xl[c] = lv
        dsm = rfslam_b200.ResidentCpiu(xm, d.y, d.rt, d.strat, device=local, xvar_types=xt, xvar_levels=xl)
        mixed_trees = 296                                                # one full wave of the 256-thread build
        fm = None
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            fm = dsm.grow(splitrule=cfg["rule"], ntree=mixed_trees, nodesize=cfg["nodesize"], seed=cfg["seed"], k_for_alpha=cfg["k_for_alpha"])
            torch.cuda.synchronize(); mixed_s = time.perf_counter() - t0
        line["mixed_factor"] = {"workload": f"{n} CPIUs x {p} covariates, {sum(t == 'C' for t in xt)} of them unordered factors with 3..8 levels, ntree={mixed_trees}, nodesize={cfg['nodesize']}",
                                "value": mixed_trees / mixed_s, "unit": "trees/s", "ms_per_step": 1e3 * mixed_s, "grow_kernel_ms": fm["grow_kernel_ms"],
                                "factor_splits": int(fm["mwcpSZ"].sum()), "forest_nodes": int(fm["treeID"].size),
                                "split_candidates_per_sec": int(fm["split_candidates"]) / mixed_s}
        dsm.close(); del dsm, xm, fm
        # (f3) permutation importance on configs[1]'s table: one wave of trees, every covariate (importance = TRUE); the
        # reference's permute() is O(m^2) per (tree, covariate) -- 6.8e10 slot visits at these 368 k OOB rows -- so there is
        # no CPU figure beside it at this size (parity: tests/test_gpu_parity.py at sizes the reference finishes)
        vimp_trees = 148
        t0 = time.perf_counter()
        fv = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=cfg["rule"], ntree=vimp_trees, nodesize=cfg["nodesize"],
                               seed=cfg["seed"], k_for_alpha=cfg["k_for_alpha"], device=local, importance="permute")
        vimp_s = time.perf_counter() - t0
        line["importance"] = {"workload": f"{n} CPIUs x {p} covariates, ntree={vimp_trees}, per-tree permutation importance of all {p} covariates",
                              "vimp_device_ms": fv["vimp_ms"], "call_s": vimp_s, "permutations": vimp_trees * p,
                              "value": vimp_trees * p / max(fv["vimp_ms"] / 1e3, 1e-9), "unit": "(tree, covariate) permutations/s",
                              "oob_row_walks_per_s": vimp_trees * (p + 1) * 0.3679 * n / max(fv["vimp_ms"] / 1e3, 1e-9),
                              "top5_covariates": [int(i) + 1 for i in (-fv["importance"][:, 0]).argsort()[:5]]}
        del fv
        # (f1) CPIU construction: persons -> counting-process information units with derived time-varying covariates
        persons = 1_000_000
        t_out = np.round(rng.uniform(30, 2900, persons), 1); dth = (rng.random(persons) < 0.3).astype(np.int32)
        t_dth = np.where(dth == 1, t_out, t_out + 100.0)
        ev = rng.uniform(0, 3000, (persons, 10)); ev[rng.random((persons, 10)) < 0.6] = np.nan
        mt = np.sort(rng.uniform(0, 2500, (persons, 4)), axis=1); mt[:, 0] = 0.0; mv = rng.normal(50, 10, (persons, 4))
        to_create = {"i.hf": ev, "n.hf": ev, "ni.hf": ev, "t.hf": ev, "i.sca": ev[:, :1], "t.sca": ev[:, :1], "c.lab": {"times": mt, "values": mv}}
        cres = None; ctimes = []
        for it in range(3):
            t0 = time.perf_counter()
            cres = rfslam_b200.cpiu_fc(t_out, t_dth, dth, to_create=to_create, k_to_persist=2, device=local)
            ctimes.append(time.perf_counter() - t0)
        rows = int(cres["rt"].size); nv = len(to_create)
        cbytes = rows * (6 * 8 + nv * 8)                                 # bytes written per CPIU: six interval columns + the derived ones
        from oracle import cpiu_port
        sub = 2000
        kinds = [0, 1, 2, 3, 0, 3, 4]
        t0 = time.perf_counter()
        cpiu_port.cpiu_fc(t_out[:sub], t_dth[:sub], dth[:sub], var_kind=kinds, var_times=[ev[:sub]] * 4 + [ev[:sub, :1]] * 2 + [mt[:sub]],
                          var_values=[None] * 6 + [mv[:sub]], k_to_persist=2)
        cpu_s = time.perf_counter() - t0
        line["cpiu"] = {"workload": f"{persons} persons, half-year intervals, {nv} derived covariates (i./n./ni./t./c., up to 10 event slots)",
                        "value": rows / float(np.min(ctimes[1:])), "unit": "CPIUs/s (host buffers in, host columns out)", "rows": rows,
                        "kernel_ms": cres["kernel_ms"], "ms_per_call": 1e3 * float(np.min(ctimes[1:])),
                        "roofline": {"bound": "hbm", "kernel": "cpiu_vars_kernel + cpiu_base_kernel", "achieved": cbytes / 1e9 / max(cres["kernel_ms"] / 1e3, 1e-9),
                                     "peak": peak, "unit": "GB/s", "frac": cbytes / 1e9 / max(cres["kernel_ms"] / 1e3, 1e-9) / peak, "algorithmic_bytes_per_call": cbytes},
                        "cpu_baseline": {"value": float(int(np.ceil(t_out[:sub] / 182.5).sum()) / cpu_s), "unit": "CPIUs/s", "cores": 1, "kind": "reference",
                                         "sample": f"the reference's cpiu.cpp helpers (compiled unmodified) under the restated R driver, {sub} persons"}}
        del cres, ev, mt, mv

    # ---- cpu baseline (rank 0, N = 1): the unmodified reference on a bounded sample, and the SAME job on the device ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tests import parity
        from oracle import refdriver
        cores = os.cpu_count() or 1
        nt = max(cores, 8)
        sd, ref, dt = cpu_reference_sample(cfg, sample_n=args.ref_sample_n, ntree=nt, threads=-1, membership=True)
        same = None
        gts = []
        for it in range(3):
            t0 = time.perf_counter()
            same = rfslam_b200.rfsrc(sd.x, sd.y, sd.rt, sd.strat, splitrule=cfg["rule"], ntree=nt, nodesize=cfg["nodesize"], seed=cfg["seed"],
                                     k_for_alpha=cfg["k_for_alpha"], membership=True, split_stat=True)
            gts.append(time.perf_counter() - t0)
        # outside every timed region: the two forests of this sample must be the same forest
        rt = refdriver.trim_forest(ref, sd.n)
        order = np.argsort(rt["treeID"], kind="stable")
        for k in ("treeID", "nodeID", "parmID", "contPT", "mwcpSZ"):
            rt[k] = rt[k][order]
        from oracle import port
        chk = parity.assert_forest_equal(same, rt, sd.n, nt, label="bench sample", tie_aware=True,
                                         ref_stat=lambda: port.grow(sd.x, sd.y, sd.rt, sd.strat, rule=cfg["rule"], ntree=nt, nodesize=cfg["nodesize"], seed=cfg["seed"]))
        g_s = float(np.min(gts[1:]))
        line["cpu_baseline"] = {
            "value": nt / dt, "unit": "trees/s", "cores": cores, "kind": "reference",
            "split_candidates_per_sec": int(same["split_candidates"]) / dt,
            "sample": (f"unmodified reference C (OpenMP, {cores} host threads) on n={args.ref_sample_n} of {cfg['n']} CPIUs x p={cfg['p']}, "
                       f"L={cfg['levels']}, ntree={nt}, nodesize={cfg['nodesize']}, poisson.split{cfg['rule']}: {dt:.1f} s"),
            "same_job": {"what": "the device path through rfslam_b200.rfsrc() from host buffers on exactly this sample and ntree",
                         "value": nt / g_s, "unit": "trees/s", "ms": 1e3 * g_s, "ratio": (nt / g_s) / (nt / dt),
                         "forest_identical": True, "ties": chk["ties"], "trees_compared": chk["trees_compared"]}}
        del sd, ref, same

    # ---- c3: BASELINE configs[2], ONE forest of C3.ntree trees sharded over the ranks (strong scaling) ----
    if not args.no_c3:
        from rfslam_b200.synth import make_cpiu_torch
        c3 = dict(C3)
        if args.c3_rows:
            c3["n"] = args.c3_rows
        if args.c3_trees:
            c3["ntree"] = args.c3_trees
        ds.close()
        del ds, d, f
        flush = None
        torch.cuda.empty_cache()
        capi.lib().rfslam_b200_host_pool_trim()
        d3 = make_cpiu_torch(c3["n"], c3["p"], levels=c3["levels"], device=f"cuda:{local}")
        torch.cuda.empty_cache()
        ds3 = rfslam_b200.ResidentCpiu(d3.x, d3.y, d3.rt, d3.strat, device=local)
        barrier()
        t0 = time.perf_counter()
        f3 = ds3.grow(splitrule=c3["rule"], ntree=c3["ntree"], nodesize=c3["nodesize"], seed=c3["seed"], k_for_alpha=c3["k_for_alpha"],
                      tree_first=rank + 1, tree_step=world, finalize=True, comm=comm, gather_forest=False)
        barrier()
        c3_s = time.perf_counter() - t0
        c3_s, c3_kern = max_over_ranks([c3_s, f3["grow_kernel_ms"]])
        c3_algo, c3_cand, c3_nodes = sum_over_ranks([f3["grow_kernel_algorithmic_bytes"], f3["split_candidates"], float(f3["treeID"].size)])
        a3 = (c3_algo / world) / 1e9 / (c3_kern / 1e3)
        line["c3"] = {"workload": (f"BASELINE configs[2]: {c3['n']} CPIUs x {c3['p']} covariates (mode Q, {c3['levels']} levels, generated on the device), "
                                   f"ONE forest of ntree={c3['ntree']} sharded over {world} GPU(s), mtry=ceil(sqrt(p))=10, nodesize={c3['nodesize']}, poisson.split{c3['rule']}"),
                      "scaling": "strong", "trees": c3["ntree"], "trees_per_gpu": len(range(rank + 1, c3["ntree"] + 1, world)), "steps": 1, "warmup": 0,
                      "value": c3["ntree"] / c3_s, "unit": "trees/s", "ms_per_step": 1e3 * c3_s, "grow_kernel_ms": c3_kern,
                      # where this rank's step went: the tree kernel, the other device work (bootstrap replay, OOB / full-ensemble
                      # traversal, compaction), and the host side -- a COLD step: ~20 GB of forest records come home into page-locked
                      # blocks that are allocated here for the first time (a second step of the same shape reuses them: 0.1 s)
                      "device_ms": f3["total_device_ms"], "bootstrap_ms": f3["bootstrap_ms"], "membership_kernel_ms": f3["membership_kernel_ms"],
                      "host_side_ms": 1e3 * c3_s - f3["total_device_ms"],
                      "split_candidates_per_sec": c3_cand / c3_s, "forest_nodes": c3_nodes, "table_bytes_hbm": int(ds3.hbm_bytes),
                      "roofline": {"bound": "hbm", "kernel": "rfslam_grow_kernel", "achieved": a3, "peak": peak, "unit": "GB/s", "frac": a3 / peak,
                                   "algorithmic_bytes_per_launch": c3_algo / world, "traffic": ncu_traffic("configs2"),
                                   "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum)"}}
        line["gpu_launches"] += int(f3["kernel_launches"])
        ds3.close()

    if rank == 0:
        print(json.dumps(line))
    sys.stdout.flush(); sys.stderr.flush()
    if comm is not None:
        comm.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    # stdout carries exactly ONE line (the JSON record): anything a library writes to fd 1 meanwhile -- NCCL prints its
    # version there when the first communicator comes up -- is sent to stderr
    sys.stdout.flush()
    _out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _out
    rc = main()
    _out.flush()
    sys.exit(rc)
