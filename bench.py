#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (trees/sec; split candidates scored/sec; % of HBM peak) on
BASELINE.json's configuration, for the B200-native path and, with --impl reference, for the
reference's own CPU implementation on the box's host cores.

  python bench.py --gpus N --steps K --warmup W            (N = 1: plain python)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

One step = one pass of the hot path over the workload: grow the whole forest (bootstrap replay,
split scoring, partition, leaf estimates, OOB / full ensembles) from the CPIU table resident in HBM.
Workload at N = 1 is BASELINE.json configs[1]: 1M CPIUs x 50 covariates, ntree = 500,
mtry = ceil(sqrt(p)) = 8, nodesize = 20, rule poisson.split4, cardinality mode Q (64 levels).
At N > 1 every rank holds a replica of the table and grows 500 trees of a 500*N-tree forest
(weak scaling; trees are independent, no collective during growth); one NCCL all-reduce of the
ensemble sums and a gather of the forest sizes close the step.
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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
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
                "reasons": sorted(reasons), "samples": len(sm)}


def make_data(cfg):
    from rfslam_b200.synth import make_cpiu
    return make_cpiu(cfg["n"], cfg["p"], levels=cfg["levels"])


def cpu_reference_sample(cfg, *, sample_n: int, ntree: int, threads: int):
    """The reference's own OpenMP CPU path (oracle/_ref, the unmodified reference compiled against
    the fake-R shim) on a bounded sample of the same synthetic workload."""
    from oracle import refdriver
    from rfslam_b200.synth import make_cpiu
    d = make_cpiu(sample_n, cfg["p"], levels=cfg["levels"])
    t0 = time.perf_counter()
    r = refdriver.grow(d.x, d.y, d.rt, d.strat, rule=cfg["rule"], ntree=ntree, nodesize=cfg["nodesize"], seed=cfg["seed"],
                       k_for_alpha=cfg["k_for_alpha"], membership=False, nthreads=threads)
    dt = time.perf_counter() - t0
    return d, r, dt


def count_candidates_gpu(d, cfg, ntree):
    """Split candidates of the sample forest: identical trees => identical count, taken from the device path."""
    import rfslam_b200
    f = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=cfg["rule"], ntree=ntree, nodesize=cfg["nodesize"], seed=cfg["seed"],
                          k_for_alpha=cfg["k_for_alpha"])
    return int(f["split_candidates"])


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
              f"O(m * #distinct) per covariate, so the full n=1e6 x ntree=500 point is out of its reach (BASELINE.md 3)")
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


def ncu_traffic(cfg):
    """DRAM bytes of one tree-kernel launch from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic_r01.json); None for any other workload (nothing is measured under a profiler here)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")))
        wl = t["workload"]
        same = all(int(wl[k]) == int(cfg[k]) for k in ("n", "p", "ntree", "levels", "nodesize", "rule"))
        return float(t["dram_bytes_per_launch"]) if same else None
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
    ap.add_argument("--ref-sample-n", type=int, default=100_000)
    ap.add_argument("--ref-ntree", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-only", action="store_true", help="internal: measure only the host-buffer leg and print {e2e_s, h2d, d2h}")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import rfslam_b200
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    d = make_data(cfg)
    n, p = d.n, d.p
    forest_ntree = cfg["ntree"] * world                      # weak scaling: every rank grows cfg.ntree trees
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    if args.e2e_only:                                        # child of a failed in-process e2e leg (see below)
        times = []
        fe = None
        for it in range(2 + max(1, min(args.steps, 2))):
            fe = None
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fe = rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=cfg["rule"], ntree=forest_ntree, nodesize=cfg["nodesize"], seed=cfg["seed"],
                                   k_for_alpha=cfg["k_for_alpha"], device=local)
            torch.cuda.synchronize()
            if it > 1:
                times.append(time.perf_counter() - t0)
        print(json.dumps({"e2e_s": float(np.mean(times)), "h2d": int(d.x.nbytes + d.y.nbytes + d.rt.nbytes + d.strat.nbytes),
                          "d2h": int(sum(int(v.nbytes) for v in fe.values() if isinstance(v, np.ndarray)))}))
        return 0
    ds = rfslam_b200.ResidentCpiu(d.x, d.y, d.rt, d.strat, device=local)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def one_step():
        f = ds.grow(splitrule=cfg["rule"], ntree=forest_ntree, nodesize=cfg["nodesize"], seed=cfg["seed"], k_for_alpha=cfg["k_for_alpha"],
                    tree_first=rank + 1, tree_step=world, finalize=(world == 1))
        if dist is not None:
            # the one exchange of the path: sum the ensemble numerators / denominators over ranks
            # (rf.c:944-949) and tell every rank the shards' forest sizes; the records stay sharded by tree
            from rfslam_b200 import shard
            f = shard.allreduce_ensembles(f, device="cuda")
            sizes = torch.tensor([f["treeID"].size, int(f["leafCount"].sum())], device="cuda")
            gathered = [torch.zeros_like(sizes) for _ in range(world)]
            dist.all_gather(gathered, sizes)
            torch.cuda.synchronize()
        return f

    def timed_resident(count, sample_clocks=False):
        dev_ms, kern_ms, walls = [], [], []
        last = None
        # one sampler for the whole job (rank 0, all of the job's GPUs): N polling nvidia-smi processes contend on the driver
        sampler = ClockSampler(",".join(str(i) for i in range(world))) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        for _ in range(count):
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
    step_s = float(np.mean(walls))
    if dist is not None:
        t = torch.tensor([step_s, float(np.mean(kern_ms)), float(np.mean(dev_ms))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_s, kern_mean, dev_mean = [float(v) for v in t.tolist()]
        tot = torch.tensor([f["split_candidates"], f["algorithmic_bytes"], f["treeID"].size], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot)
        cand_total, algo_total, nodes_total = [float(v) for v in tot.tolist()]
    else:
        kern_mean, dev_mean = float(np.mean(kern_ms)), float(np.mean(dev_ms))
        cand_total, algo_total, nodes_total = float(f["split_candidates"]), float(f["algorithmic_bytes"]), float(f["treeID"].size)

    # ---- e2e: the same forest through the public host-buffer call (upload + rank-code + grow + download) ----
    def e2e_step():
        return rfslam_b200.rfsrc(d.x, d.y, d.rt, d.strat, splitrule=cfg["rule"], ntree=forest_ntree, nodesize=cfg["nodesize"], seed=cfg["seed"],
                                 k_for_alpha=cfg["k_for_alpha"], device=local, tree_first=rank + 1, tree_step=world, finalize=(world == 1))
    # No collective inside or after this leg: a device fault in one rank must not hang the others (the
    # one collective of the path, the ensemble all-reduce, is part of the resident step above).  Ranks
    # start together, time their own calls, and hand rank 0 their mean through a file; the job's e2e
    # time is the max over ranks.
    e2e_note = None
    h2d = d.x.nbytes + d.y.nbytes + d.rt.nbytes + d.strat.nbytes
    d2h = 0
    barrier()
    try:
        e2e_times = []
        fe = None
        for it in range(2 + max(1, min(args.steps, 2))):         # two untimed calls: page-locked output blocks and device scratch get recycled from the third on
            fe = None                                            # the caller drops the previous forest before asking for the next
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fe = e2e_step()
            torch.cuda.synchronize()
            if it > 1:
                e2e_times.append(time.perf_counter() - t0)
        e2e_s = float(np.mean(e2e_times))
        d2h = sum(int(v.nbytes) for v in fe.values() if isinstance(v, np.ndarray))
    except Exception as e:                                       # a CUDA fault poisons this process
        e2e_s = None
        e2e_note = f"in-process e2e leg failed on rank {rank} ({type(e).__name__}: {str(e)[:160]})"
        sys.stderr.write("[bench] " + e2e_note + "\n")
        if world == 1:                                           # measure the leg in a fresh process
            cmd = [sys.executable, os.path.abspath(__file__), "--e2e-only", "--steps", str(args.steps), "--warmup", str(args.warmup), "--no-cpu-baseline"]
            if args.rows: cmd += ["--rows", str(args.rows)]
            if args.trees: cmd += ["--trees", str(args.trees)]
            outp = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            child = json.loads(outp.stdout.strip().split("\n")[-1])
            e2e_s, h2d, d2h = child["e2e_s"], child["h2d"], child["d2h"]
            e2e_note += "; measured in a fresh process"
    if world > 1:
        tag = os.environ.get("MASTER_PORT", "0") + "_" + os.environ.get("TORCHELASTIC_RUN_ID", "x")
        mine = os.path.join("/tmp", f"rfslam_bench_e2e_{tag}_{rank}.json")
        with open(mine + ".tmp", "w") as fh:
            json.dump({"e2e_s": e2e_s, "d2h": int(d2h), "note": e2e_note}, fh)
        os.replace(mine + ".tmp", mine)
        if rank == 0:
            got, deadline = {}, time.time() + 900
            while len(got) < world and time.time() < deadline:
                for r in range(world):
                    fn = os.path.join("/tmp", f"rfslam_bench_e2e_{tag}_{r}.json")
                    if r not in got and os.path.exists(fn):
                        got[r] = json.load(open(fn)); os.remove(fn)
                time.sleep(0.05)
            ok = [v["e2e_s"] for v in got.values() if v.get("e2e_s")]
            notes = [v["note"] for v in got.values() if v.get("note")] + ([f"{world - len(got)} rank(s) did not report"] if len(got) < world else [])
            e2e_s = max(ok) if ok else None
            d2h = max([v["d2h"] for v in got.values()] + [0])
            e2e_note = "; ".join(notes) if notes else None

    trees_total = cfg["ntree"] * world
    peak, peak_src = measured_peak_gbs()
    achieved = (algo_total / world) / 1e9 / (kern_mean / 1e3)        # per GPU, the tree kernel of one launch
    line = {
        "metric": "trees_per_sec", "value": trees_total / step_s, "unit": "trees/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "n": n, "p": p, "ntree_total": trees_total, "l2": "256 MB flush buffer written between timed steps",
                   "parallelism": f"tree-sharded x{world}"},
        "split_candidates_per_sec": cand_total / step_s,
        "device_ms_per_step": dev_mean, "grow_kernel_ms_per_step": kern_mean, "forest_nodes": nodes_total,
        "clocks": clocks,
        "e2e": dict({"value": (trees_total / e2e_s) if e2e_s else None, "unit": "trees/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                     "ms_per_step": (1e3 * e2e_s) if e2e_s else None},
                    **({"note": e2e_note} if e2e_note else {})),
        "gpu_launches": int(f["kernel_launches"]) * args.steps,
        "roofline": {"bound": "hbm", "kernel": "rfslam_grow_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_total / world, "kernel_ms": kern_mean,
                     "traffic": ncu_traffic(cfg), "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum)"},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            nt = max(cores, 8)
            sd, _, dt = cpu_reference_sample(cfg, sample_n=args.ref_sample_n, ntree=nt, threads=-1)
            cands = count_candidates_gpu(sd, cfg, nt)
            line["cpu_baseline"] = {
                "value": nt / dt, "unit": "trees/s", "cores": cores, "kind": "reference",
                "split_candidates_per_sec": cands / dt,
                "sample": (f"unmodified reference C (OpenMP, {cores} host threads) on n={args.ref_sample_n} of {cfg['n']} CPIUs x p={cfg['p']}, "
                           f"L={cfg['levels']}, ntree={nt}, nodesize={cfg['nodesize']}, poisson.split{cfg['rule']}: {dt:.1f} s")}
        except Exception as e:   # the baseline is a reported figure, never the product path
            line["cpu_baseline"] = {"value": None, "unit": "trees/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line))
    sys.stdout.flush(); sys.stderr.flush()
    if dist is not None or e2e_note:
        os._exit(0)              # no collective teardown: a rank whose context faulted in the e2e leg must not hang the rest
    ds.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
