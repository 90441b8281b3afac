#!/usr/bin/env python
"""bench.py — diploid samples/second of the QUILT2 per-sample Gibbs hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" = every Gibbs call of one batch of synthetic samples at QUILT2 defaults: per sample 8 chains x
(3 calls on the common SNPs + 1 call on all SNPs) = 32 calls, each 20 burn-in + 1 sampling sweep with shard
passes after sweeps 3/6/9, AND the haplotype re-selection between the calls of a chain
(select_new_haps_mspbwt_v3, functions.R:856-868): the calls are staged as four batches (call 1 / 2 / 3 of every chain,
then the all-SNP calls) and each batch receives its which_haps_to_use from the previous one ON THE DEVICE
(quilt_gpu_batch_chain_select) — hapProbs_t of the intermediate calls never travel to the host.  Workload at N = 1 = the configuration the metric is quoted on:
chr20 2 Mb (+2x0.5 Mb buffer) at 1x, K = 4096 of a 5008-haplotype panel, 32 000 common / 96 000 total SNPs.
Samples shard across ranks (weak scaling: the per-GPU batch is fixed); the only collectives are the one-time
NCCL broadcast of the prepared reference and the max-reduction of the timed region.

value      device-timed (CUDA events on the library stream), inputs resident in HBM before the timed region
e2e        the same batch through quilt_gpu_gibbs_batch with HOST buffers: host preparation, H2D, kernels, D2H
roofline   the sweep kernel: algorithmic bytes (8 K (5 nHap T + R) per job and sweep) / event time
parity_gate  before anything is timed: one whole sample's calls (32 at the headline workload) through the CPU checker on
           the host threads, compared with the staged batch's fetched outputs; any label / GT mismatch or
           max |dDS|, |dGP| > 1e-4 voids the run (non-zero exit)
cpu_baseline / --impl reference
           the reference's OWN C++ for this path (oracle/_ref/libquiltref.so: the unmodified QUILT/src sources compiled
           against the header-only RcppArmadillo stand-in, kind "reference"), one call per thread on all host cores
           like the reference's one-forked-worker-per-core model (QUILT/R/quilt.R:690-692); if that library is absent
           the statement-order oracle port is timed instead (kind "port")
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: K, K_full, nSNPs (common), all-SNP factor, region bp, coverage, samples per GPU per step
    "chr20_2Mb_1x_K4096": dict(K=4096, K_full=5008, nSNPs=32000, factor=3, region_bp=3_000_000, coverage=1.0, samples=37),
    "chr20_2Mb_1x_K512": dict(K=512, K_full=5008, nSNPs=32000, factor=3, region_bp=3_000_000, coverage=1.0, samples=64),
    # BASELINE.json config #4 (per-GPU share): NIPT, three haplotypes, fetal fraction 10 %, 0.5x, K = 2048; common-SNP calls only
    "nipt_2Mb_0.5x_K2048": dict(K=2048, K_full=5008, nSNPs=32000, factor=0, region_bp=3_000_000, coverage=0.5, samples=48, ff=0.1),
    # BASELINE.json config #5 (per-GPU share, scaled): K = 8192 (two-CTA cluster kernels) drawn from a 20 000-haplotype synthetic panel
    # (a 200 000-haplotype panel needs 6.4 GB of NumPy bits per rank just to be synthesised); common-SNP calls only
    "chr20_2Mb_1x_K8192_20k": dict(K=8192, K_full=20000, nSNPs=32000, factor=0, region_bp=3_000_000, coverage=1.0, samples=37),
    "tiny": dict(K=256, K_full=600, nSNPs=3200, factor=3, region_bp=300_000, coverage=1.0, samples=4),
}
METRIC = "diploid samples/sec, chr20 2Mb @1x cov, K=4096, 5008-hap panel; 1/2/4/8 GPU"
WORLD_SEED = 20260118


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per sweep launch of THIS workload from the committed ncu captures (profiles/sweep_traffic.json), or None when no
    capture of the workload exists (a number taken on another shape would be meaningless)"""
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as fh:
            return json.load(fh)["workloads"].get(workload)
    except Exception:
        return None


class ClockSampler:
    QUERY = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(wl, rank, world_obj, log):
    from quilt_b200 import schedule, synth

    t0 = time.time()
    calls = []
    n_samples = wl["samples"]
    for s in range(n_samples):
        seed = 1000 * (rank + 1) + s
        ff = wl.get("ff", 0.0)
        sr = synth.make_sample_reads(world_obj, seed, coverage=wl["coverage"], region_bp=wl["region_bp"], n_true_haps=3 if ff > 0 else 2,
                                     hap_probs=(0.5, 0.5 - ff / 2, ff / 2) if ff > 0 else None)
        calls += schedule.sample_calls(world_obj, sr, seed + 7, K=wl["K"], ff=ff, impute_rare_common=wl["factor"] > 0)
    log(f"inputs: {n_samples} samples -> {len(calls)} Gibbs calls in {time.time() - t0:.1f}s")
    return calls


def split_stages(calls, wl):
    """calls come chain-major (schedule.sample_calls: per chain n_seek_its common-SNP calls [+ the all-SNP call]); stage s holds
    call s of every chain.  Intermediate common-SNP calls keep their probabilities on the device (only the last common call and
    the all-SNP call of a chain are averaged by the caller, functions.R:999-1020 with n_burn_in_seek_its = 2)."""
    from quilt_b200 import cabi

    per_chain = 4 if wl["factor"] > 0 else 3
    stages = [calls[s::per_chain] for s in range(per_chain)]
    for s in range(2):
        for c in stages[s]:
            c.flags |= cabi.F_OUTPUT_NO_PROBS
    return stages


class ChainedStep:
    """the staged batches of one step and the device-resident chain call -> select -> call -> select -> call -> select -> all-SNP call"""

    def __init__(self, lib, stages, seed):
        from quilt_b200 import api

        self.lib = lib
        self.stages = stages
        self.batches = [api.Batch(lib, st) for st in stages]
        rng = np.random.default_rng(seed)
        self.pad = [rng.random(len(stages[s]) * stages[s + 1][0].K) for s in range(len(stages) - 1)]

    def run(self):
        for s, b in enumerate(self.batches):
            if s > 0:
                self.batches[s - 1].chain_select_into(b, self.pad[s - 1])
            b.run()
        for b in self.batches:
            b.sync()  # (the elapsed times of a batch's events are read at sync)

    def timing(self):
        t = {"total_ms": 0.0, "sweep_ms": 0.0, "n_sweep_launches": 0, "select_ms": 0.0}
        for s, b in enumerate(self.batches):
            tm = b.timing()
            t["total_ms"] += tm["total_ms"]
            t["sweep_ms"] += tm["sweep_ms"]
            t["n_sweep_launches"] += tm["n_sweep_launches"]
            if s > 0:
                t["select_ms"] += b.chain_ms()
        t["total_ms"] += t["select_ms"]
        return t

    def bytes(self):
        out = {"h2d_bytes": 0, "d2h_bytes": 0, "sweep_algorithmic_bytes": 0.0}
        for b in self.batches:
            for k, v in b.bytes().items():
                out[k] += v
        out["h2d_bytes"] += int(sum(p.nbytes for p in self.pad))
        return out

    def fetch(self):
        return [b.fetch() for b in self.batches]

    def free(self):
        for b in self.batches:
            b.free()


def cpu_checker(prefer: str = "reference"):
    """-> (library object with .gibbs(call), kind).  "reference" = oracle/_ref (the reference's own sources compiled against
    the RcppArmadillo stand-in), "port" = the oracle restatement.  Test / baseline infrastructure only."""
    if prefer == "reference":
        try:
            from oracle import ref_py

            if ref_py.available():
                return ref_py.Ref(), "reference"
        except Exception:
            pass
    from oracle.oracle_py import Oracle

    return Oracle(), "port"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def parity_gate(step, n_chains, log, prefer="reference"):
    """One whole sample (the first n_chains jobs of every stage) re-done on the CPU, stage by stage like the device chain: the CPU
    checker runs the Gibbs calls (exact labels / H_class / GT, DS and GP to 1e-4 against the GPU outputs that are shipped), the oracle's
    selection section re-derives every haplotype list from the CPU's own hapProbs_t and it must equal the list the device wrote."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle.oracle_py import Oracle

    chk, kind = cpu_checker(prefer)
    orc = Oracle()
    threads = max(1, min(host_cores(), n_chains))
    t0 = time.perf_counter()
    gpu_res = step.fetch()
    calls, results, ref = [], [], []
    sel_mismatch, sel_checked = 0, 0
    prev_cpu = None
    for s, st in enumerate(step.stages):
        cs = st[:n_chains]
        if s > 0:
            K = cs[0].K
            for j, c in enumerate(cs):
                dev_list = step.batches[s].which_haps(j)
                want = orc.select_haps_padded(c.panel, prev_cpu[j].hapProbs_t, K, step.pad[s - 1][j * K:(j + 1) * K], nHap=2 if c.ff == 0 else 3)
                sel_checked += 1
                sel_mismatch += int(not np.array_equal(dev_list, want))
                c.which_haps_to_use = dev_list  # the CPU call below runs on the list the device used
        with ThreadPoolExecutor(max_workers=threads) as ex:
            cpu = list(ex.map(chk.gibbs, cs))
        prev_cpu = cpu
        calls += cs
        results += gpu_res[s][:n_chains]
        ref += cpu
    gate = {"calls": len(calls), "checker": kind, "label_mismatches": 0, "H_class_mismatches": 0, "gt_mismatches": 0, "underflow_mismatches": 0,
            "max_dDS": 0.0, "max_dGP": 0.0, "tolerance": 1e-4, "selections_checked": sel_checked, "selection_mismatches": sel_mismatch}
    for c, g, o in zip(calls, results, ref):
        if bool(g.underflow_problem) != bool(o.underflow_problem):
            gate["underflow_mismatches"] += 1
            continue
        if o.underflow_problem:
            continue
        gate["label_mismatches"] += int(np.sum(g.H != o.H))
        gate["H_class_mismatches"] += int(np.sum(g.H_class != o.H_class))
        from quilt_b200 import cabi as _cabi

        if c.flags & _cabi.F_OUTPUT_NO_PROBS:
            continue  # probabilities of intermediate calls stay on the device; the NEXT stage's list check covers them
        gate["gt_mismatches"] += int(np.sum(np.argmax(g.genProbsM_t, axis=0) != np.argmax(o.genProbsM_t, axis=0)))
        nh = 2 if c.ff == 0 else 3
        gate["max_dDS"] = max(gate["max_dDS"], float(np.max(np.abs(g.hapProbs_t[:nh].sum(0) - o.hapProbs_t[:nh].sum(0)))))
        gate["max_dGP"] = max(gate["max_dGP"], float(np.max(np.abs(g.genProbsM_t - o.genProbsM_t))), float(np.max(np.abs(g.genProbsF_t - o.genProbsF_t))))
    gate["seconds"] = time.perf_counter() - t0
    gate["passed"] = (gate["label_mismatches"] == 0 and gate["H_class_mismatches"] == 0 and gate["gt_mismatches"] == 0 and gate["underflow_mismatches"] == 0
                      and gate["selection_mismatches"] == 0 and gate["max_dDS"] <= 1e-4 and gate["max_dGP"] <= 1e-4)
    log(f"parity gate ({kind}, {threads} threads, {gate['seconds']:.1f}s): {gate}")
    return gate


def cpu_arm(wl, world_obj, log, prefer="reference"):
    """The CPU implementation timed on the host cores: each thread runs one common-SNP call and one all-SNP call of its own
    sample concurrently with all the others (even threads: an iterative-initialisation call, odd threads: a normal one);
    per-sample time = 8 t_iterative + 16 t_normal + 8 t_all (the QUILT2 schedule) — a bounded sample, extrapolated."""
    from concurrent.futures import ThreadPoolExecutor

    import psutil

    from quilt_b200 import synth

    orc, kind_impl = cpu_checker(prefer)
    cores = host_cores()
    K = wl["K"]
    per_thread_gb = (8.0 * K * 20000 + 4.0 * K * 20000 + 3 * 2 * 8.0 * K * 3000) / 1e9 * 1.3 + 0.3
    mem_cap = max(1, int(psutil.virtual_memory().available / 1e9 / per_thread_gb))
    threads = max(1, min(cores, mem_cap))
    jobs = []
    ff = wl.get("ff", 0.0)
    has_all = wl["factor"] > 0
    for t in range(threads):
        sr = synth.make_sample_reads(world_obj, 777000 + t, coverage=wl["coverage"], region_bp=wl["region_bp"], n_true_haps=3 if ff > 0 else 2,
                                     hap_probs=(0.5, 0.5 - ff / 2, ff / 2) if ff > 0 else None)
        kind = "iterative" if t % 2 == 0 else "normal"
        jobs.append((kind, synth.make_call(world_obj, sr.common, 5000 + t, K=K, first_iteration=(kind == "iterative"), ff=ff)))
        if has_all:
            jobs.append(("all", synth.make_call(world_obj, sr.all, 6000 + t, K=K, all_snps=True, sort_haps=False, ff=ff)))

    def run(job):
        t0 = time.perf_counter()
        orc.gibbs(job[1])
        return job[0], time.perf_counter() - t0

    t0 = time.perf_counter()
    # two phases so that every core is busy with the same kind of call while it is being timed
    with ThreadPoolExecutor(max_workers=threads) as ex:
        r1 = list(ex.map(run, [j for j in jobs if j[0] != "all"]))
        r2 = list(ex.map(run, [j for j in jobs if j[0] == "all"]))
    wall = time.perf_counter() - t0
    tt = {"iterative": [], "normal": [], "all": []}
    for k, v in r1 + r2:
        tt[k].append(v)
    t_it = statistics.mean(tt["iterative"]) if tt["iterative"] else statistics.mean(tt["normal"])
    t_no = statistics.mean(tt["normal"]) if tt["normal"] else t_it
    t_all = statistics.mean(tt["all"]) if tt["all"] else 0.0
    per_sample = 8 * t_it + 16 * t_no + 8 * t_all
    value = threads / per_sample
    log(f"cpu arm: {threads} threads ({cores} cores), t_iterative {t_it:.2f}s t_normal {t_no:.2f}s t_all {t_all:.2f}s -> {value:.4f} samples/s (wall {wall:.1f}s)")
    return {
        "value": value,
        "unit": "samples/s",
        "cores": threads,
        "kind": kind_impl,
        "impl": ("oracle/_ref/libquiltref.so: unmodified QUILT/src sources compiled against the RcppArmadillo stand-in" if kind_impl == "reference"
                 else "oracle/libquiltoracle.so: statement-order restatement"),
        "sample": f"{threads} concurrent threads x (1 common-SNP call + 1 all-SNP call) of the same workload, timed once and EXTRAPOLATED to the "
                  f"32-call schedule: per-sample time = 8*t_iterative + 16*t_normal + 8*t_all",
        "seconds_per_call": {"iterative": t_it, "normal": t_no, "all_snps": t_all},
        "host_cores": cores,
    }, wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chr20_2Mb_1x_K4096", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="samples per GPU per step (default: workload's)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-gate", action="store_true", help="experiments only: a line without the gate is not a benchmark result")
    ap.add_argument("--cpu-impl", default="reference", choices=["reference", "port"], help="CPU arm / gate checker: oracle/_ref or the oracle port")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = dict(WORKLOADS[args.workload])
    global METRIC
    if args.workload != "chr20_2Mb_1x_K4096":
        METRIC = f"samples/sec, workload {args.workload} (not the BASELINE.json headline configuration)"
    if args.samples > 0:
        wl["samples"] = args.samples

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    def log(msg):
        print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    from quilt_b200 import synth

    config = {
        "workload": args.workload,
        "K": wl["K"], "K_full": wl["K_full"], "nSNPs_common": wl["nSNPs"], "nSNPs_all": wl["nSNPs"] * wl["factor"],
        "nGrids": (wl["nSNPs"] + 31) // 32, "nGrids_all": (wl["nSNPs"] * wl["factor"] + 31) // 32,
        "coverage": wl["coverage"], "samples_per_gpu_per_step": wl["samples"], "calls_per_sample": 32 if wl["factor"] > 0 else 24, "ff": wl.get("ff", 0.0),
        "schedule": "8 chains x (3 common-SNP calls + 1 all-SNP call), 20+1 sweeps, shard passes at sweeps 3/6/9",
        "parallelism": f"samples sharded over {args.gpus} GPU(s), no data-path collective",
        "l2": "per-wave working set (>= 29 GB of alpha/beta/eMatGrid columns) is far larger than the 126 MB L2",
    }

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        w = synth.make_world(WORLD_SEED, K_full=wl["K_full"], nSNPs=wl["nSNPs"], region_bp=wl["region_bp"], all_snps_factor=wl["factor"])
        vals, walls = [], []
        cb = None
        # each "step" is the bounded sample; warm-up steps would only repeat ~25 s of CPU work, one is enough for page-in
        for i in range(max(1, min(args.steps, 2))):
            cb, wall = cpu_arm(wl, w, log, args.cpu_impl)
            vals.append(cb["value"])
            walls.append(wall)
        v = statistics.mean(vals)
        cb["value"] = v
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": 0,
            "ms_per_step": 1e3 * statistics.mean(walls), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU arm = " + cb["impl"] + "; one single-threaded call per host thread on all host cores (the reference has no OpenMP); "
                    "the per-step value is extrapolated from one timed call of each kind per thread (see cpu_baseline.sample)",
        }
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------------------------------ B200 arm
    import torch

    from quilt_b200 import api, dist

    rank, local_rank, world, dev = dist.init()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    lib = api.GpuLib()
    lib.set_device(local_rank)
    t0 = time.time()
    w = None
    if rank == 0:
        w = synth.make_world(WORLD_SEED, K_full=wl["K_full"], nSNPs=wl["nSNPs"], region_bp=wl["region_bp"], all_snps_factor=wl["factor"])
    w = dist.broadcast_world(w, src=0)
    log(f"world ready in {time.time() - t0:.1f}s")
    calls = build_inputs(wl, rank, w, log)
    n_samples = wl["samples"]

    # ---- staged batches: inputs resident in HBM; the haplotype lists of stages 2 .. 4 are produced on the device
    t0 = time.time()
    stages = split_stages(calls, wl)
    batch = ChainedStep(lib, stages, 4242 + rank)
    nbytes = batch.bytes()
    log(f"staged {len(calls)} calls in {len(stages)} chained batches: h2d {nbytes['h2d_bytes'] / 1e9:.2f} GB in {time.time() - t0:.1f}s")
    for i in range(args.warmup):
        batch.run()
        log(f"warmup {i}: {batch.timing()['total_ms']:.1f} ms")
    # ---- parity gate (rank 0, its first sample): nothing below counts unless the GPU outputs match the CPU checker
    gate = None
    if rank == 0 and not args.no_parity_gate:
        gate = parity_gate(batch, 8, log, args.cpu_impl)
    dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    n0 = lib.kernel_launches()
    wall0 = time.perf_counter()
    dev_ms, sweep_ms, sweep_launches, select_ms = 0.0, 0.0, 0, 0.0
    for i in range(args.steps):
        batch.run()
        tm = batch.timing()
        dev_ms += tm["total_ms"]
        sweep_ms += tm["sweep_ms"]
        sweep_launches += tm["n_sweep_launches"]
        select_ms += tm["select_ms"]
    torch.cuda.synchronize()
    dist.barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches = lib.kernel_launches() - n0
    clocks = sampler.stop()
    ms_per_step_local = dev_ms / args.steps
    ms_per_step = dist.max_over_ranks(ms_per_step_local)
    wall_per_step = dist.max_over_ranks(wall_ms / args.steps)
    total_samples = n_samples * world
    value = total_samples / (ms_per_step / 1e3)
    log(f"timed: {ms_per_step_local:.1f} ms/step device, {wall_ms / args.steps:.1f} ms/step wall, sweep share {sweep_ms / dev_ms:.3f}")

    # ---- roofline of the dominant kernel (rank 0's launches)
    peak, peak_src = measured_peak()
    alg_bytes_per_run = nbytes["sweep_algorithmic_bytes"]
    launches_per_run = sweep_launches / args.steps
    avg_launch_s = (sweep_ms / 1e3) / max(sweep_launches, 1)
    achieved = (alg_bytes_per_run / max(launches_per_run, 1)) / avg_launch_s / 1e9
    tr = ncu_traffic(args.workload)
    roofline = {
        "bound": "hbm", "kernel": "k_sweep", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": (tr or {}).get("dram_bytes_per_launch"),
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes_per_run / max(launches_per_run, 1),
        "avg_launch_ms": 1e3 * avg_launch_s, "launches_per_step": launches_per_run, "share_of_step": sweep_ms / dev_ms,
        "selection_ms_per_step": select_ms / args.steps,
        "note": "algorithmic bytes follow SURVEY.md §8(d) (dense fp64 eMatRead columns counted); the kernel moves fewer bytes "
                "because emission columns are rebuilt from 2^nb-entry tables + bit-packed alleles, see DESIGN.md",
    }
    batch_pad = batch.pad
    batch.free()

    # ---- e2e: host buffers in, host buffers out, every step
    # the caller's result buffers exist before the timed region (allocated and touched once, as a host that processes
    # batch after batch would keep them); everything else of the call is timed: job preparation, pinned staging, H2D,
    # kernels, D2H, unpacking into the caller's arrays
    # every step: the whole chain through quilt_gpu_gibbs_chain with HOST buffers — per stage job preparation, pinned staging, H2D,
    # device-side selection of the haplotype lists from the previous stage, kernels, D2H, unpacking into the caller's arrays, one wave
    # pipeline across the stages; the caller's argument / result structures exist before the timed region (a host that processes
    # batch after batch keeps them, like R keeps its per-worker scratch, quilt.R:731-762)
    preps = [lib.prepare(st, touch=True) for st in stages]
    e2e_s = []
    n_under = 0
    for i in range(max(1, args.e2e_steps)):
        dist.barrier()
        t0 = time.perf_counter()
        res = api.run_chain_prepared(lib, preps, batch_pad)
        e2e_s.append(time.perf_counter() - t0)
        log(f"e2e step {i}: {1e3 * e2e_s[-1]:.1f} ms")
        n_under = sum(int(r.underflow_problem) for st_res in res for r in st_res)
    e2e_step = dist.max_over_ranks(statistics.mean(e2e_s))
    e2e = {"value": total_samples / e2e_step, "unit": "samples/s", "h2d_bytes_per_step": nbytes["h2d_bytes"], "d2h_bytes_per_step": nbytes["d2h_bytes"],
           "ms_per_step": 1e3 * e2e_step, "steps": len(e2e_s), "underflow_calls": n_under,
           "path": "quilt_gpu_gibbs_chain: all stages of the call chain in one call (host buffers in / out, one wave pipeline, lists selected on the device)"}
    log(f"e2e: {1e3 * e2e_step:.1f} ms/step")

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_arm(wl, w, log, args.cpu_impl)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cb,
            "wall_ms_per_step": wall_per_step, "parity_gate": gate,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.shutdown()
    if gate is not None and not gate["passed"]:
        log("PARITY GATE FAILED: the throughput above is void")
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
