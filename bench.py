#!/usr/bin/env python
"""bench.py -- edge-samples/s of the Monte-Carlo hot path (run_many) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3] [--no-secondary]

Main line (BASELINE.json config 3, the configuration `metric` is quoted on): the 100k-event / 400k-activity network
DAG, mixed gamma + empirical-relative delays, **2^20 samples in total, strong-scaled**: each of the N ranks (one
process per GPU, seeds sharded in contiguous blocks, no data-path collective) runs 2^20 / N samples per step as
back-to-back launches of equal, machine-filling size into recycled device buffers -- full outputs (durations[A,S],
realized[E,S], cause[E,S]) written to HBM.  `value` = 2^20 x activities x steps / time (CUDA events, max over ranks).

* `roofline`  -- the sweep kernel's launches inside the timed region, timed one by one with CUDA events on the launch
                 stream: algorithmic bytes per launch / mean launch time against the measured HBM copy bandwidth.
* `e2e`       -- the same metric through the reference-facing calls with HOST buffers: `mcdp_run_many_host[_multi]`
                 (pinned caller buffers), and the drop-in `MonteCarloPropagator.run_many` -> list[SimResult] /
                 `run_many_arrays`; host-to-device seeds and device-to-host results inside the timed region, against
                 the pinned D2H bandwidth measured in the same run (`e2e.roofline`).
* `secondary` -- the other named configurations at their stated sizes: C2 (2^20 samples, full outputs), C4 (262 144
                 samples over the N GPUs, reduced statistics, **NCCL all-reduce of the statistics inside the timed
                 region**), C5 (10 M samples reduced to per-event histograms), and the C3 DAG with generic-shape gamma
                 (Marsaglia-Tsang), each with its own roofline.
* `latency`   -- `run(seed)` through the drop-in class next to the reference's per-run time.
* `cpu_baseline` -- the unmodified reference (oracle/_ref) on this box's host cores, bounded sample (N = 1 only).

`--impl reference` times only the CPU arm on the main line's config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TOTAL_SAMPLES = {"c1": 1 << 20, "c2": 1 << 20, "c3": 1 << 20, "c4": 1 << 18, "c5": 10_000_000}
WORKLOADS = {
    # name: (generator, description)
    "c1": ("c1_toy", "toy demo DAG 10 events / 12 activities, 200-point empirical tables"),
    "c2": ("c2_layered", "layered timetable DAG 10k events / 30k activities, exponential delays"),
    "c3": ("c3_network", "network DAG 100k events / 400k activities, gamma + empirical-relative"),
    "c4": ("c4_national", "national DAG 1M events / 4M activities, fan-in up to 16 (reduced statistics mode)"),
    "c5": ("c5_deep_chain", "50k-event chain + 200 merge nodes fan-in 256, empirical-absolute (reduced statistics mode)"),
}
REDUCED_WORKLOADS = {"c4", "c5"}
FULL_BUFFER_BUDGET = 100e9  # bytes of recycled output buffers per launch (HBM: 180 GB)
THRESHOLDS, N_BINS = (60.0, 180.0, 300.0), 64


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the benchmark runs; `window` summarises a time span."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in list(self.rows) if t0 - 0.05 <= r[0] <= t1 + 0.15] or list(self.rows)[-3:]
        for _, line in rows:
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


def build_workload(name, variant=None):
    from mc_dagprop_b200 import synth
    from mc_dagprop_b200.flat import FlatDists

    dag, dists = getattr(synth, WORKLOADS[name][0])()
    if variant == "mt":  # the same DAG with generic-shape gamma: every gamma draw goes through Marsaglia-Tsang
        x = np.linspace(0.0, 3.0, 256)
        dists = FlatDists()
        dists.add_gamma(1, 2.3, 0.1, 5.0)
        dists.add_gamma(2, 0.6, 0.3, 5.0)
        dists.add_empirical_relative(3, x, np.exp(-x))
        dists.add_empirical_relative(4, x, np.exp(-x))
    return dag, dists


def bytes_per_edge_sample(E, A, reduced):
    # SURVEY 8(d) / DESIGN.md section 6: full 16 + 12 E/A, reduced 8 + 8 E/A
    return (8.0 + 8.0 * E / A) if reduced else (16.0 + 12.0 * E / A)


# ----------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref/libmcdp_ref.so), one Simulator per host thread
# ----------------------------------------------------------------------------------------------
def cpu_reference_throughput(dag, dists, target_seconds: float, threads: int | None = None, steps: int = 1,
                             warmup: int = 0):
    """edge-samples/s of the reference's own C++ Simulator::run loop (== run_many, _core.cpp:355-361)
    with one Simulator per thread (the reference's only parallel pattern, test_simulator.py:201-215)."""
    import oracle

    kind = "reference" if oracle.have_ref() else "port"
    Sim = oracle.RefSim if kind == "reference" else oracle.OracleSim
    threads = threads or os.cpu_count() or 1
    sims = [None] * threads

    def make(i):
        sims[i] = Sim(dag, dists)

    ts = [threading.Thread(target=make, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    A = sims[0].A
    # calibrate: one seed on one thread (also the reference's run(seed) latency)
    sims[0].run_many(np.arange(1, dtype=np.int32), realized=True, durations=True, cause=True)
    t0 = time.perf_counter()
    sims[0].run_many(np.arange(1, 3, dtype=np.int32), realized=True, durations=True, cause=True)
    per_seed = max((time.perf_counter() - t0) / 2, 1e-6)
    per_thread = max(1, int(target_seconds / per_seed / max(steps + warmup, 1)))
    per_thread = min(per_thread, 1 << 16)

    def work(i, base):
        sims[i].run_many(np.arange(base + i * per_thread, base + (i + 1) * per_thread, dtype=np.int32))

    times = []
    for s in range(warmup + steps):
        ts = [threading.Thread(target=work, args=(i, s * threads * per_thread)) for i in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = threads * per_thread * A * len(times) / total
    return {"value": value, "unit": "edge-samples/s", "cores": threads, "kind": kind,
            "sample": f"{per_thread} seeds x {threads} threads x {len(times)} steps, one Simulator per thread, "
                      f"full SimResult copies",
            "run_seed_ms_one_thread": per_seed * 1e3}, total / len(times)


def main_config(wl, E, A, world, total, per_gpu, launch, n_launches, reduced):
    return {"workload": f"{wl}: {WORKLOADS[wl][1]}", "events": E, "activities": A,
            "total_samples_per_step": total, "samples_per_gpu_per_step": per_gpu,
            "launches_per_gpu_per_step": n_launches, "samples_per_launch": launch,
            "mode": "reduced per-event statistics" if reduced else
                    "full outputs (realized, durations, cause_event per sample) written to HBM",
            "l2": "outputs per launch exceed L2 by orders of magnitude (no reuse between launches)",
            "parallelism": f"seeds sharded over {world} GPU(s), " +
                           ("NCCL all-reduce of the statistics" if reduced and world > 1 else "no data-path collective")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    dag, dists = build_workload(wl)
    E, A = dag.n_events, dag.n_activities
    cb, ms = cpu_reference_throughput(dag, dists, target_seconds=60.0, steps=args.steps, warmup=args.warmup)
    total = args.samples or TOTAL_SAMPLES[wl]
    # the same `config` the GPU arm prints for this command line (launch split from a host-only plan: no GPU needed)
    from mc_dagprop_b200 import capi

    world = max(args.gpus, 1)
    per_gpu = total // world
    reduced = wl in REDUCED_WORKLOADS or args.reduced
    if reduced:
        n_launches, launch = 1, per_gpu
    else:
        plan = capi.Plan(dag, dists, device=capi.DEVICE_NONE)
        cap = int(FULL_BUFFER_BUDGET // (12 * E + 8 * A))
        shape = plan.launch_shape(max(min(per_gpu, cap), 1))
        quad = shape["samples_per_lane"] == 4
        ctas_per_sm = max(1, (20 if quad else 32) // (shape["warps_per_group"] * shape["groups_per_cta"]))
        n_launches, launch = split_launches(per_gpu, cap, 148 * ctas_per_sm * shape["groups_per_cta"] * (128 if quad else 64))
    line = {
        "impl": "reference", "metric": "edge-samples/s", "value": cb["value"], "unit": "edge-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": main_config(wl, E, A, world, total, per_gpu, launch, n_launches, reduced),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "edge-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Ctx:
    """torch / torch.distributed plumbing of one rank."""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.stream = torch.cuda.current_stream()
        self.sp = self.stream.cuda_stream

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def split_launches(per_gpu: int, cap: int, wave: int):
    """Launches of whole waves of sample groups, as large as the recycled output buffers allow; what is left over is
    one smaller launch (a launch with fewer groups than SMs is spread over thread-block clusters by the library, so
    a small tail costs less than an equal share of a full launch would)."""
    if per_gpu <= 0:
        return 0, 0
    size = max(wave, cap // wave * wave)
    size = min(size, -(-per_gpu // 128) * 128)
    return -(-per_gpu // size), size


def timed_workload(cx: Ctx, wl: str, total: int, steps: int, warmup: int, variant=None, force_reduced=False, opts=None,
                   sampler=None, small_warmup=False, burst=False):
    """One workload at its stated size: `steps` timed passes over `total` samples (all ranks together).
    `small_warmup`: only the first warm-up pass runs at full size (it allocates the scratch the timed passes use),
    the others at an eighth -- for configurations whose single pass takes seconds."""
    torch = cx.torch
    from mc_dagprop_b200 import capi

    reduced = wl in REDUCED_WORKLOADS or force_reduced
    dag, dists = build_workload(wl, variant)
    plan = capi.Plan(dag, dists, device=cx.local_rank)
    for k, v in (opts or {}).items():
        plan.set_option(k, v)
    E, A = plan.E, plan.A
    per_gpu_lo = total * cx.rank // cx.world
    per_gpu = total * (cx.rank + 1) // cx.world - per_gpu_lo
    bpe = bytes_per_edge_sample(E, A, reduced)
    evs_launch = []  # (start, stop, samples) of every launch in the timed region
    if reduced:
        desc = capi.make_stats_desc(thresholds=THRESHOLDS, n_bins=N_BINS, hist_range=(0.0, dag.max_delay))
        s_sum = torch.zeros(E, dtype=torch.float64, device=cx.dev)
        s_sq = torch.zeros(E, dtype=torch.float64, device=cx.dev)
        s_late = torch.zeros((len(THRESHOLDS), E), dtype=torch.int64, device=cx.dev)
        s_hist = torch.zeros((E, N_BINS), dtype=torch.int32, device=cx.dev)
        stats = (s_sum, s_sq, s_late, s_hist)
        # one call per step; the library splits it into launches by its scratch budget (whole waves of sample groups)
        launch = plan.reduced_chunk(per_gpu, N_BINS) if per_gpu else 0
        n_launches = -(-per_gpu // launch) if launch else 0

        def step(i, timed, frac=1):
            for b in stats:
                b.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(cx.stream)
            plan.run_reduced_device(per_gpu // frac, desc, s_sum, s_sq, s_late, s_hist, seed0=per_gpu_lo + i * total, stream=cx.sp)
            e1.record(cx.stream)
            if timed:
                evs_launch.append((e0, e1, per_gpu))
            if cx.dist is not None:  # the one collective of the system: the statistics over NVLink / NVSwitch
                for b in stats:
                    cx.dist.all_reduce(b)
    else:
        cap = int(FULL_BUFFER_BUDGET // (12 * E + 8 * A))
        shape = plan.launch_shape(max(min(per_gpu, cap), 1))
        sms = torch.cuda.get_device_properties(cx.dev).multi_processor_count
        quad = shape["samples_per_lane"] == 4
        ctas_per_sm = max(1, (20 if quad else 32) // (shape["warps_per_group"] * shape["groups_per_cta"]))
        wave = sms * ctas_per_sm * shape["groups_per_cta"] * (128 if quad else 64)
        n_launches, launch = split_launches(per_gpu, cap, wave)
        ld = launch
        realized = torch.empty((E, ld), dtype=torch.float64, device=cx.dev)
        durations = torch.empty((A, ld), dtype=torch.float64, device=cx.dev)
        cause = torch.empty((E, ld), dtype=torch.int32, device=cx.dev)

        def step(i, timed, frac=1):
            done = 0
            while done < per_gpu // frac:
                m = min(launch, per_gpu - done)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(cx.stream)
                plan.run_full_device(m, realized, durations, cause, ld, seed0=per_gpu_lo + done + i * total, stream=cx.sp)
                e1.record(cx.stream)
                if timed:
                    evs_launch.append((e0, e1, m))
                done += m

    for i in range(warmup):
        step(-1 - i, False, 8 if (small_warmup and i > 0) else 1)
    cx.barrier()
    t_wall0 = time.time()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(cx.stream)
    for i in range(steps):
        step(i, True)
    ev1.record(cx.stream)
    cx.barrier()
    t_wall1 = time.time()
    total_ms = cx.max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None

    value = total * A * steps / (total_ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    shape = plan.launch_shape(max(launch, 1), reduced, N_BINS if reduced else 0)
    if reduced:  # timed per call; a call is n_launches launches back to back
        full = [(a.elapsed_time(b) / max(n_launches, 1), min(launch, m)) for a, b, m in evs_launch]
        achieved = per_gpu * A * bpe / (float(np.mean([a.elapsed_time(b) for a, b, _ in evs_launch])) * 1e-3) / 1e9
        avg_ms = float(np.mean([t for t, _ in full]))
    else:
        full = [(a.elapsed_time(b), m) for a, b, m in evs_launch if m == launch] or [(a.elapsed_time(b), m) for a, b, m in evs_launch]
        avg_ms = float(np.mean([t for t, _ in full]))
        achieved = full[0][1] * A * bpe / (avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "traffic_source": None, "peak_source": peak_src, "bytes_per_edge_sample": bpe,
                "kernel": "mcdp::quad_sweep_kernel" if shape["samples_per_lane"] == 4 else "mcdp::chunk_sweep_kernel",
                "avg_launch_ms": avg_ms, "samples_per_launch": full[0][1], "launches_timed": len(full),
                "launch": {k: shape[k] for k in ("samples_per_lane", "warps_per_group", "groups_per_cta", "threads", "grid", "cluster")}}
    if reduced:
        roofline["note"] = (f"one API call per step, split by the library into {n_launches} launch(es) of <= {launch} samples "
                            "(scratch budget, whole waves); avg_launch_ms = call time / launches")
    tail = [] if reduced else [(a.elapsed_time(b), m) for a, b, m in evs_launch if m != launch]
    if tail:
        roofline["tail_launch"] = {"samples": tail[0][1], "avg_ms": float(np.mean([t for t, _ in tail]))}
    prof = os.path.join(ROOT, "profiles", f"traffic_{wl}.json")
    if os.path.exists(prof) and variant is None:
        try:
            with open(prof) as f:
                tj = json.load(f)
            # DRAM bytes of ONE captured launch (ncu --set full), scaled to this run's launch size: a property of the
            # capture named in traffic_source, not of this run
            per_es = tj["dram_bytes_per_launch"] / (tj["samples_per_launch"] * A)
            roofline["traffic"] = per_es * full[0][1] * A
            roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass
    if not reduced and burst:
        # the same launch timed alone after the device has idled: what the sustained number above loses to the power cap
        # (a sustained run of this kernel holds the board at its power limit and the SM clock ~9 % under its maximum)
        cx.torch.cuda.synchronize()
        time.sleep(3.0)
        bs = []
        for i in range(3):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(cx.stream)
            plan.run_full_device(min(launch, per_gpu), realized, durations, cause, ld, seed0=per_gpu_lo + (steps + 1 + i) * total, stream=cx.sp)
            e1.record(cx.stream)
            cx.torch.cuda.synchronize()
            bs.append(e0.elapsed_time(e1))
            time.sleep(1.0)
        b_ms = float(np.min(bs))
        b_ach = min(launch, per_gpu) * A * bpe / (b_ms * 1e-3) / 1e9
        roofline["burst"] = {"launch_ms": b_ms, "achieved": b_ach, "frac": b_ach / peak,
                             "how": "one launch of the same size timed alone after 3 s of idle, best of 3 (SM clock at its maximum)"}
    out = {"value": value, "ms_per_step": total_ms / steps, "steps": steps, "warmup": warmup,
           "config": main_config(wl, E, A, cx.world, total, per_gpu, launch, n_launches, reduced),
           "roofline": roofline, "clocks": clocks, "gpu_launches": len(evs_launch) * (n_launches if reduced else 1)}
    if variant == "mt":
        out["config"]["workload"] = "c3-mt: the C3 DAG with gamma shapes 2.3 / 0.6 (Marsaglia-Tsang) + empirical-relative"
    if reduced and cx.dist is not None:
        out["config"]["collective"] = "dist.all_reduce (NCCL) of sum / sumsq / late / hist inside the timed region, every step"
    if not reduced:
        del realized, durations, cause
    plan.close()
    torch.cuda.empty_cache()
    return out, (dag, dists)


def pinned_d2h_gbs(cx: Ctx, gib: float = 2.0):
    """Pinned device-to-host copy bandwidth, all ranks copying at the same time.  Returns (best, sustained_all):
    `best` = this rank's best of five copies of `gib` GiB, each timed on its own; `sustained_all` = what all ranks
    together move when every rank issues four copies back to back after a barrier, total bytes / slowest rank's time
    -- the figure a call that shards equal blocks over the GPUs can reach."""
    torch = cx.torch
    n = int(gib * (1 << 30))
    d = torch.empty(n, dtype=torch.uint8, device=cx.dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.copy_(d, non_blocking=True)
    cx.barrier()
    best = 0.0
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = cx.max_over_ranks(time.perf_counter() - t0)
    del d, h
    return best, 4.0 * n * cx.world / dt / 1e9


def e2e_section(cx: Ctx, wl: str, dag, dists, steps: int, n_e2e: int):
    """Host seeds -> device -> host results through the public calls, copies inside the timed region."""
    torch = cx.torch
    from mc_dagprop_b200 import capi

    E, A = dag.n_events, dag.n_activities
    reduced = wl in REDUCED_WORKLOADS
    ke = max(1, min(steps, 3))
    d2h_peak, d2h_sustained_all = pinned_d2h_gbs(cx)
    d2h_peak_all = d2h_peak
    if cx.dist is not None:
        t = torch.tensor([d2h_peak], dtype=torch.float64, device=cx.dev)
        cx.dist.all_reduce(t)
        d2h_peak_all = float(t.item())
    res = {}
    if reduced:
        plan = capi.Plan(dag, dists, device=cx.local_rank)
        ne = n_e2e
        seeds = np.arange(cx.rank * ne, (cx.rank + 1) * ne, dtype=np.int32)
        plan.run_reduced_host(seeds, thresholds=THRESHOLDS, n_bins=N_BINS, hist_range=(0.0, dag.max_delay))
        cx.barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            plan.run_reduced_host(seeds + i * ne * cx.world, thresholds=THRESHOLDS, n_bins=N_BINS, hist_range=(0.0, dag.max_delay))
        dt = cx.max_over_ranks(time.perf_counter() - t0)
        plan.close()
        return {"value": cx.world * ne * A * ke / dt, "unit": "edge-samples/s", "h2d_bytes_per_step": 4 * ne,
                "d2h_bytes_per_step": E * (8 + 8 + 3 * 8 + 64 * 4), "samples_per_step": cx.world * ne, "steps": ke,
                "api": "mcdp_run_reduced_host, one call per rank"}

    bytes_per_sample = 12 * E + 8 * A
    # (1) C ABI with caller-pinned buffers, one call per rank on its own GPU (n_e2e samples in total)
    ne = max(128, n_e2e // cx.world)
    plan = capi.Plan(dag, dists, device=cx.local_rank)
    h_seeds = torch.empty(ne, dtype=torch.int32).pin_memory()
    h_r = torch.empty((ne, E), dtype=torch.float64).pin_memory()
    h_d = torch.empty((ne, A), dtype=torch.float64).pin_memory()
    h_c = torch.empty((ne, E), dtype=torch.int32).pin_memory()
    out = (h_r.numpy(), h_d.numpy(), h_c.numpy())

    def capi_step(i):
        h_seeds.copy_(torch.arange(i * ne, (i + 1) * ne, dtype=torch.int32) + cx.rank * 1000003)
        plan.run_many_host(h_seeds.numpy(), out=out)
        return float(out[0][-1, -1])

    capi_step(-1)
    cx.barrier()
    t0 = time.perf_counter()
    for i in range(ke):
        capi_step(i)
    dt = cx.max_over_ranks(time.perf_counter() - t0)
    plan.close()
    per_rank = {"value": cx.world * ne * A * ke / dt, "unit": "edge-samples/s", "samples_per_step": cx.world * ne, "steps": ke,
                "h2d_bytes_per_step": 4 * ne * cx.world, "d2h_bytes_per_step": bytes_per_sample * ne * cx.world,
                "d2h_gbs": bytes_per_sample * ne * cx.world * ke / dt / 1e9,
                "api": "mcdp_run_many_host (caller-pinned buffers), one process and one call per GPU"}
    res["capi_per_rank"] = per_rank
    headline = per_rank

    # (2) one process, all N devices behind ONE call: mcdp_run_many_host_multi (rank 0; the other ranks wait)
    if cx.world > 1:
        cx.barrier()
        multi = None
        if cx.rank == 0:
            try:
                ps = capi.PlanSet(dag, dists, list(range(cx.world)))
                nm = n_e2e
                seeds_m = np.arange(nm, dtype=np.int32)
                m_r = torch.empty((nm, E), dtype=torch.float64).pin_memory()
                m_d = torch.empty((nm, A), dtype=torch.float64).pin_memory()
                m_c = torch.empty((nm, E), dtype=torch.int32).pin_memory()
                out = (m_r.numpy(), m_d.numpy(), m_c.numpy())
                ps.run_many_host(seeds_m, out=out)
                t0 = time.perf_counter()
                for i in range(ke):
                    ps.run_many_host(seeds_m + (i + 1) * nm, out=out)
                dtm = time.perf_counter() - t0
                ps.close()
                del m_r, m_d, m_c
                multi = {"value": nm * A * ke / dtm, "unit": "edge-samples/s", "samples_per_step": nm, "steps": ke,
                         "h2d_bytes_per_step": 4 * nm, "d2h_bytes_per_step": bytes_per_sample * nm,
                         "d2h_gbs": bytes_per_sample * nm * ke / dtm / 1e9,
                         "api": f"mcdp_run_many_host_multi: one process, one call, {cx.world} devices (caller-pinned buffers)"}
            except Exception as exc:  # reported, not hidden
                multi = {"error": str(exc)}
        cx.barrier()
        if cx.rank == 0:
            res["capi_multi_one_process"] = multi
            if multi and "value" in multi:
                headline = multi
    del h_seeds, h_r, h_d, h_c, out
    ne = n_e2e

    # (3) the drop-in class (what a user of the reference calls): run_many -> list[SimResult], run_many_arrays
    if cx.rank == 0:
        try:
            from mc_dagprop import GenericDelayGenerator, MonteCarloPropagator

            gen = GenericDelayGenerator()
            for t in range(len(dists.dist_type)):
                ty, kind = int(dists.dist_type[t]), int(dists.kind[t])
                lo, hi = int(dists.tab_off[t]), int(dists.tab_off[t + 1])
                if kind == 0:
                    gen.add_constant(ty, float(dists.p0[t]))
                elif kind == 1:
                    gen.add_exponential(ty, float(dists.p0[t]), float(dists.p1[t]))
                elif kind == 2:
                    gen.add_gamma(ty, float(dists.p0[t]), float(dists.p1[t]), float(dists.p2[t]))
                elif kind == 3:
                    gen.add_empirical_absolute(ty, list(dists.tab_values[lo:hi]), list(dists.tab_weights[lo:hi]))
                else:
                    gen.add_empirical_relative(ty, list(dists.tab_values[lo:hi]), list(dists.tab_weights[lo:hi]))
            prop = MonteCarloPropagator.from_arrays(dag.earliest, dag.act_idx, dag.act_base, dag.act_type, dag.prec_target,
                                                    dag.prec_off, dag.pred_src, dag.pred_act, dag.max_delay, gen,
                                                    devices=list(range(cx.world)))
            nd = ne
            seeds_l = list(range(nd))
            r = prop.run_many(seeds_l)  # warm: pins the result pool
            del r
            t0 = time.perf_counter()
            for i in range(ke):
                r = prop.run_many([s + (i + 1) * nd for s in seeds_l])
                _ = float(r[-1].realized[-1])
                del r
            dtd = time.perf_counter() - t0
            a = prop.run_many_arrays(np.asarray(seeds_l, np.int32))
            del a
            t0 = time.perf_counter()
            for i in range(ke):
                a = prop.run_many_arrays(np.asarray(seeds_l, np.int32) + (i + 1) * nd)
                _ = float(a[0][-1, -1])
                del a
            dta = time.perf_counter() - t0
            res["drop_in_run_many"] = {"value": nd * A * ke / dtd, "unit": "edge-samples/s", "samples_per_step": nd, "steps": ke,
                                       "d2h_gbs": bytes_per_sample * nd * ke / dtd / 1e9,
                                       "api": "mc_dagprop.MonteCarloPropagator.run_many -> list[SimResult] (pinned result pool)"}
            res["drop_in_run_many_arrays"] = {"value": nd * A * ke / dta, "unit": "edge-samples/s", "samples_per_step": nd,
                                              "steps": ke, "d2h_gbs": bytes_per_sample * nd * ke / dta / 1e9,
                                              "api": "MonteCarloPropagator.run_many_arrays -> (realized, durations, cause) arrays"}
            # run(seed) latency, the interactive call
            prop.run(seed=1)
            lat = []
            for i in range(20):
                t0 = time.perf_counter()
                prop.run(seed=2 + i)
                lat.append((time.perf_counter() - t0) * 1e3)
            res["run_seed_ms"] = sorted(lat)[len(lat) // 2]  # median of 20 calls
            res["run_seed_ms_min"] = min(lat)
            del prop
        except Exception as exc:
            res["drop_in_error"] = str(exc)
    cx.barrier()
    e2e = dict(headline)
    e2e["roofline"] = {"bound": "pcie_d2h", "achieved": headline.get("d2h_gbs"), "unit": "GB/s",
                       "peak": d2h_sustained_all, "sum_of_per_gpu_best_of_5": d2h_peak_all, "this_gpu_best_of_5": d2h_peak,
                       "frac": (headline.get("d2h_gbs") or 0.0) / d2h_sustained_all if d2h_sustained_all else None,
                       "peak_source": f"pinned cudaMemcpy D2H, {cx.world} GPU(s) x 4 back-to-back copies of 2 GiB started together, total bytes / slowest rank's time, measured in this run"}
    e2e.update(res)
    return e2e


def small_dag_latency(cx: Ctx, with_reference: bool):
    """run(seed) on the toy DAG of demo/monte_carlo.py (config 1) through the drop-in class: pure call overhead."""
    import oracle
    from mc_dagprop import GenericDelayGenerator, MonteCarloPropagator

    dag, dists = build_workload("c1")
    gen = GenericDelayGenerator()
    for t in range(len(dists.dist_type)):
        lo, hi = int(dists.tab_off[t]), int(dists.tab_off[t + 1])
        gen.add_empirical_absolute(int(dists.dist_type[t]), list(dists.tab_values[lo:hi]), list(dists.tab_weights[lo:hi]))
    prop = MonteCarloPropagator.from_arrays(dag.earliest, dag.act_idx, dag.act_base, dag.act_type, dag.prec_target, dag.prec_off,
                                            dag.pred_src, dag.pred_act, dag.max_delay, gen, device=cx.local_rank)
    prop.run(seed=0)
    lat = []
    for i in range(200):
        t0 = time.perf_counter()
        prop.run(seed=i)
        lat.append((time.perf_counter() - t0) * 1e3)
    out = {"run_seed_ms": sorted(lat)[len(lat) // 2], "run_seed_ms_min": min(lat), "run_seed_ms_mean": sum(lat) / len(lat)}
    seeds = list(range(10000))
    prop.run_many(seeds)  # first call of this size: device buffers and the pinned block are allocated
    t0 = time.perf_counter()
    for _ in range(5):
        prop.run_many(seeds)
    out["run_many_10k_seeds_ms"] = (time.perf_counter() - t0) / 5 * 1e3  # 10k SimResult objects included
    prop.run_many_arrays(seeds)
    t0 = time.perf_counter()
    for _ in range(5):
        prop.run_many_arrays(seeds)
    out["run_many_arrays_10k_seeds_ms"] = (time.perf_counter() - t0) / 5 * 1e3
    if with_reference:
        sim = (oracle.RefSim if oracle.have_ref() else oracle.OracleSim)(dag, dists)
        sim.run_many(np.arange(10, dtype=np.int32))
        t0 = time.perf_counter()
        sim.run_many(np.arange(10000, dtype=np.int32))
        out["reference_run_many_10k_seeds_ms_one_thread"] = (time.perf_counter() - t0) * 1e3  # C++ engine, no Python objects
        out["reference_run_seed_ms_one_thread"] = out["reference_run_many_10k_seeds_ms_one_thread"] / 10000
    return out


def run_b200(args):
    import torch

    from mc_dagprop_b200 import capi

    if capi.device_count() < 1 or not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (mc_dagprop_b200 has no CPU fallback)")
    cx = Ctx()
    wl = args.workload
    total = args.samples or TOTAL_SAMPLES[wl]
    sampler = ClockSampler(cx.local_rank) if cx.rank == 0 else None
    opts = {}
    if args.wpg:
        opts[capi.OPT_WARPS_PER_GROUP] = args.wpg
    if args.gpc:
        opts[capi.OPT_GROUPS_PER_CTA] = args.gpc
    if args.spl:
        opts[capi.OPT_SAMPLES_PER_LANE] = args.spl
    main, (dag, dists) = timed_workload(cx, wl, total, args.steps, args.warmup, force_reduced=args.reduced, opts=opts,
                                        sampler=sampler, burst=True)
    reduced = wl in REDUCED_WORKLOADS or args.reduced

    secondary = []
    if not args.no_secondary:
        ks = max(1, min(args.steps, 3))
        for swl, variant, stotal, ssteps, swarm in (("c2", None, TOTAL_SAMPLES["c2"], ks, 3),
                                                     ("c3", "mt", 148 * 128 * cx.world * 4, ks, 3),
                                                     ("c4", None, TOTAL_SAMPLES["c4"], max(1, min(ks, 2)), 3),
                                                     ("c5", None, TOTAL_SAMPLES["c5"], max(1, min(ks, 2)), 3)):
            if swl == wl and variant is None:
                continue
            try:
                sec, _ = timed_workload(cx, swl, stotal, ssteps, swarm, variant=variant, sampler=sampler,
                                        small_warmup=swl in ("c4", "c5"))
                entry = {"workload": sec["config"]["workload"], "value": sec["value"], "unit": "edge-samples/s",
                         "n_gpus": cx.world, "ms_per_step": sec["ms_per_step"], "steps": sec["steps"], "warmup": sec["warmup"],
                         "config": sec["config"], "roofline": sec["roofline"], "clocks": sec["clocks"],
                         "gpu_launches": sec["gpu_launches"]}
            except Exception as exc:
                entry = {"workload": swl, "error": str(exc)}
                cx.torch.cuda.empty_cache()
            secondary.append(entry)

    e2e = None
    if not args.no_e2e:
        n_e2e = args.e2e_samples or {"c1": 1 << 18, "c2": 1 << 15, "c3": 1 << 12, "c4": 1 << 15, "c5": 1 << 18}[wl]
        e2e = e2e_section(cx, wl, dag, dists, args.steps, n_e2e)

    if cx.rank == 0:
        cb = None
        if cx.world == 1 and not args.no_cpu:
            cb, _ = cpu_reference_throughput(dag, dists, target_seconds=args.cpu_seconds)
        latency = None
        if e2e and "run_seed_ms" in e2e:
            latency = {"api": "mc_dagprop.MonteCarloPropagator.run(seed) -> SimResult (host result, one sample)",
                       wl: {"run_seed_ms": e2e.pop("run_seed_ms"), "run_seed_ms_min": e2e.pop("run_seed_ms_min", None),
                            "reference_run_seed_ms_one_thread": cb.get("run_seed_ms_one_thread") if cb else None}}
            try:
                latency["c1"] = small_dag_latency(cx, with_reference=cx.world == 1 and not args.no_cpu)
            except Exception as exc:
                latency["c1"] = {"error": str(exc)}
        line = {
            "metric": "edge-samples/s", "value": main["value"], "unit": "edge-samples/s", "n_gpus": cx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main["config"], "roofline": main["roofline"], "cpu_baseline": cb, "e2e": e2e,
            "gpu_launches": main["gpu_launches"], "clocks": main["clocks"], "latency": latency, "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if sampler:
        sampler.stop()
    if cx.dist is not None:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="TOTAL samples per step over all GPUs (default: the configuration's)")
    ap.add_argument("--e2e-samples", type=int, default=0)
    ap.add_argument("--reduced", action="store_true", help="force the reduced statistics mode")
    ap.add_argument("--wpg", type=int, default=0)
    ap.add_argument("--gpc", type=int, default=0)
    ap.add_argument("--spl", type=int, default=0, choices=[0, 2, 4], help="samples per lane: 2 pair kernel, 4 quad kernel, 0 auto")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
