#!/usr/bin/env python
"""bench.py -- edge-samples/s of the Monte-Carlo hot path (run_many) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5|c1] [--impl b200|reference]

A *step* is one run_many pass over a batch of seeds on every GPU (weak scaling: the per-GPU batch
is fixed, seeds are sharded by rank, there is no data-path collective in full-output mode).
`value` = edge-samples/s = ranks x samples x activities x steps / time, outputs written to HBM
(durations[A,S], realized[E,S], cause[E,S]), inputs resident.  `e2e` is the same metric through
the host-buffer C-ABI call (seeds from pinned host memory, all three result arrays copied back
to pinned host memory inside the timed region).  `roofline` relates the sweep kernel to the
measured HBM bandwidth; `cpu_baseline` is the unmodified reference (oracle/_ref) timed on this
box's host cores on a bounded sample.  `--impl reference` times only that CPU arm.
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

WORKLOADS = {
    # name: (generator, per-GPU samples per step (full mode), e2e samples per step, description)
    "c1": ("c1_toy", 1 << 20, 1 << 18, "toy demo DAG 10 events / 12 activities, 200-point empirical tables"),
    "c2": ("c2_layered", 1 << 18, 1 << 15, "layered timetable DAG 10k events / 30k activities, exponential delays"),
    "c3": ("c3_network", 18944, 1 << 11, "network DAG 100k events / 400k activities, gamma + empirical-relative"),
    "c4": ("c4_national", 1 << 15, 1 << 15, "national DAG 1M events / 4M activities (reduced statistics mode)"),
    "c5": ("c5_deep_chain", 1 << 18, 1 << 18, "50k-event chain + 200 merge nodes fan-in 256 (reduced statistics mode)"),
}
REDUCED_WORKLOADS = {"c4", "c5"}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

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

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than one sample: take whatever was seen
            for ts, line in self.rows:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[0]))
                    smax.append(float(parts[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name):
    from mc_dagprop_b200 import synth

    gen = getattr(synth, WORKLOADS[name][0])
    return gen()


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
    # calibrate: one seed on one thread
    t0 = time.perf_counter()
    sims[0].run_many(np.arange(1, dtype=np.int32), realized=True, durations=True, cause=True)
    per_seed = max(time.perf_counter() - t0, 1e-6)
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
                      f"full SimResult copies"}, total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dag, dists = build_workload(args.workload)
    E, A = dag.n_events, dag.n_activities
    cb, ms = cpu_reference_throughput(dag, dists, target_seconds=60.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "edge-samples/s", "value": cb["value"], "unit": "edge-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload][3]}", "events": E, "activities": A,
                   "mode": "full outputs (realized, durations, cause_event per sample)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "edge-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    from mc_dagprop_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.device_count() < 1 or not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (mc_dagprop_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    wl = args.workload
    reduced = wl in REDUCED_WORKLOADS or args.reduced
    dag, dists = build_workload(wl)
    plan = capi.Plan(dag, dists, device=local_rank)
    if args.wpg:
        plan.set_option(capi.OPT_WARPS_PER_GROUP, args.wpg)
    if args.gpc:
        plan.set_option(capi.OPT_GROUPS_PER_CTA, args.gpc)
    if args.spl:
        plan.set_option(capi.OPT_SAMPLES_PER_LANE, args.spl)
    E, A = plan.E, plan.A
    n = args.samples or WORKLOADS[wl][1]
    shape = plan.launch_shape(n, reduced, 64 if reduced else 0)
    kernel_name = "mcdp::quad_sweep_kernel" if shape["samples_per_lane"] == 4 else (
        "mcdp::sweep_kernel" if shape["batches"] > 1 else "mcdp::chunk_sweep_kernel")
    ld = n
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    if reduced:
        desc = capi.make_stats_desc(thresholds=(60.0, 180.0, 300.0), n_bins=64, hist_range=(0.0, dag.max_delay))
        s_sum = torch.zeros(E, dtype=torch.float64, device=dev)
        s_sq = torch.zeros(E, dtype=torch.float64, device=dev)
        s_late = torch.zeros((3, E), dtype=torch.int64, device=dev)
        s_hist = torch.zeros((E, 64), dtype=torch.int32, device=dev)

        def step(i):
            plan.run_reduced_device(n, desc, s_sum, s_sq, s_late, s_hist, seed0=(rank * 1000003 + i) * n, stream=sp)
    else:
        realized = torch.empty((E, ld), dtype=torch.float64, device=dev)
        durations = torch.empty((A, ld), dtype=torch.float64, device=dev)
        cause = torch.empty((E, ld), dtype=torch.int32, device=dev)

        def step(i):
            plan.run_full_device(n, realized, durations, cause, ld, seed0=(rank * 1000003 + i) * n, stream=sp)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(-1 - i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    evs[0].record(stream)
    for i in range(args.steps):
        step(i)
        evs[i + 1].record(stream)
    barrier()
    t_wall1 = time.time()
    total_ms = evs[0].elapsed_time(evs[-1])
    kernel_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        if reduced:  # the one collective of the system: reduce the statistics buffers over NVLink
            for buf in (s_sum, s_sq, s_late, s_hist):
                dist.all_reduce(buf)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    value = world * n * A * args.steps / (total_ms * 1e-3)
    bpe = bytes_per_edge_sample(E, A, reduced)
    peak, peak_src = measured_peak_gbs()
    avg_launch_s = float(np.mean(kernel_ms)) * 1e-3
    achieved = n * A * bpe / avg_launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "bytes_per_edge_sample": bpe,
                "kernel": kernel_name, "avg_launch_ms": avg_launch_s * 1e3,
                "launch": {k: shape[k] for k in ("samples_per_lane", "warps_per_group", "groups_per_cta", "threads", "grid")}}
    prof = os.path.join(ROOT, "profiles", f"traffic_{wl}.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- e2e: host seeds -> device -> host results through mcdp_run_many_host / run_reduced_host ----
    e2e = None
    if rank == 0 or world > 1:
        ne = args.e2e_samples or WORKLOADS[wl][2]
        if not reduced:
            del realized, durations, cause
            torch.cuda.empty_cache()
            h_seeds = torch.empty(ne, dtype=torch.int32).pin_memory()
            h_r = torch.empty((ne, E), dtype=torch.float64).pin_memory()
            h_d = torch.empty((ne, A), dtype=torch.float64).pin_memory()
            h_c = torch.empty((ne, E), dtype=torch.int32).pin_memory()
            out = (h_r.numpy(), h_d.numpy(), h_c.numpy())

            def e2e_step(i):
                h_seeds.copy_(torch.arange(i * ne, (i + 1) * ne, dtype=torch.int32))
                plan.run_many_host(h_seeds.numpy(), out=out)
                return float(out[0][-1, -1])

            h2d, d2h = 4 * ne, (12 * E + 8 * A) * ne
        else:
            def e2e_step(i):
                st = plan.run_reduced_host(np.arange(i * ne, (i + 1) * ne, dtype=np.int32),
                                           thresholds=(60.0, 180.0, 300.0), n_bins=64, hist_range=(0.0, dag.max_delay))
                return float(st.sum[-1])

            h2d, d2h = 4 * ne, E * (8 + 8 + 3 * 8 + 64 * 4)
        e2e_step(-1)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        ke = max(1, min(args.steps, 3))
        for i in range(ke):
            e2e_step(i)
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * ne * A * ke / dt, "unit": "edge-samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "samples_per_step": ne, "steps": ke,
               "api": "mcdp_run_reduced_host" if reduced else "mcdp_run_many_host (pinned host buffers)"}
        if not reduced:
            # for information: the same DAG through the statistics API (host seeds in, per-event mean / variance /
            # lateness counts / 64-bin histogram out) -- what a caller who does not need every sample would use
            nr = n
            th = (60.0, 180.0, 300.0)
            seeds_r = np.arange(nr, dtype=np.int32)
            plan.run_reduced_host(seeds_r, thresholds=th, n_bins=64, hist_range=(0.0, dag.max_delay))
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(ke):
                plan.run_reduced_host(seeds_r + i * nr, thresholds=th, n_bins=64, hist_range=(0.0, dag.max_delay))
            dtr = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dtr], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtr = float(t.item())
            e2e["reduced_api"] = {"value": world * nr * A * ke / dtr, "unit": "edge-samples/s", "samples_per_step": nr,
                                  "h2d_bytes_per_step": 4 * nr, "d2h_bytes_per_step": E * (8 + 8 + 3 * 8 + 64 * 4),
                                  "api": "mcdp_run_reduced_host"}

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu:
            cb, _ = cpu_reference_throughput(dag, dists, target_seconds=args.cpu_seconds)
        line = {
            "metric": "edge-samples/s", "value": value, "unit": "edge-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wl}: {WORKLOADS[wl][3]}", "events": E, "activities": A,
                       "samples_per_gpu_per_step": n,
                       "mode": "reduced per-event statistics" if reduced else
                               "full outputs (realized, durations, cause_event per sample) written to HBM",
                       "l2": "outputs per step exceed L2 by orders of magnitude (no reuse between steps)",
                       "parallelism": f"seeds sharded over {world} GPU(s), no data-path collective" +
                                      (", NCCL all-reduce of statistics" if reduced and world > 1 else "")},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": args.steps, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="per-GPU samples per step (default per workload)")
    ap.add_argument("--e2e-samples", type=int, default=0)
    ap.add_argument("--reduced", action="store_true", help="force the reduced statistics mode")
    ap.add_argument("--wpg", type=int, default=0)
    ap.add_argument("--gpc", type=int, default=0)
    ap.add_argument("--spl", type=int, default=0, choices=[0, 2, 4], help="samples per lane: 2 pair kernel, 4 quad kernel, 0 auto")
    ap.add_argument("--no-cpu", action="store_true")
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
