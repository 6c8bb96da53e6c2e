"""Scratch timing of the fused kernel (device buffers, CUDA events). Not the contract bench."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mc_dagprop_b200 import capi, synth
if os.environ.get("MCDP_LIB"):  # scratch A/B builds
    capi.LIB_PATH = os.path.abspath(os.environ["MCDP_LIB"])

def run(name, dag, dists, n, wpg=0, gpc=0, reps=3, spl=0):
    plan = capi.Plan(dag, dists, device=0)
    if spl: plan.set_option(capi.OPT_SAMPLES_PER_LANE, spl)
    if wpg: plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
    if gpc: plan.set_option(capi.OPT_GROUPS_PER_CTA, gpc)
    E, A = plan.E, plan.A
    ld = n
    dev = torch.device("cuda:0")
    realized = torch.empty((E, ld), dtype=torch.float64, device=dev)
    dur = torch.empty((A, ld), dtype=torch.float64, device=dev)
    cause = torch.empty((E, ld), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run_full_device(n, realized, dur, cause, ld, seed0=i * n, stream=st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:]) * 1e-3
    es = n * A / t
    bpe = 16 + 12 * E / A
    print(f"{name}: E={E} A={A} n={n} spl={spl} wpg={wpg} gpc={gpc} {t*1e3:.2f} ms  {es:.3e} edge-samples/s  "
          f"{es*bpe/1e9:.0f} GB/s algorithmic ({es*bpe/6546.9e9*100:.1f}% of 6546.9)", flush=True)
    del realized, dur, cause
    plan.close()
    return es


def run_reduced(name, dag, dists, n, wpg=0, gpc=0, reps=2, spl=0):
    """Reduced statistics (mean / variance / 3 lateness counts / 64-bin histogram), device buffers, CUDA events."""
    plan = capi.Plan(dag, dists, device=0)
    if spl: plan.set_option(capi.OPT_SAMPLES_PER_LANE, spl)
    if wpg: plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
    if gpc: plan.set_option(capi.OPT_GROUPS_PER_CTA, gpc)
    E, A = plan.E, plan.A
    dev = torch.device("cuda:0")
    desc = capi.make_stats_desc(thresholds=(60.0, 180.0, 300.0), n_bins=64, hist_range=(0.0, dag.max_delay))
    s_sum = torch.zeros(E, dtype=torch.float64, device=dev)
    s_sq = torch.zeros(E, dtype=torch.float64, device=dev)
    s_late = torch.zeros((3, E), dtype=torch.int64, device=dev)
    s_hist = torch.zeros((E, 64), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run_reduced_device(n, desc, s_sum, s_sq, s_late, s_hist, seed0=i * n, stream=st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:]) * 1e-3
    es = n * A / t
    shape = plan.launch_shape(n, True, 64)
    print(f"{name} reduced: E={E} A={A} n={n} spl={shape['samples_per_lane']} wpg={shape['warps_per_group']} "
          f"gpc={shape['groups_per_cta']} grid={shape['grid']} {t*1e3:.2f} ms  {es:.3e} edge-samples/s", flush=True)
    plan.close()
    return es

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    if which == "c2":
        dag, d = synth.c2_layered()
        for n in (262144,):
            for wpg, gpc in ((8, 1), (16, 1), (0, 0)):
                run("c2", dag, d, n, wpg, gpc)
    elif which == "c3":
        dag, d = synth.c3_network()
        for n in (18944,):
            for wpg, gpc in ((16, 1), (0, 0)):
                run("c3", dag, d, n, wpg, gpc)
