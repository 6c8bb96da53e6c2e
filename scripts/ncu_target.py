"""One workload / mode / launch size for ncu captures and quick timings (not the contract bench).

    python scripts/ncu_target.py c3 full 18944 [--mt] [--spl 4] [--reps 3] [--wpg W] [--gpc G] [--cluster C]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mc_dagprop_b200 import capi, synth
from mc_dagprop_b200.flat import FlatDists
if os.environ.get('MCDP_LIB'):  # scratch A/B builds
    capi.LIB_PATH = os.path.abspath(os.environ['MCDP_LIB'])

ap = argparse.ArgumentParser()
ap.add_argument("workload"); ap.add_argument("mode", choices=["full", "reduced"]); ap.add_argument("n", type=int)
ap.add_argument("--mt", action="store_true", help="generic-shape gamma (Marsaglia-Tsang) instead of the Erlang shapes")
ap.add_argument("--spl", type=int, default=0); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--wpg", type=int, default=0); ap.add_argument("--gpc", type=int, default=0)
ap.add_argument("--opt", action="append", default=[], help="option=value pairs passed to set_option")
a = ap.parse_args()
dag, d = getattr(synth, {"c1": "c1_toy", "c2": "c2_layered", "c3": "c3_network", "c4": "c4_national", "c5": "c5_deep_chain"}[a.workload])()
if a.mt:
    x = np.linspace(0.0, 3.0, 256)
    d = FlatDists()
    d.add_gamma(1, 2.3, 0.1, 5.0); d.add_gamma(2, 0.6, 0.3, 5.0)
    d.add_empirical_relative(3, x, np.exp(-x)); d.add_empirical_relative(4, x, np.exp(-x))
plan = capi.Plan(dag, d, device=0)
if a.spl: plan.set_option(capi.OPT_SAMPLES_PER_LANE, a.spl)
if a.wpg: plan.set_option(capi.OPT_WARPS_PER_GROUP, a.wpg)
if a.gpc: plan.set_option(capi.OPT_GROUPS_PER_CTA, a.gpc)
for kv in a.opt:
    k, v = kv.split("="); plan.set_option(int(k), int(v))
E, A, n = plan.E, plan.A, a.n
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
reduced = a.mode == "reduced"
if reduced:
    desc = capi.make_stats_desc(thresholds=(60.0, 180.0, 300.0), n_bins=64, hist_range=(0.0, dag.max_delay))
    bufs = [torch.zeros(E, dtype=torch.float64, device=dev), torch.zeros(E, dtype=torch.float64, device=dev),
            torch.zeros((3, E), dtype=torch.int64, device=dev), torch.zeros((E, 64), dtype=torch.int32, device=dev)]
    step = lambda i: plan.run_reduced_device(n, desc, *bufs, seed0=i * n, stream=st)
else:
    ld = (n + 63) // 64 * 64  # row length of the event-major arrays: a multiple of 64
    r = torch.empty((E, ld), dtype=torch.float64, device=dev); du = torch.empty((A, ld), dtype=torch.float64, device=dev)
    c = torch.empty((E, ld), dtype=torch.int32, device=dev)
    step = lambda i: plan.run_full_device(n, r, du, c, ld, seed0=i * n, stream=st)
ts = []
for i in range(a.reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(i); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
t = min(ts[1:]) * 1e-3 if len(ts) > 1 else ts[0] * 1e-3
bpe = (8 + 8 * E / A) if reduced else (16 + 12 * E / A)
sh = plan.launch_shape(n, reduced, 64 if reduced else 0)
print(f"{a.workload} {a.mode}{' mt' if a.mt else ''} n={n} shape={sh['samples_per_lane']}/{sh['warps_per_group']}/{sh['groups_per_cta']} grid={sh['grid']} "
      f"{t*1e3:.2f} ms {n*A/t:.3e} es/s {n*A*bpe/t/1e9:.0f} GB/s ({n*A*bpe/t/6545e9*100:.1f}%)", flush=True)
