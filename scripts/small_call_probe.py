"""Device time of one full-output launch over n samples on configuration 3 through the small-call path
(mcdp_small_sweep.cuh) and through the sweep kernels, n = 1 ... 512: the data behind the auto rule (kSmallAutoMax in
mcdp_capi.cu).  python scripts/small_call_probe.py [workload]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_workload  # noqa: E402
from mc_dagprop_b200 import capi  # noqa: E402

if os.environ.get("MCDP_LIB"):  # scratch A/B builds (scripts/ab_variants.sh)
    capi.LIB_PATH = os.path.abspath(os.environ["MCDP_LIB"])
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
dag, dists = build_workload(wl)
plan = capi.Plan(dag, dists, device=0)
E, A = plan.E, plan.A


def med(f, reps=15):
    f()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


print(f"{wl}: E={E} A={A}; device time of one launch, ms (median of 15)")
print(f"{'samples':>8} {'small-call path':>16} {'sweep kernels':>14}")
for n in (1, 2, 4, 8, 16, 32, 64, 128, 148, 256, 512):
    ld = 128 * ((n + 127) // 128)
    r = torch.empty((E, ld), dtype=torch.float64, device="cuda")
    d = torch.empty((A, ld), dtype=torch.float64, device="cuda")
    c = torch.empty((E, ld), dtype=torch.int32, device="cuda")
    out = []
    for limit in (4096, 0):
        plan.set_option(capi.OPT_SMALL_CALL_MAX, limit)
        out.append(med(lambda: plan.run_full_device(n, r, d, c, ld, seed0=0)))
    print(f"{n:>8} {out[0]:>16.3f} {out[1]:>14.3f}")
