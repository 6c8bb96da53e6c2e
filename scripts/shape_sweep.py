"""Scratch: launch-shape sweep (warps per group x groups per CTA) for one workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_dagprop_b200 import synth
from scripts.quick_bench import run

which = sys.argv[1]
n = int(sys.argv[2])
dag, d = getattr(synth, which)()
for wpg, gpc in ((0, 0), (16, 1), (8, 1), (8, 2), (4, 2), (4, 4), (2, 8), (1, 8), (1, 16)):
    try:
        run(which, dag, d, n, wpg, gpc)
    except Exception as e:  # noqa: BLE001
        print(which, wpg, gpc, "failed:", e)
