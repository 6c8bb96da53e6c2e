"""Scratch A/B round 3: resident warps of the quad kernel (MCDP_LIB selects a build with another CTA limit)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_dagprop_b200 import synth
from scripts.quick_bench import run, run_reduced

wpg = int(sys.argv[1])
dag, d = synth.c3_network()
run("c3 default", dag, d, 18944, spl=4, wpg=wpg, gpc=1)
run("c3 default", dag, d, 18944, spl=4, wpg=16, gpc=1)
dag2, d2 = synth.c2_layered()
run("c2", dag2, d2, 262144, spl=4, wpg=wpg // 2, gpc=2)
