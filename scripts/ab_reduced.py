"""Scratch A/B of the reduced-statistics launches (MCDP_LIB selects the build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_dagprop_b200 import synth
from scripts.quick_bench import run_reduced

dag5, d5 = synth.c5_deep_chain()
run_reduced("c5", dag5, d5, 1 << 18, reps=4)
dag, d = synth.c3_network()
run_reduced("c3", dag, d, 18944, reps=3)
dag4, d4 = synth.c4_national()
run_reduced("c4", dag4, d4, 1 << 15, reps=2)
