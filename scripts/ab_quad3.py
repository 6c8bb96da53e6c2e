"""Scratch A/B round 3: the quad kernel at its auto launch shapes (MCDP_LIB selects another build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mc_dagprop_b200 import synth
from mc_dagprop_b200.flat import FlatDists
from scripts.quick_bench import run, run_reduced

dag, d = synth.c3_network()
run("c3 default", dag, d, 18944)
run("c3 none", dag, FlatDists(), 18944)
x = np.linspace(0.0, 3.0, 256)
g = FlatDists()
g.add_gamma(1, 2.3, 0.1, 5.0); g.add_gamma(2, 0.6, 0.3, 5.0)
g.add_empirical_relative(3, x, np.exp(-x)); g.add_empirical_relative(4, x, np.exp(-x))
run("c3 generic-shape gamma", dag, g, 18944)
run_reduced("c3", dag, d, 18944)
dag2, d2 = synth.c2_layered()
run("c2", dag2, d2, 262144)
if "c4" in sys.argv:
    dag4, d4 = synth.c4_national()
    run_reduced("c4", dag4, d4, 1 << 15)
