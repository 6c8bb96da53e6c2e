"""Scratch: the three C3 points that steer kernel work (default mix, no distributions, generic-shape mix) + C2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mc_dagprop_b200 import synth
from mc_dagprop_b200.flat import FlatDists
from scripts.quick_bench import run

dag, d = synth.c3_network()
n = 18944
run("c3 default", dag, d, n)
run("c3 none", dag, FlatDists(), n)
x = np.linspace(0.0, 3.0, 256)
g = FlatDists()
g.add_gamma(1, 2.3, 0.1, 5.0); g.add_gamma(2, 0.6, 0.3, 5.0)
g.add_empirical_relative(3, x, np.exp(-x)); g.add_empirical_relative(4, x, np.exp(-x))
run("c3 mix generic-shape gamma", dag, g, n)
dag2, d2 = synth.c2_layered()
run("c2", dag2, d2, 262144)
