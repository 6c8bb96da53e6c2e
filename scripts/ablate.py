"""Scratch ablation on the C3 DAG: cost of each distribution kind (device buffers, CUDA events)."""
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


def all_types(name, add):
    g = FlatDists()
    for t in (1, 2, 3, 4):
        add(g, t)
    run(name, dag, g, n)


all_types("c3 all-empirical", lambda g, t: g.add_empirical_relative(t, x, np.exp(-x)))
all_types("c3 all-gamma2.0", lambda g, t: g.add_gamma(t, 2.0, 0.1, 5.0))
all_types("c3 all-gamma0.5", lambda g, t: g.add_gamma(t, 0.5, 0.3, 5.0))
all_types("c3 all-gamma2.3", lambda g, t: g.add_gamma(t, 2.3, 0.1, 5.0))
all_types("c3 all-gamma0.6", lambda g, t: g.add_gamma(t, 0.6, 0.3, 5.0))
all_types("c3 all-exponential", lambda g, t: g.add_exponential(t, 0.2, 5.0))
all_types("c3 all-constant", lambda g, t: g.add_constant(t, 0.1))
g = FlatDists()
g.add_gamma(1, 2.3, 0.1, 5.0); g.add_gamma(2, 0.6, 0.3, 5.0)
g.add_empirical_relative(3, x, np.exp(-x)); g.add_empirical_relative(4, x, np.exp(-x))
run("c3 mix generic-shape gamma", dag, g, n)
