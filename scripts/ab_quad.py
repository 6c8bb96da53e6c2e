"""Scratch A/B: pair kernel (2 samples per lane) vs quad kernel (4 samples per lane) on the named workloads."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mc_dagprop_b200 import synth
from mc_dagprop_b200.flat import FlatDists
from scripts.quick_bench import run, run_reduced

which = sys.argv[1:] or ["c3", "c2", "c5", "c4"]
if "c3" in which:
    dag, d = synth.c3_network()
    n = 18944
    for spl in (2, 4):
        run("c3 default", dag, d, n, spl=spl)
    for spl in (2, 4):
        run("c3 none", dag, FlatDists(), n, spl=spl)
    x = np.linspace(0.0, 3.0, 256)
    g = FlatDists()
    g.add_gamma(1, 2.3, 0.1, 5.0); g.add_gamma(2, 0.6, 0.3, 5.0)
    g.add_empirical_relative(3, x, np.exp(-x)); g.add_empirical_relative(4, x, np.exp(-x))
    for spl in (2, 4):
        run("c3 generic-shape gamma", dag, g, n, spl=spl)
    run("c3 default", dag, d, n, spl=4, wpg=8, gpc=2)
    run("c3 default n=9472", dag, d, 9472, spl=4)
    run("c3 default n=9472", dag, d, 9472, spl=2)
    for spl in (2, 4):
        run_reduced("c3", dag, d, n, spl=spl)
if "c2" in which:
    dag2, d2 = synth.c2_layered()
    for spl in (2, 4):
        run("c2", dag2, d2, 262144, spl=spl)
    run("c2", dag2, d2, 262144, spl=4, wpg=8, gpc=2)
    run("c2", dag2, d2, 262144, spl=4, wpg=4, gpc=4)
if "c1" in which:
    dag1, d1 = synth.c1_toy()
    for spl in (2, 4):
        run("c1", dag1, d1, 1 << 20, spl=spl)
if "c5" in which:
    dag5, d5 = synth.c5_deep_chain()
    for spl in (2, 4):
        run_reduced("c5", dag5, d5, 1 << 18, spl=spl)
if "c4" in which:
    dag4, d4 = synth.c4_national()
    for spl in (2, 4):
        run_reduced("c4", dag4, d4, 1 << 15, spl=spl)
    run_reduced("c4 n=37888", dag4, d4, 37888, spl=4)
