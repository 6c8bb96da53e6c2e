"""Scratch A/B round 2: launch shapes of the quad kernel on C2, auto choice, C5 check."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_dagprop_b200 import synth
from scripts.quick_bench import run, run_reduced

dag2, d2 = synth.c2_layered()
run("c2 auto", dag2, d2, 262144)
run("c2", dag2, d2, 262144, spl=4, wpg=8, gpc=1)
run("c2", dag2, d2, 262144, spl=4, wpg=8, gpc=2)
run("c2", dag2, d2, 262144, spl=2, wpg=8, gpc=2)
run("c2", dag2, d2, 262144, spl=2, wpg=8, gpc=1)
dag, d = synth.c3_network()
run("c3 auto", dag, d, 18944)
run("c3 auto n=9472", dag, d, 9472)
dag5, d5 = synth.c5_deep_chain()
for spl in (2, 4):
    run_reduced("c5", dag5, d5, 1 << 18, spl=spl, reps=4)
dag1, d1 = synth.c1_toy()
for spl in (2, 4):
    run("c1", dag1, d1, 1 << 20, spl=spl)
