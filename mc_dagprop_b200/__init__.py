"""mc_dagprop_b200 -- B200-native (sm_100a) Monte-Carlo DAG propagation.

A from-scratch implementation of the hot path of WonJayne/mc_dagprop
(``MonteCarloPropagator.run / run_many`` over a ``DagContext`` with a ``GenericDelayGenerator``,
reference ``src/mc_dagprop/monte_carlo/_core.cpp``) behind the reference's own Python surface.

* ``mc_dagprop_b200.monte_carlo`` -- the drop-in classes (pybind11 module ``_core``)
* ``mc_dagprop_b200.capi``        -- ctypes binding of the C ABI (``include/mcdp_b200.h``)
* ``mc_dagprop_b200.flat``        -- flat-array DAG / generator descriptions
* ``mc_dagprop_b200.synth``       -- synthetic DAGs of the benchmark configurations

There is no CPU execution path: importing works without a GPU, running needs one.
"""
__version__ = "0.1.0"
