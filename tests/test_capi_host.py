"""CPU tests of the C-ABI library: it loads, exports every symbol include/mcdp_b200.h declares,
and the host plan compiler (levels, order, tables, validation) behaves -- no compute calls."""
import os
import re

import numpy as np
import pytest

import oracle
from mc_dagprop_b200 import capi, synth
from mc_dagprop_b200.flat import FlatDag, FlatDists

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mcdp_b200.h")).read()
    declared = set(re.findall(r"\b(mcdp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    lib = capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mcdp_abi_version() == 5


def test_no_cpu_execution_path():
    dag, d = synth.c1_toy()
    plan = capi.Plan(dag, d, device=capi.DEVICE_NONE)
    with pytest.raises(RuntimeError, match="no CPU execution path"):
        plan.run_many_host([1, 2, 3])
    with pytest.raises(RuntimeError, match="no CPU execution path"):
        plan.run_reduced_host([1, 2, 3])
    if capi.device_count() == 0:
        with pytest.raises(RuntimeError, match="CUDA"):
            capi.Plan(dag, d, device=0)


@pytest.mark.parametrize("n,seed", [(1, 0), (50, 1), (700, 2)])
def test_plan_order_is_topological_and_levelled(n, seed):
    dag = synth.random_dag(n, seed)
    plan = capi.Plan(dag, synth.mixed_small_dists(), device=capi.DEVICE_NONE)
    osim = oracle.OracleSim(dag, synth.mixed_small_dists())
    assert (plan.E, plan.A, plan.P) == (osim.E, osim.A, len(osim.csr()[1]))
    order, level = plan.order()
    assert sorted(order.tolist()) == list(range(plan.E))
    pos = np.empty(plan.E, int)
    pos[order] = np.arange(plan.E)
    lvl = np.empty(plan.E, int)
    lvl[order] = level
    off, src, _ = osim.csr()
    for e in range(plan.E):
        preds = src[off[e]:off[e + 1]]
        assert all(pos[p] < pos[e] for p in preds)
        # longest-path layering: level = 1 + max level of the predecessors (0 for roots)
        assert lvl[e] == (1 + max(lvl[p] for p in preds) if len(preds) else 0)
    assert np.all(np.diff(level) >= 0) and plan.n_levels == (level.max() + 1 if plan.E else 0)


def test_plan_cumulative_tables_equal_libstdcpp_discrete_distribution():
    d = FlatDists()
    rng = np.random.default_rng(3)
    w = rng.random(257)
    d.add_empirical_absolute(1, np.arange(257.0), w)
    d.add_empirical_relative(2, [1.0, 2.0], [3.0, 1.0])
    d.add_empirical_absolute(3, [7.0], [2.0])
    dag = FlatDag.from_precedence_list([0.0, 1.0], [(0, 1.0, 1), (1, 1.0, 2), (2, 1.0, 3)], [(1, [(0, 0)])], 5.0)
    plan = capi.Plan(dag, d, device=capi.DEVICE_NONE)
    osim = oracle.OracleSim(dag, d)
    for t in (1, 2):
        assert np.array_equal(plan.cumulative(t).view(np.uint64), osim.cumulative(t).view(np.uint64))
    assert plan.cumulative(3).tolist() == [1.0]  # single entry: index 0 without a draw
    assert plan.cumulative(99) is None


def test_validation_messages_match_reference():
    d = FlatDists()
    d.add_constant(1, 0.0)
    cyc = FlatDag.from_precedence_list([0.0, 0.0], [(0, 1.0, 1), (1, 1.0, 1)], [(1, [(0, 0)]), (0, [(1, 1)])], 1e6)
    with pytest.raises(RuntimeError, match="Invalid DAG: cycle detected in precedence list"):
        capi.Plan(cyc, d, device=capi.DEVICE_NONE)
    ok = FlatDag.from_precedence_list([0.0, 0.0], [(0, 1.0, 1)], [(1, [(0, 0)])], -1.0)
    with pytest.raises(RuntimeError, match="max_delay must be non-negative"):
        capi.Plan(ok, d, device=capi.DEVICE_NONE)
    d2 = FlatDists()
    d2.add_constant(-1, 0.0)
    ok.max_delay = 1.0
    with pytest.raises(RuntimeError, match="Activity type -1 is reserved for no delay"):
        capi.Plan(ok, d2, device=capi.DEVICE_NONE)


def test_validation_of_what_the_reference_leaves_undefined():
    d = FlatDists()
    for prec, msg in [([(5, [(0, 0)])], "target"), ([(1, [(7, 0)])], "predecessor"), ([(1, [(0, -2)])], "activity")]:
        dag = FlatDag.from_precedence_list([0.0, 0.0], [(0, 1.0, 1)], prec, 1.0)
        with pytest.raises(RuntimeError, match=msg):
            capi.Plan(dag, d, device=capi.DEVICE_NONE)
    dag = FlatDag.from_precedence_list([0.0, 0.0], [(0, 1.0, 1)], [(1, [(0, 0)])], 1.0)
    for bad in (lambda g: g.add_exponential(1, -1.0, 1.0), lambda g: g.add_exponential(1, 1.0, -1.0),
                lambda g: g.add_gamma(1, 0.0, 1.0), lambda g: g.add_empirical_absolute(1, [], []),
                lambda g: g.add_empirical_absolute(1, [1.0, 2.0], [0.0, 0.0]),
                lambda g: g.add_empirical_relative(1, [1.0, 2.0], [1.0, -1.0])):
        g = FlatDists()
        bad(g)
        with pytest.raises(RuntimeError):
            capi.Plan(dag, g, device=capi.DEVICE_NONE)
    with pytest.raises(RuntimeError, match="same length"):
        FlatDists().add_empirical_absolute(1, [1.0, 2.0], [1.0])


def test_activity_index_past_the_end_is_a_zero_duration_link():
    """The reference reads past actual_durations_ here (UB); its LargeScaleTest relies on 0.0."""
    dag = FlatDag.from_precedence_list([0.0, 1.0], [(0, 1.0, 1)], [(1, [(0, 9)])], 10.0)
    plan = capi.Plan(dag, FlatDists(), device=capi.DEVICE_NONE)
    assert plan.A == 1 and plan.P == 1
    r, d, c = oracle.OracleSim(dag, FlatDists()).run_many([0])
    assert r[0].tolist() == [0.0, 1.0] and c[0].tolist() == [-1, -1] and d.shape == (1, 1)


def test_duplicate_target_entries_last_wins():
    """reference _core.cpp:240 (preds_by_target[tgt] = entry.second): the last entry's preds win."""
    dag = FlatDag.from_precedence_list([0.0, 0.0, 0.0], [(0, 10.0, 0), (1, 50.0, 0)],
                                       [(2, [(0, 0)]), (2, [(1, 1)])], 1e6)
    plan = capi.Plan(dag, FlatDists(), device=capi.DEVICE_NONE)
    assert plan.P == 1
    r, _, c = oracle.OracleSim(dag, FlatDists()).run_many([0])
    assert r[0, 2] == 50.0 and c[0, 2] == 1


def test_benchmark_configs_compile():
    for gen, (E, A) in [(synth.c1_toy, (10, 12)), (synth.c2_layered, (10_000, 29_700))]:
        dag, d = gen()
        plan = capi.Plan(dag, d, device=capi.DEVICE_NONE)
        assert (plan.E, plan.A) == (E, A)
    dag, d = synth.c3_network()
    plan = capi.Plan(dag, d, device=capi.DEVICE_NONE)
    assert plan.E == 100_000 and 390_000 < plan.A < 410_000 and plan.n_levels == 250
    dag, d = synth.c5_deep_chain()
    plan = capi.Plan(dag, d, device=capi.DEVICE_NONE)
    assert plan.n_levels == 50_001 + 1 - 1 or plan.n_levels >= 50_000


def test_launch_shape_of_both_kernels():
    """mcdp_plan_launch_shape on a host-only plan (148 SMs assumed): the pair kernel covers 64 samples per group at
    32 resident warps per SM, the quad kernel 128 samples per group at 16; the bench workload's C3 batch is one full
    wave for either."""
    dag, d = synth.c2_layered()
    plan = capi.Plan(dag, d, device=capi.DEVICE_NONE)
    n = 18944  # 296 pair groups = 148 quad groups
    plan.set_option(capi.OPT_SAMPLES_PER_LANE, 2)
    plan.set_option(capi.OPT_WARPS_PER_GROUP, 16)
    s2 = plan.launch_shape(n)
    assert (s2["samples_per_lane"], s2["warps_per_group"], s2["threads"], s2["grid"]) == (2, 16, 512, 296)
    plan.set_option(capi.OPT_SAMPLES_PER_LANE, 4)
    s4 = plan.launch_shape(n)
    assert (s4["samples_per_lane"], s4["warps_per_group"], s4["threads"], s4["grid"]) == (4, 16, 512, 148)
    assert s4["smem_bytes"] == s2["smem_bytes"]  # same tables + one chunk ring per warp
    # ragged counts round up to whole groups
    plan.set_option(capi.OPT_WARPS_PER_GROUP, 1)
    plan.set_option(capi.OPT_GROUPS_PER_CTA, 1)
    assert plan.launch_shape(129)["grid"] == 2 and plan.launch_shape(128)["grid"] == 1 and plan.launch_shape(1)["grid"] == 1
    plan.set_option(capi.OPT_SAMPLES_PER_LANE, 2)
    assert plan.launch_shape(129)["grid"] == 3
    # reduced launches of any size are single-batch launches of the quad kernel (it stages its statistics per warp)
    plan.set_option(capi.OPT_SAMPLES_PER_LANE, 0)
    plan.set_option(capi.OPT_WARPS_PER_GROUP, 0)
    plan.set_option(capi.OPT_GROUPS_PER_CTA, 0)
    big = plan.launch_shape(1 << 20, reduced=True, n_bins=64)
    assert big["samples_per_lane"] == 4 and big["grid"] * big["groups_per_cta"] * 128 >= 1 << 20
    assert plan.launch_shape(1 << 15, reduced=True, n_bins=64)["samples_per_lane"] == 4
    full = plan.launch_shape(1 << 15, reduced=False)
    assert plan.launch_shape(1 << 15, reduced=True, n_bins=64)["smem_bytes"] > full["smem_bytes"]  # + the staging areas
    with pytest.raises(RuntimeError, match="samples per lane"):
        plan.set_option(capi.OPT_SAMPLES_PER_LANE, 3)


def test_library_is_sm_100a_code_with_bulk_copies_and_256_bit_accesses():
    """What the build claims, read from the shipped binary (no GPU needed): one sm_100a cubin, the TMA bulk copy that
    feeds the per-warp record ring (UBLKCP) with its mbarrier waits (SYNCS), and the 256-bit row loads / stores of the
    quad kernel (LDG.E...256 / STG.E...256, new with sm_100)."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elfs = subprocess.run([cuobjdump, "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elfs and not re.search(r"sm_(?!100a)\d+", elfs), elfs
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "quad_sweep_kernel", capi.LIB_PATH], capture_output=True, text=True).stdout
    if "Function" not in sass:  # older cuobjdump: no -fun filter on mangled substrings
        sass = subprocess.run([cuobjdump, "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "quad_sweep_kernel" in sass and "UBLKCP" in sass and "SYNCS" in sass
    assert re.search(r"LDG\.E[.\w]*\.256", sass) and re.search(r"STG\.E[.\w]*\.256", sass)


@pytest.mark.parametrize("gen", ["c1_toy", "c2_layered", "c5_deep_chain"])
def test_launch_shape_invariants(gen):
    """Whatever the options, a launch covers every sample, respects the CTA limits of its kernel (512 threads pair,
    640 quad), keeps its named-barrier ids within the 15 a CTA has, and fits the SM's shared memory."""
    dag, d = getattr(synth, gen)()
    plan = capi.Plan(dag, d, device=capi.DEVICE_NONE)
    rng = np.random.default_rng(1)
    ns = [1, 63, 64, 65, 127, 128, 129, 18944, 1 << 15, 1 << 18, 1 << 20] + rng.integers(1, 1 << 21, size=20).tolist()
    for spl in (0, 2, 4):
        for wpg, gpc in [(0, 0), (1, 0), (3, 0), (16, 0), (20, 0), (0, 3), (5, 2), (8, 15), (20, 15)]:
            plan.set_option(capi.OPT_SAMPLES_PER_LANE, spl)
            plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
            plan.set_option(capi.OPT_GROUPS_PER_CTA, gpc)
            for n in ns:
                for reduced in (False, True):
                    s = plan.launch_shape(n, reduced, 64 if reduced else 0)
                    k = s["samples_per_lane"]
                    if k == 1:  # the small-call path: only by the auto rule, only full-output calls of a few samples
                        assert not reduced and (spl, wpg, gpc) == (0, 0, 0) and n <= 128 and s["grid"] == n
                        continue
                    assert k in (2, 4) and (spl == 0 or k == spl)
                    assert s["threads"] == 32 * s["warps_per_group"] * s["groups_per_cta"]
                    assert s["threads"] <= (640 if k == 4 else 512) and 1 <= s["groups_per_cta"] <= 15
                    c = s["cluster"]
                    assert c in (1, 2, 4, 8) and s["grid"] % c == 0
                    groups = s["grid"] // c * s["groups_per_cta"]
                    assert groups * 32 * k >= n
                    assert (s["grid"] // c - 1) * s["groups_per_cta"] * 32 * k < n  # no empty CTA
                    assert s["smem_bytes"] <= 227 * 1024
                    if c > 1:  # a cluster shares ONE group: quad kernel, one group per CTA, several warps, few groups
                        assert k == 4 and s["groups_per_cta"] == 1 and s["warps_per_group"] > 1 and s["grid"] <= 148


def test_plan_set_compiles_once_for_host_only_devices_and_has_no_cpu_path():
    """A plan set validates and compiles without a GPU (MCDP_DEVICE_NONE entries); every run call on it fails."""
    dag, d = synth.random_dag(50, 3), synth.mixed_small_dists()
    ps = capi.PlanSet(dag, d, [capi.DEVICE_NONE, capi.DEVICE_NONE])
    assert len(ps) == 2 and ps.E == 50
    for call in (lambda: ps.run_many_host(np.arange(4, dtype=np.int32)),
                 lambda: ps.run_reduced_host(np.arange(4, dtype=np.int32), n_bins=4)):
        with pytest.raises(RuntimeError, match="no CPU execution path"):
            call()
    bad = synth.random_dag(50, 3)
    bad.max_delay = -1.0
    with pytest.raises(RuntimeError, match="max_delay must be non-negative"):
        capi.PlanSet(bad, d, [capi.DEVICE_NONE])
