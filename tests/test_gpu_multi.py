"""Multi-GPU tests (``-m gpu``; skipped on a single-GPU box): seed sharding over two ranks with NCCL
reproduces the single-GPU statistics exactly for the integer accumulators and to summation order
for the f64 sums; full outputs per seed are identical whichever rank computes them."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from mc_dagprop_b200 import capi, multi, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dag, dists = synth.random_dag(300, 8, max_delay=60.0), synth.mixed_small_dists()
        plan = capi.Plan(dag, dists, device=rank)
        seeds = np.arange(11, 11 + 5000, dtype=np.int32)
        th = (1.0, 10.0)
        st = multi.run_reduced_sharded(plan, seeds, thresholds=th, n_bins=12, hist_range=(0.0, 60.0))
        mine = multi.shard_seeds(seeds, rank, world)
        r, d, c = plan.run_many_host(mine)
        q.put((rank, st.sum, st.sumsq, st.late, st.hist, mine, r, d, c))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharding_matches_single_gpu():
    import torch

    from mc_dagprop_b200 import capi, synth

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    [p.join(timeout=60) for p in procs]
    dag, dists = synth.random_dag(300, 8, max_delay=60.0), synth.mixed_small_dists()
    plan = capi.Plan(dag, dists, device=0)
    seeds = np.arange(11, 11 + 5000, dtype=np.int32)
    ref = plan.run_reduced_host(seeds, thresholds=(1.0, 10.0), n_bins=12, hist_range=(0.0, 60.0))
    r, d, c = plan.run_many_host(seeds)
    for rank, s_sum, s_sq, late, hist, mine, rr, dd, cc in res:
        np.testing.assert_allclose(s_sum, ref.sum, rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(s_sq, ref.sumsq, rtol=1e-12, atol=1e-9)
        assert np.array_equal(late, ref.late) and np.array_equal(hist, ref.hist)
        lo = int(np.searchsorted(seeds, mine[0]))
        assert np.array_equal(rr.view(np.uint64), r[lo:lo + mine.size].view(np.uint64))
        assert np.array_equal(dd.view(np.uint64), d[lo:lo + mine.size].view(np.uint64))
        assert np.array_equal(cc, c[lo:lo + mine.size])
