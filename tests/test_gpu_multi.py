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


def _solo_worker(port, q):
    """world_size 1 over NCCL on cuda:0: the all-reduce path of the reduced mode on a single-GPU box."""
    import torch
    import torch.distributed as dist

    from mc_dagprop_b200 import capi, multi, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        dag, dists = synth.random_dag(300, 8, max_delay=60.0), synth.mixed_small_dists()
        plan = capi.Plan(dag, dists, device=0)
        seeds = np.arange(11, 11 + 5000, dtype=np.int32)
        # two shards of the same seeds on the one device, accumulated into the same buffers, then the NCCL all-reduce
        dev = torch.device("cuda", 0)
        E = plan.E
        s_sum = torch.zeros(E, dtype=torch.float64, device=dev)
        s_sq = torch.zeros(E, dtype=torch.float64, device=dev)
        s_late = torch.zeros((2, E), dtype=torch.int64, device=dev)
        s_hist = torch.zeros((E, 12), dtype=torch.int32, device=dev)
        desc = capi.make_stats_desc((1.0, 10.0), 12, (0.0, 60.0))
        for r in range(2):
            mine = multi.shard_seeds(seeds, r, 2)
            d_seeds = torch.from_numpy(mine.copy()).to(dev)
            plan.run_reduced_device(mine.size, desc, s_sum, s_sq, s_late, s_hist, seeds=d_seeds,
                                    stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        multi.allreduce_stats([s_sum, s_sq, s_late, s_hist])
        torch.cuda.synchronize()
        st = multi.run_reduced_sharded(plan, seeds, thresholds=(1.0, 10.0), n_bins=12, hist_range=(0.0, 60.0))
        q.put((s_sum.cpu().numpy(), s_sq.cpu().numpy(), s_late.cpu().numpy(), s_hist.cpu().numpy(), st.sum, st.late, st.hist))
    finally:
        dist.destroy_process_group()


def test_nccl_allreduce_path_on_one_gpu():
    """The driver's test box has one GPU: run the NCCL leg there too (one rank; two seed shards folded on the
    one device), against the plain single-call statistics."""
    import torch.multiprocessing as mp

    from mc_dagprop_b200 import capi, synth

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_solo_worker, args=(port, q))
    p.start()
    s_sum, s_sq, s_late, s_hist, st_sum, st_late, st_hist = q.get(timeout=300)
    p.join(timeout=60)
    dag, dists = synth.random_dag(300, 8, max_delay=60.0), synth.mixed_small_dists()
    plan = capi.Plan(dag, dists, device=0)
    seeds = np.arange(11, 11 + 5000, dtype=np.int32)
    ref = plan.run_reduced_host(seeds, thresholds=(1.0, 10.0), n_bins=12, hist_range=(0.0, 60.0))
    np.testing.assert_allclose(s_sum, ref.sum, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(s_sq, ref.sumsq, rtol=1e-12, atol=1e-9)
    assert np.array_equal(s_late.astype(np.uint64), ref.late) and np.array_equal(s_hist.astype(np.uint32), ref.hist)
    np.testing.assert_allclose(st_sum, ref.sum, rtol=1e-12, atol=1e-9)
    assert np.array_equal(st_late, ref.late) and np.array_equal(st_hist, ref.hist)
