"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: seed sharding covers the seed list
exactly once, and the statistics all-reduce reproduces the unsharded accumulators.  The per-shard
accumulators are produced by the CPU oracle here (test infrastructure); on the GPUs the same
functions are fed by mcdp_run_reduced_device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from mc_dagprop_b200 import multi, synth


def test_shard_bounds_partition():
    for n in (0, 1, 63, 64, 65, 1000, 4096, 100_003):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi = multi.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n and (lo % 64 == 0 or lo == n)
                covered.extend(range(lo, hi))
            assert covered == list(range(n))
    with pytest.raises(ValueError):
        multi.shard_bounds(10, 2, 2)


def _local_stats(dag, dists, seeds, thresholds, n_bins, hist_hi):
    # the device generator contract is a pure function of (seed, activity); the reference stream is not
    # (gamma's cached normal survives reseeding), so sharding equivalence is stated on the contract
    r, _, _ = oracle.OracleSim(dag, dists).run_many_spec(seeds, durations=False, cause=False)
    delay = r - dag.earliest[None, :]
    E = dag.n_events
    s = torch.from_numpy(delay.sum(0)) if seeds.size else torch.zeros(E, dtype=torch.float64)
    q = torch.from_numpy((delay * delay).sum(0)) if seeds.size else torch.zeros(E, dtype=torch.float64)
    late = torch.from_numpy(np.stack([(delay > t).sum(0) for t in thresholds]).astype(np.int64))
    bins = np.clip(np.floor(delay * (n_bins / hist_hi)).astype(int), 0, n_bins - 1)
    hist = np.zeros((E, n_bins), np.int32)
    for e in range(E):
        hist[e] = np.bincount(bins[:, e], minlength=n_bins)
    return s, q, late, torch.from_numpy(hist)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dag, dists = synth.random_dag(60, 5, max_delay=40.0), synth.mixed_small_dists()
        seeds = np.arange(7, 7 + 777, dtype=np.int32)
        th, nb, hi = (1.0, 10.0), 8, 40.0
        mine = multi.shard_seeds(seeds, rank, world)
        bufs = list(_local_stats(dag, dists, mine, th, nb, hi))
        multi.allreduce_stats(bufs)
        full = _local_stats(dag, dists, seeds, th, nb, hi)
        ok = (torch.allclose(bufs[0], full[0], rtol=1e-12, atol=1e-9) and torch.allclose(bufs[1], full[1], rtol=1e-12, atol=1e-9)
              and torch.equal(bufs[2], full[2]) and torch.equal(bufs[3], full[3]))
        q.put((rank, bool(ok), int(mine.size)))
    finally:
        dist.destroy_process_group()


def test_stats_allreduce_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res)
    assert sum(r[2] for r in res) == 777
