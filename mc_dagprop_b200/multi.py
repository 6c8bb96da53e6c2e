"""Multi-GPU host logic: one process per GPU (``torch.distributed``), seeds sharded by rank.

Samples are independent (reference ``run_many`` is a loop over seeds, ``_core.cpp:355-361``) and the
device generator is keyed on (seed, activity), so partitioning the seeds over ranks changes
nothing per sample and needs NO data-path collective in full-output mode.  The reduced
(statistics) mode has one real exchange step: a sum all-reduce of the per-event accumulators
(NCCL over NVLink on the GPUs; gloo in the CPU tests of this logic).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int, align: int = 64) -> tuple[int, int]:
    """Contiguous block of ``range(n)`` owned by ``rank``: blocks are multiples of ``align`` samples
    (one warp's sample group) except the last, so the union is exactly ``range(n)``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    groups = (n + align - 1) // align
    per, extra = divmod(groups, world)
    g0 = rank * per + min(rank, extra)
    g1 = g0 + per + (1 if rank < extra else 0)
    return min(g0 * align, n), min(g1 * align, n)


def shard_seeds(seeds, rank: int, world: int) -> np.ndarray:
    seeds = np.asarray(seeds, dtype=np.int32).reshape(-1)
    lo, hi = shard_bounds(seeds.size, rank, world)
    return seeds[lo:hi]


def allreduce_stats(tensors, group=None) -> None:
    """In-place sum over ranks of the statistics accumulators (f64 sums, integer counters and
    histograms).  Integer buffers reduce exactly; f64 sums differ from a single-GPU run only by
    summation order."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        if t is not None and t.numel():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def run_reduced_sharded(plan, seeds, thresholds=(), n_bins=0, hist_range=(0.0, 1.0), group=None):
    """All ranks call this with the SAME ``seeds``; each runs its shard on its own GPU through the
    C ABI, then the accumulators are all-reduced.  Returns a :class:`mc_dagprop_b200.capi.Stats`
    holding the global statistics on every rank."""
    import torch
    import torch.distributed as dist

    from . import capi

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    seeds = np.asarray(seeds, dtype=np.int32).reshape(-1)
    mine = shard_seeds(seeds, rank, world)
    dev = torch.device("cuda", plan.device)
    E = plan.E
    nt = len(thresholds)
    s_sum = torch.zeros(E, dtype=torch.float64, device=dev)
    s_sq = torch.zeros(E, dtype=torch.float64, device=dev)
    s_late = torch.zeros((nt, E), dtype=torch.int64, device=dev)
    s_hist = torch.zeros((E, n_bins), dtype=torch.int32, device=dev)
    d_seeds = torch.from_numpy(mine.copy()).to(dev)
    desc = capi.make_stats_desc(thresholds, n_bins, hist_range)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        plan.run_reduced_device(mine.size, desc, s_sum, s_sq, s_late if nt else None, s_hist if n_bins else None,
                                seeds=d_seeds if mine.size else None, stream=stream)
        allreduce_stats([s_sum, s_sq, s_late, s_hist], group)
        torch.cuda.synchronize()
    return capi.Stats(seeds.size, s_sum.cpu().numpy(), s_sq.cpu().numpy(), s_late.cpu().numpy().astype(np.uint64),
                      s_hist.cpu().numpy().astype(np.uint32), tuple(thresholds), tuple(hist_range))
