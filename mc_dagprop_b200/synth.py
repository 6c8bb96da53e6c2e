"""Deterministic synthetic DAGs for the five BASELINE.json configurations.

All generators return ``(FlatDag, FlatDists)`` built directly as flat arrays (no per-object
conversion), seeded with ``numpy.random.default_rng``.  Times are seconds.
"""
from __future__ import annotations

import numpy as np

from .flat import FlatDag, FlatDists


def _discretised_exponential(scale: float, stop: int = 200) -> np.ndarray:
    """Probabilities of the discretised exponential the reference demo uses
    (``analytic/distributions.py:77-102`` with step 1, start 0): diff of the CDF at integer edges."""
    edges = np.arange(0, stop + 1, 1)
    cdf = 1.0 - np.exp(-edges / scale)
    d = np.diff(cdf)
    return d / d.sum()


def c1_toy(variant: str = "empirical"):
    """Config 1: the 10-event / 12-activity toy DAG of ``demo/_shared.py:20-76`` driven as in
    ``demo/monte_carlo.py:21-44`` (every activity its own type, ``minimal_duration`` 0,
    ``max_delay`` 1800).  ``variant="empirical"`` is what the shipped demo samples (200-point
    discretised exponentials); ``variant="const_exp"`` is BASELINE.json's wording: constant on four
    activities, exponential on eight."""
    earliest = [0.0, 2.0, 4.0, 6.0, 7.0, 10.0, 8.0, 11.0, 12.0, 14.0]
    edges = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (1, 6), (6, 7), (7, 5), (2, 8), (8, 9), (6, 8), (4, 9)]
    scales = [2.0, 3.0, 4.0, 2.0, 3.5, 3.5, 2.5, 4.5, 5.0, 2.0, 2.5, 3.0]
    prec = [(1, [(0, 0)]), (2, [(1, 1)]), (3, [(2, 2)]), (4, [(3, 3)]), (5, [(4, 4), (7, 7)]), (6, [(1, 5)]),
            (7, [(6, 6)]), (8, [(2, 8), (6, 10)]), (9, [(8, 9), (4, 11)])]
    dists = FlatDists()
    if variant == "empirical":
        acts = [(i, 0.0, i) for i in range(len(edges))]
        for i, sc in enumerate(scales):
            dists.add_empirical_absolute(i, np.arange(200.0), _discretised_exponential(sc))
    elif variant == "const_exp":
        acts = [(i, max(0.0, float(earliest[d] - earliest[s])), i) for i, (s, d) in enumerate(edges)]
        for i, sc in enumerate(scales):
            if i % 3 == 0:
                dists.add_constant(i, 0.1)
            else:
                dists.add_exponential(i, sc / 10.0, 5.0)
    else:
        raise ValueError(variant)
    return FlatDag.from_precedence_list(earliest, acts, prec, 1800.0), dists


def _layered(n_layers: int, width: int, fan_in_fn, rng, n_types: int, headway: float = 120.0, back: int = 3):
    """Layered timetable DAG: event (l, j) = train j at stop l.  Every event of layer l >= 1 has the
    run/dwell predecessor (l-1, j) first, then extra predecessors (headways / connections) from
    uniformly random trains in layers l-1 .. l-back."""
    E = n_layers * width
    layer = np.repeat(np.arange(n_layers), width)
    earliest = headway * layer + rng.integers(0, 60, size=E)
    earliest = earliest.astype(np.float64)
    tgt = np.arange(width, E, dtype=np.int64)  # layers >= 1
    fan = fan_in_fn(tgt.size).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(fan)])
    P = int(off[-1])
    pred_tgt = np.repeat(tgt, fan)
    first = np.zeros(P, bool)
    first[off[:-1]] = True
    lt = pred_tgt // width
    jt = pred_tgt % width
    lb = np.minimum(rng.integers(1, back + 1, size=P), lt)  # how many layers back (>= 1, never before layer 0)
    src = np.where(first, (lt - 1) * width + jt, (lt - lb) * width + rng.integers(0, width, size=P))
    buffer = rng.integers(0, 61, size=P).astype(np.float64)
    base = np.maximum(0.0, earliest[pred_tgt] - earliest[src] - buffer)
    # run / dwell alternate along the train, the extras alternate headway / connection
    k_in = np.arange(P) - np.repeat(off[:-1], fan)
    if n_types == 4:
        act_type = np.where(first, 1 + (lt % 2), 3 + (k_in % 2))
    else:
        act_type = 1 + (np.arange(P) % n_types)
    act_idx = np.arange(P, dtype=np.int32)
    return earliest, act_idx, base, act_type.astype(np.int32), tgt.astype(np.int32), off, src.astype(np.int32)


def c2_layered(n_layers: int = 100, width: int = 100, seed: int = 20261002):
    """Config 2: layered railway timetable, 10k events / ~30k activities, fan-in 3, exponential
    delays per activity_type, ``max_delay`` 1800."""
    rng = np.random.default_rng(seed)
    e, ai, base, at, tgt, off, src = _layered(n_layers, width, lambda n: np.full(n, 3), rng, 4)
    dag = FlatDag(e, ai, base, at, tgt, off, src, ai.copy(), 1800.0)
    d = FlatDists()
    for t, lam in zip((1, 2, 3, 4), (0.05, 0.1, 0.2, 0.3)):
        d.add_exponential(t, lam, 5.0)
    return dag, d


def _mixed_dists() -> FlatDists:
    d = FlatDists()
    d.add_gamma(1, 2.0, 0.1, 5.0)
    d.add_gamma(2, 0.5, 0.3, 5.0)
    x = np.linspace(0.0, 3.0, 256)
    d.add_empirical_relative(3, x, np.exp(-x))
    d.add_empirical_relative(4, x, np.exp(-x))
    return d


def c3_network(n_layers: int = 250, width: int = 400, seed: int = 20261003):
    """Config 3: network DAG, 100k events / ~400k activities, fan-in Poisson(3)+1 clipped to 1..16,
    gamma on types 1-2 and 256-entry empirical-relative tables on types 3-4, ``max_delay`` 1800."""
    rng = np.random.default_rng(seed)
    fan = lambda n: np.clip(1 + rng.poisson(3.0, size=n), 1, 16)  # noqa: E731
    e, ai, base, at, tgt, off, src = _layered(n_layers, width, fan, rng, 4)
    return FlatDag(e, ai, base, at, tgt, off, src, ai.copy(), 1800.0), _mixed_dists()


def c4_national(n_layers: int = 1000, width: int = 1000, seed: int = 20261004):
    """Config 4: national-scale DAG, 1M events / ~4M activities, fan-in truncated geometric
    (mean 4, 1..16), same distribution mix as config 3."""
    rng = np.random.default_rng(seed)
    fan = lambda n: np.clip(rng.geometric(0.25, size=n), 1, 16)  # noqa: E731
    e, ai, base, at, tgt, off, src = _layered(n_layers, width, fan, rng, 4)
    return FlatDag(e, ai, base, at, tgt, off, src, ai.copy(), 1800.0), _mixed_dists()


def c5_deep_chain(chain: int = 50_000, merges: int = 200, merge_fan_in: int = 256, seed: int = 20261005):
    """Config 5: a ``chain``-event serial chain plus ``merges`` merge nodes with fan-in
    ``merge_fan_in`` drawn from random chain positions, plus a final sink over the merge nodes and
    the chain end; one 256-entry empirical-absolute table on every edge, ``minimal_duration`` 30."""
    rng = np.random.default_rng(seed)
    E = chain + merges + 1
    earliest = np.empty(E, np.float64)
    earliest[:chain] = 35.0 * np.arange(chain)
    tgt, off, src = [], [0], []
    # chain
    tgt.extend(range(1, chain))
    src.extend(range(0, chain - 1))
    off.extend(range(1, chain))
    for m in range(merges):
        picks = rng.choice(chain, size=min(merge_fan_in, chain), replace=False)
        node = chain + m
        earliest[node] = 35.0 * (picks.max() + 1)
        tgt.append(node)
        src.extend(picks.tolist())
        off.append(len(src))
    sink = chain + merges
    earliest[sink] = max(earliest[:sink].max(), 35.0 * chain) + 35.0
    tgt.append(sink)
    src.extend(list(range(chain, chain + merges)) + [chain - 1])
    off.append(len(src))
    P = len(src)
    ai = np.arange(P, dtype=np.int32)
    dag = FlatDag(earliest, ai, np.full(P, 30.0), np.ones(P, np.int32), np.asarray(tgt, np.int32),
                  np.asarray(off, np.int64), np.asarray(src, np.int32), ai.copy(), 1800.0)
    d = FlatDists()
    v = np.arange(256.0)
    d.add_empirical_absolute(1, v, np.exp(-v / 30.0))
    return dag, d


def random_dag(n_events: int, seed: int, max_fan_in: int = 5, n_types: int = 6, max_delay: float = 50.0,
               tie_prone: bool = True, idx_gaps: bool = True, shuffle: bool = True):
    """Small adversarial DAG for parity tests: random order of event ids vs topology, integer-valued
    times (ties and clamps are frequent), activity idx gaps, activities shared by two entries,
    unreferenced activities, events without predecessors, types without a distribution."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n_events) if shuffle else np.arange(n_events)  # topological position -> event id
    earliest = np.empty(n_events)
    earliest[perm] = np.sort(rng.integers(0, max(2, n_events // 2) + 1, size=n_events)).astype(float)
    tgt, off, src, act = [], [0], [], []
    acts = []
    next_idx = 0
    for pos in range(1, n_events):
        if rng.random() < 0.15:
            continue  # root
        k = int(rng.integers(1, max_fan_in + 1))
        ps = rng.integers(0, pos, size=k)
        tgt.append(int(perm[pos]))
        for p in ps:
            if acts and rng.random() < 0.05:
                a = int(rng.integers(0, len(acts)))  # reuse an activity in a second entry
                idx = acts[a][0]
            else:
                if idx_gaps and rng.random() < 0.1:
                    next_idx += int(rng.integers(1, 3))
                idx = next_idx
                next_idx += 1
                base = float(rng.integers(0, 30)) if tie_prone else float(rng.random() * 30)
                acts.append((idx, base, int(rng.integers(0, n_types + 1))))
            src.append(int(perm[p]))
            act.append(idx)
        off.append(len(src))
    for _ in range(3):  # unreferenced activities
        acts.append((next_idx, float(rng.integers(0, 30)), int(rng.integers(0, n_types + 1))))
        next_idx += 1
    order = rng.permutation(len(tgt)) if shuffle else np.arange(len(tgt))
    tgt2, off2, src2, act2 = [], [0], [], []
    for i in order:
        tgt2.append(tgt[i])
        src2.extend(src[off[i]:off[i + 1]])
        act2.extend(act[off[i]:off[i + 1]])
        off2.append(len(src2))
    dag = FlatDag(earliest, [a[0] for a in acts], [a[1] for a in acts], [a[2] for a in acts], tgt2, off2, src2, act2,
                  max_delay)
    return dag


def mixed_small_dists(n_types: int = 6) -> FlatDists:
    """One distribution of every kind on types 1..5 (type 0 and types > 5 have none)."""
    d = FlatDists()
    d.add_constant(1, 0.5)
    d.add_exponential(2, 0.7, 2.0)
    d.add_gamma(3, 2.5, 0.4, 4.0)
    d.add_empirical_absolute(4, [0.0, 1.0, 2.0, 5.0, 9.0], [0.3, 0.3, 0.2, 0.15, 0.05])
    d.add_empirical_relative(5, np.linspace(0, 2, 37), np.exp(-np.linspace(0, 2, 37)))
    if n_types >= 6:
        d.add_gamma(6, 0.6, 0.5, 3.0)
    return d
