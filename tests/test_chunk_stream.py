"""Host logic of the plan compiler's chunk stream (CPU only; host-only plans, ``MCDP_DEVICE_NONE``).

The walker below is a Python restatement of how ``chunk_sweep_kernel`` consumes the stream
(``csrc/mcdp_chunk_sweep.cuh``): units in order, a header closes the running event and opens the next, every
entry applies the max-plus recurrence, the realized row an entry gathers is requested one unit early (or taken
over from the event just closed).  It models memory order: a row may only be requested after it was written.
Propagating the oracle's durations through the walked stream must reproduce the oracle bit for bit, for the
level-aligned and the dense packing, with event-id rows and with recycled scratch-slot rows."""
import numpy as np
import pytest

import oracle
from mc_dagprop_b200 import capi, synth
from mc_dagprop_b200.flat import FlatDag

U = capi.CHUNK_UNITS


def _walk(plan, dag, durations, rows, dense):
    """Returns realized[n, E], cause[n, E] computed by walking the chunk stream like one warp would (dense) or, for
    the level-aligned stream, like warps that may take the chunks of a level in ANY order."""
    units, clb = plan.chunks(rows=rows, dense=dense)
    n = durations.shape[0]
    E = plan.E
    n_rows = plan.n_slots if rows else E
    store = np.full((n_rows, n), np.nan)        # realized rows as the kernel's global memory
    written = np.zeros(n_rows, bool)
    realized = np.full((n, E), np.nan)
    cause = np.full((n, E), -2, np.int64)
    seen_events, seen_entries = set(), 0
    rng = np.random.default_rng(1)
    order = []
    for l in range(len(clb) - 1):
        cs = list(range(clb[l], clb[l + 1]))
        if not dense:
            # chunks of a level are independent work items, except that a continuation chunk follows its opener
            heads = [c for c in cs if (units[c, 0]["meta"] >> 29) == capi.KIND_EVENT]
            rng.shuffle(heads)
            cs2 = []
            for c in heads:
                cs2.append(c)
                k = c
                while units[k, 0]["c"] > 0:  # remaining continuation chunks
                    k += 1
                    assert (units[k, 0]["meta"] >> 29) == capi.KIND_END and (units[k, 0]["meta"] & 1)
                    cs2.append(k)
            assert sorted(cs2) == cs
            cs = cs2
        order.append(cs)
    state = None  # the running event: [row, event, latest, ub, cause]
    nrs = None    # value requested one unit early

    def close():
        nonlocal state
        row, ev, latest, ub, cz = state
        r = np.minimum(latest, ub)
        store[row] = r
        written[row] = True
        realized[:, ev] = r
        cause[:, ev] = cz
        state = None
        return r

    last_closed = None
    for level_chunks in order:
        level_rows_written = []
        for c in level_chunks:
            ch = units[c]
            is_cont = (ch[0]["meta"] >> 29) == capi.KIND_END and (ch[0]["meta"] & 1)
            assert (ch[0]["meta"] >> 29) == capi.KIND_EVENT or is_cont
            remaining = int(ch[0]["c"])
            for u in range(1 if is_cont else 0, U):
                rec = ch[u]
                kind = rec["meta"] >> 29
                if kind >= capi.KIND_EVENT:
                    if dense:  # the close precedes the header's own request
                        if state is not None:
                            last_closed = close()
                    else:      # level-aligned: the request may be issued before the close (no dependency inside a level)
                        if rec["nxt"] != capi.NO_ROW and kind == capi.KIND_EVENT:
                            assert written[rec["nxt"]], "header gathers a row that is not written yet"
                            nrs = store[rec["nxt"]].copy()
                        if state is not None:
                            last_closed = close()
                    if kind == capi.KIND_END:
                        break
                    if dense:
                        if rec["d"]:
                            assert rec["nxt"] == capi.NO_ROW and last_closed is not None
                            nrs = last_closed.copy()
                        elif rec["nxt"] != capi.NO_ROW:
                            assert written[rec["nxt"]], "header gathers a row that is not written yet"
                            nrs = store[rec["nxt"]].copy()
                    else:
                        assert rec["d"] == 0
                    ev = int(rec["b"])
                    assert ev not in seen_events
                    seen_events.add(ev)
                    e0 = rec["x"]
                    assert e0 == dag.earliest[ev] or (np.isnan(e0) and np.isnan(dag.earliest[ev]))
                    state = [int(rec["a"]), ev, np.full(n, e0), e0 + dag.max_delay, np.full(n, -1, np.int64)]
                    continue
                # entry unit
                seen_entries += 1
                rs = nrs
                nrs = None
                assert rs is not None, "entry whose source row was never requested"
                if rec["nxt"] != capi.NO_ROW:
                    assert written[rec["nxt"]], "entry requests a row that is not written yet"
                    nrs = store[rec["nxt"]].copy()
                act = rec["b"]
                d = durations[:, act] if act != capi.NO_ACT else 0.0
                assert np.array_equal(rs, store[rec["a"]], equal_nan=True), "prefetched value is not the source row's"
                t = np.minimum(rs + d, state[3])
                take = t >= state[2]
                state[2] = np.where(take, t, state[2])
                # cause_event reports event ids; with slot rows the walker maps back through the open events
                state[4] = np.where(take, int(rec["a"]), state[4])
            if state is not None and remaining == 0:
                last_closed = close()
            if not dense:
                assert state is None or remaining > 0
    assert state is None and seen_events == set(range(E)) and seen_entries == plan.P
    return realized, cause


@pytest.mark.parametrize("n_events,seed", [(1, 0), (40, 1), (400, 2), (1500, 3)])
@pytest.mark.parametrize("dense", [False, True])
def test_walking_the_chunk_stream_reproduces_the_oracle(n_events, seed, dense):
    dag = synth.random_dag(n_events, seed)
    dists = synth.mixed_small_dists()
    osim = oracle.OracleSim(dag, dists)
    _, dur, _ = osim.run_many(np.arange(5, dtype=np.int32))
    r_o, c_o = osim.run_injected(dur)
    plan = capi.Plan(dag, dists, device=capi.DEVICE_NONE)
    r, c = _walk(plan, dag, dur, rows=0, dense=dense)
    assert np.array_equal(r.view(np.uint64), r_o.view(np.uint64))
    assert np.array_equal(c, c_o)
    r2, _ = _walk(plan, dag, dur, rows=1, dense=dense)  # recycled scratch-slot rows: same realized times
    assert np.array_equal(r2.view(np.uint64), r_o.view(np.uint64))


def _chain_with_merges(chain=300, fan=40):
    """A serial chain (every event reads the one closed just before it) plus one merge node over `fan` chain events
    (a long event: continuation chunks) plus a sink: the shapes the dense packing exists for."""
    E = chain + 2
    earliest = np.concatenate([7.0 * np.arange(chain), [7.0 * chain, 7.0 * chain + 5.0]])
    acts, prec = [], []
    for i in range(1, chain):
        acts.append((len(acts), 6.5, 1))
        prec.append((i, [(i - 1, len(acts) - 1)]))
    merge_preds = []
    for s in np.linspace(0, chain - 1, fan).astype(int):
        acts.append((len(acts), 3.0, 2))
        merge_preds.append((int(s), len(acts) - 1))
    prec.append((chain, merge_preds))
    acts.append((len(acts), 1.0, 1))
    acts.append((len(acts), 2.0, 2))
    prec.append((chain + 1, [(chain, len(acts) - 2), (chain - 1, len(acts) - 1)]))
    return FlatDag.from_precedence_list(earliest, acts, prec, 50.0)


@pytest.mark.parametrize("dense", [False, True])
def test_chain_forwarding_and_continuation_chunks(dense):
    dag = _chain_with_merges()
    dists = synth.mixed_small_dists()
    osim = oracle.OracleSim(dag, dists)
    _, dur, _ = osim.run_many(np.arange(4, dtype=np.int32))
    r_o, c_o = osim.run_injected(dur)
    plan = capi.Plan(dag, dists, device=capi.DEVICE_NONE)
    units, clb = plan.chunks(rows=0, dense=dense)
    kinds = units["meta"] >> 29
    assert ((kinds[:, 0] == capi.KIND_END) & ((units["meta"][:, 0] & 1) == 1)).sum() >= 2  # continuation chunks exist
    if dense:
        assert units["d"][kinds == capi.KIND_EVENT].sum() > 200  # the chain forwards from registers
        assert clb.tolist() == [0, units.shape[0]]
    else:
        assert len(clb) == plan.n_levels + 1 and clb[-1] == units.shape[0]
    for rows in (0, 1):
        r, c = _walk(plan, dag, dur, rows=rows, dense=dense)
        assert np.array_equal(r.view(np.uint64), r_o.view(np.uint64))
        if rows == 0:
            assert np.array_equal(c, c_o)
