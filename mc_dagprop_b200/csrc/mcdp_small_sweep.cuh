// mcdp_small_sweep.cuh -- the path of calls with a handful of samples (run(seed), reference _core.cpp:312-353).
//
// The sweep kernels give a lane two or four SAMPLES and walk the activities of a level one after the other; with one
// sample in the call, 127 of a group's 128 sample columns are padding and the time of the launch is the length of
// the dependent chain: chunks per warp x units per chunk x the latency of one unit, level after level (4.4 ms on the
// 100k-event DAG).  A call this small is turned around: a thread owns an ACTIVITY, then an EVENT.
//
//   small_sample_kernel     one thread per activity and sample draws the duration (the same generator contract, the
//                           same functions as the pair kernel, keyed by (seed, activity index): the same bits) --
//                           the reference's sample phase, _core.cpp:323-329, all activities at once;
//   small_propagate_kernel  one CTA per sample walks the levels, one thread per event of the level: the reference's
//                           propagate phase, _core.cpp:332-350, with the same min / >= / clamp order as the sweep
//                           kernels.  It also serves injected durations and the reference-compatible stream.
//
// The sample kernel also leaves every duration in precedence-entry order ([sample][entry]): together with the static
// source rows in the same order, everything the propagate kernel needs apart from the realized times of earlier levels
// is a contiguous stream it can fetch ahead.  Both write the same event-major [row][ld] arrays as the sweep kernels.
#pragma once
#include "mcdp_chunk_sweep.cuh"
#include "mcdp_sampling.cuh"
#include "mcdp_sweep.cuh"

namespace mcdp {

// A tile: events of ONE level, at most kSmallTileEvents of them with at most kSmallTilePreds precedence entries (which
// are contiguous in evaluation order).  The propagate kernel works tile by tile.
constexpr int kSmallSampleThreads = 256;
constexpr int kSmallSweepThreads = 1024;
constexpr int kSmallTileEvents = kSmallSweepThreads;
constexpr int kSmallPrefetch = 4;  // precedence entries a thread fetches ahead per tile
constexpr int kSmallTilePreds = kSmallPrefetch * kSmallSweepThreads;
constexpr uint32_t kSmallNoPos = 0xFFFFFFFFu;
#ifndef MCDP_SMALL_BATCH
#define MCDP_SMALL_BATCH 4
#endif
constexpr int kSmallBatch = MCDP_SMALL_BATCH;  // realized times of this many sources are requested together

struct SmallParams {
    const PredRec* items;      // one record per precedence entry (its position in `next_src_row`) and per orphan activity
    int32_t n_items;           // (kSmallNoPos), sorted by sampler class; classes are padded to whole warps (the gamma
                               // samplers vote across the warp)
    const EventRec* events;    // evaluation order
    const uint32_t* pred_src;  // [P] source row of every precedence entry, evaluation order
    const uint32_t* pred_act;  // [P] its activity index (kNoAct: none)
    const int4* tiles;         // {first event, events, first precedence entry, entries}
    int32_t n_tiles;
    int64_t P;
    const DistRec* dists;
    const double* tab_pool;
    const double* log_tab;
    const int32_t* seeds;  // nullptr => seed0 + sample index
    int32_t seed0;
    int64_t n, ld;
    double* realized;         // [E][ld]
    double* durations;        // [A][ld] written by small_sample_kernel
    const double* durations_in;  // [A][ld] small_gather_kernel: the caller's injected durations
    double* dur_by_pred;      // [n][P] duration of every precedence entry: what the propagate kernel streams
    int32_t* cause;           // [E][ld]
    double max_delay;
    PhiloxKeys keys;
};

// one thread per precedence entry (and orphan activity) and sample: draw the duration
__global__ void __launch_bounds__(kSmallSampleThreads) small_sample_kernel(const __grid_constant__ SmallParams p) {
    __shared__ __align__(16) int4 s_log[kLogTabEntries];
    for (int i = threadIdx.x; i < kLogTabEntries; i += blockDim.x) s_log[i] = __ldg(reinterpret_cast<const int4*>(p.log_tab) + i);
    __syncthreads();
    const uint32_t log_tab = smem_u32(s_log);
    const int idx = int(blockIdx.x) * kSmallSampleThreads + int(threadIdx.x);
    if (idx >= p.n_items) return;  // n_items is a multiple of 32: whole warps leave
    const int64_t s = blockIdx.y;
    const uint32_t seed = p.seeds ? uint32_t(__ldg(p.seeds + s)) : uint32_t(p.seed0) + uint32_t(s);
    const int4 q0 = __ldg(reinterpret_cast<const int4*>(p.items + idx));
    const int4 q1 = __ldg(reinterpret_cast<const int4*>(p.items + idx) + 1);
    const uint32_t act = uint32_t(q0.y), meta = uint32_t(q1.x), pos = uint32_t(q1.z);
    const double base = __hiloint2double(q0.w, q0.z);
    double d = base;  // _core.cpp:304-305,325
    if ((meta >> 29) != kKindNone) {
        double ea, eb;  // both halves of the pair sampler draw for the same seed
        sample_extra2<false>(meta, uint32_t(q1.y), reinterpret_cast<const char*>(p.dists), uint32_t(q1.w),
                             reinterpret_cast<const char*>(p.tab_pool), base, act, seed, seed, false, p.keys, log_tab, ea, eb);
        d = __dadd_rn(base, ea);  // _core.cpp:328
    }
    if (act != kNoAct) p.durations[int64_t(act) * p.ld + s] = d;  // entries that share an activity write the same value
    if (pos != kSmallNoPos) p.dur_by_pred[s * p.P + pos] = d;
}

// injected durations / reference-compatible stream: the caller's [A][ld] durations in precedence-entry order
__global__ void __launch_bounds__(kSmallSampleThreads) small_gather_kernel(const __grid_constant__ SmallParams p) {
    const int64_t j = int64_t(blockIdx.x) * kSmallSampleThreads + threadIdx.x;
    if (j >= p.P) return;
    const int64_t s = blockIdx.y;
    const uint32_t act = __ldg(p.pred_act + j);
    p.dur_by_pred[s * p.P + j] = act != kNoAct ? __ldg(p.durations_in + int64_t(act) * p.ld + s) : 0.0;
}

// One CTA per sample, tile after tile, one thread per event of the tile.  What does not depend on earlier levels --
// the event records, the sources and durations of the precedence entries -- is fetched one tile ahead (registers,
// then shared memory); on the critical path of a level are the realized times of the sources (L2), the fold and a barrier.
__global__ void __launch_bounds__(kSmallSweepThreads) small_propagate_kernel(const __grid_constant__ SmallParams p) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    // two buffers of {duration f64[kSmallTilePreds], source row u32[kSmallTilePreds]}
    constexpr int kBufBytes = kSmallTilePreds * 12;
    const int64_t s = blockIdx.x;
    const int tid = int(threadIdx.x);
    const double* dur = p.dur_by_pred + s * p.P;

    struct Ahead {
        int4 tile;        // descriptor
        int4 h0, h1;      // this thread's event record of the tile
        uint32_t src[kSmallPrefetch];
        double d[kSmallPrefetch];
    };
    auto fetch = [&](int t, Ahead& a) {
        a.tile = make_int4(0, 0, 0, 0);
        if (t >= p.n_tiles) return;
        a.tile = __ldg(p.tiles + t);
        if (tid < a.tile.y) {
            a.h0 = __ldg(reinterpret_cast<const int4*>(p.events + a.tile.x + tid));
            a.h1 = __ldg(reinterpret_cast<const int4*>(p.events + a.tile.x + tid) + 1);
        }
#pragma unroll
        for (int r = 0; r < kSmallPrefetch; ++r) {
            const int j = tid + r * kSmallSweepThreads;
            if (j < a.tile.w) {
                a.src[r] = __ldg(p.pred_src + uint32_t(a.tile.z) + j);
                a.d[r] = __ldcs(dur + uint32_t(a.tile.z) + j);
            }
        }
    };
    auto park = [&](const Ahead& a, int buf) {  // the fetched entries into the tile's shared-memory buffer
        double* d_s = reinterpret_cast<double*>(small_smem + buf * kBufBytes);
        uint32_t* src_s = reinterpret_cast<uint32_t*>(small_smem + buf * kBufBytes + kSmallTilePreds * 8);
#pragma unroll
        for (int r = 0; r < kSmallPrefetch; ++r) {
            const int j = tid + r * kSmallSweepThreads;
            if (j < a.tile.w) {
                d_s[j] = a.d[r];
                src_s[j] = a.src[r];
            }
        }
    };

    Ahead cur, nxt;
    fetch(0, cur);
    park(cur, 0);
    __syncthreads();
    for (int t = 0; t < p.n_tiles; ++t) {
        fetch(t + 1, nxt);  // in flight while this tile is folded
        const int buf = t & 1;
        if (tid < cur.tile.y) {
            const double* d_s = reinterpret_cast<const double*>(small_smem + buf * kBufBytes);
            const uint32_t* src_s = reinterpret_cast<const uint32_t*>(small_smem + buf * kBufBytes + kSmallTilePreds * 8);
            // _core.cpp:333-337
            const double earliest = __hiloint2double(cur.h1.y, cur.h1.x);
            const double ub = __dadd_rn(earliest, p.max_delay);
            double lat = earliest;
            int cause = -1;
            const int rel = cur.h0.z - cur.tile.z, fan = cur.h0.w;
            for (int k0 = 0; k0 < fan; k0 += kSmallBatch) {
                uint32_t src[kSmallBatch];
                double rs[kSmallBatch];
#pragma unroll
                for (int k = 0; k < kSmallBatch; ++k) {  // the realized times of several sources at once
                    src[k] = k0 + k < fan ? src_s[rel + k0 + k] : 0u;
                    if (k0 + k < fan) rs[k] = __ldcg(p.realized + int64_t(src[k]) * p.ld + s);
                }
#pragma unroll
                for (int k = 0; k < kSmallBatch; ++k) {
                    if (k0 + k < fan) {
                        // _core.cpp:341-346
                        const double tt = ref_min(__dadd_rn(rs[k], d_s[rel + k0 + k]), ub);
                        if (tt >= lat) {
                            lat = tt;
                            cause = int(src[k]);
                        }
                    }
                }
            }
            // _core.cpp:348-349
            __stcg(p.realized + int64_t(uint32_t(cur.h0.x)) * p.ld + s, ref_min(lat, ub));
            __stcg(p.cause + int64_t(uint32_t(cur.h0.x)) * p.ld + s, cause);
        }
        park(nxt, buf ^ 1);
        cur = nxt;
        __syncthreads();  // the tile's realized times are in L2, the next tile's entries in shared memory
    }
}

}  // namespace mcdp
