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
constexpr int kSmallTilePreds = 4 * kSmallSweepThreads;
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
    int32_t E;
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
    double* realized_by_sample;  // [n][E] the sample's realized times side by side (the [E][ld] output has a row per event)
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
// the event records (registers), the sources and durations of the precedence entries (cp.async straight into a
// ring of shared-memory tiles, no registers) -- is fetched two tiles ahead; on the critical path of a level are
// the realized times of the sources (L2, kSmallBatch requests at a time), the fold and a barrier.
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {  // at most N of the thread's groups still pending
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kSmallStages = 3;  // shared-memory tiles: the one being folded and two in flight

__global__ void __launch_bounds__(kSmallSweepThreads) small_propagate_kernel(const __grid_constant__ SmallParams p) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    // kSmallStages buffers of {duration f64[kSmallTilePreds], source row u32[kSmallTilePreds]}
    constexpr int kBufBytes = kSmallTilePreds * 12;
    const uint32_t smem0 = smem_u32(small_smem);
    const int64_t s = blockIdx.x;
    const int tid = int(threadIdx.x);
    const double* dur = p.dur_by_pred + s * p.P;
    double* mine = p.realized_by_sample + s * int64_t(p.E);  // what later levels read: neighbours share sectors

    struct Ahead {
        int n_ev, pred0;      // of the tile
        int4 h0;              // this thread's event record: row, event, pred_begin, fan_in
        int e_lo, e_hi;       // earliest
    };
    auto load_desc = [&](int t) { return t < p.n_tiles ? __ldg(p.tiles + t) : make_int4(0, 0, 0, 0); };
    // tile t (descriptor already in registers): event record into registers, entries by cp.async into buffer t % stages
    auto fetch = [&](int t, const int4& tile, Ahead& a) {
        a.n_ev = tile.y;
        a.pred0 = tile.z;
        if (tid < tile.y) {
            a.h0 = __ldg(reinterpret_cast<const int4*>(p.events + tile.x + tid));
            const int2 e = __ldg(reinterpret_cast<const int2*>(p.events + tile.x + tid) + 2);
            a.e_lo = e.x;
            a.e_hi = e.y;
        }
        const uint32_t d_s = smem0 + uint32_t((t % kSmallStages) * kBufBytes), src_s = d_s + uint32_t(kSmallTilePreds * 8);
        for (int j = tid; j < tile.w; j += kSmallSweepThreads) {
            cp_async_8(d_s + uint32_t(j) * 8u, dur + uint32_t(tile.z) + j);
            cp_async_4(src_s + uint32_t(j) * 4u, p.pred_src + uint32_t(tile.z) + j);
        }
        cp_async_commit();
    };

    // Software pipeline: while tile t is folded, the entries of tiles t + 1 and t + 2 are in flight and the descriptor
    // of tile t + 3 is on its way -- no load of the static data is waited for in the iteration that issued it.
    Ahead cur, nx1, nx2;
    fetch(0, load_desc(0), cur);
    fetch(1, load_desc(1), nx1);
    int4 desc2 = load_desc(2);
    cp_async_wait_group<1>();
    __syncthreads();
    for (int t = 0; t < p.n_tiles; ++t) {
        const int4 desc3 = load_desc(t + 3);
        fetch(t + 2, desc2, nx2);
        desc2 = desc3;
        const int buf = t % kSmallStages;
        if (tid < cur.n_ev) {
            const double* d_s = reinterpret_cast<const double*>(small_smem + buf * kBufBytes);
            const uint32_t* src_s = reinterpret_cast<const uint32_t*>(small_smem + buf * kBufBytes + kSmallTilePreds * 8);
            // _core.cpp:333-337
            const double earliest = __hiloint2double(cur.e_hi, cur.e_lo);
            const double ub = __dadd_rn(earliest, p.max_delay);
            double lat = earliest;
            int cause = -1;
            const int rel = cur.h0.z - cur.pred0, fan = cur.h0.w;
            for (int k0 = 0; k0 < fan; k0 += kSmallBatch) {
                uint32_t src[kSmallBatch];
                double rs[kSmallBatch];
#pragma unroll
                for (int k = 0; k < kSmallBatch; ++k) {  // the realized times of several sources at once
                    src[k] = k0 + k < fan ? src_s[rel + k0 + k] : 0u;
                    if (k0 + k < fan) rs[k] = __ldcg(mine + src[k]);
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
            const double r = ref_min(lat, ub);
            __stcg(mine + uint32_t(cur.h0.x), r);
            __stcs(p.realized + int64_t(uint32_t(cur.h0.x)) * p.ld + s, r);
            __stcs(p.cause + int64_t(uint32_t(cur.h0.x)) * p.ld + s, cause);
        }
        cur = nx1;
        nx1 = nx2;
        cp_async_wait_group<1>();  // tile t + 1 has landed (tile t + 2 may still be in flight)
        __syncthreads();           // the tile's realized times are in L2, the next tile's entries in shared memory
    }
}

}  // namespace mcdp
