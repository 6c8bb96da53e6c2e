// mcdp_sweep.cuh -- the fused sample + max-plus sweep kernel for sm_100a.
//
// Replaces Simulator::run (reference _core.cpp:312-353) for a block of seeds:
//   * one warp owns 64 adjacent samples (two per lane) of the event-major / sample-minor arrays,
//     so every predecessor read and every realized / cause / duration write is a 16-byte
//     vector access and a 512-byte contiguous row segment per warp;
//   * `warps_per_group` warps share the same 64 samples and split every topological level
//     between them (round-robin over the level's events), meeting at a named barrier per level:
//     this is how large DAGs, whose per-sample output footprint limits the resident sample
//     count, still fill the SMs;
//   * the delay of each precedence entry is drawn (Philox, mcdp_sampling.cuh) at the point of
//     use and streamed out, never re-read.
#pragma once
#include "mcdp_sampling.cuh"

namespace mcdp {

// CTA shape: at most 16 warps (groups_per_cta x warps_per_group), two such CTAs per SM => 64
// registers per thread and 32 resident warps per SM.
#ifndef MCDP_MAX_THREADS
#define MCDP_MAX_THREADS 512
#endif
#ifndef MCDP_MIN_BLOCKS
#define MCDP_MIN_BLOCKS 2
#endif

// kModeAttr: reduced statistics + delay-cause attribution (per-activity counts of being the binding predecessor)
enum SweepMode { kModeFull = 0, kModeInjected = 1, kModeReduced = 2, kModeAttr = 3 };

struct SweepParams {
    const EventRec* events;
    const PredRec* preds;
    const int32_t* level_begin;
    const ChunkUnit* chunks;             // chunk stream (full / injected modes, mcdp_chunk_sweep.cuh)
    const int32_t* chunk_level_begin;    // [n_levels + 1] positions into chunks
    const PredRec* orphans;
    const DistRec* dists;
    const double* tab_pool;
    const double* log_tab;  // {rc, lc} pairs of mcdp_math.cuh (kLogTabEntries)
    const int32_t* seeds;  // nullptr => seed0 + sample index
    double* realized;      // [rows][ld]  (output in full/injected mode, scratch in reduced mode)
    double* durations;     // [A][ld]     full mode: written
    const double* inj;     // [A][ld]     injected mode: read
    int32_t* cause;        // [E][ld]
    double* sum;           // reduced mode accumulators
    double* sumsq;
    unsigned long long* late;
    uint32_t* hist;
    unsigned long long* cause_act;   // kModeAttr: [A] samples in which an entry with this activity decided its target
    unsigned long long* cause_none;  // kModeAttr: [E] samples with cause_event == -1
    double thresholds[MCDP_MAX_THRESHOLDS];
    double hist_lo, hist_scale;
    double max_delay;
    int64_t n, ld;
    uint32_t ldb8, ldb4;                // ld * 8, ld * 4 (row strides in bytes, < 2^32)
    uint32_t smem_tab_off, smem_ring_off;  // dynamic shared memory: byte offsets of the table pool and of the chunk rings
    int32_t n_levels, n_orphans, n_dists, tab_pool_len;
    int32_t n_thresholds, n_bins, E, n_chunks;
    uint32_t last_pred;  // index of the last precedence record (prefetch clamp)
    int32_t seed0;
    PhiloxKeys keys;  // round keys of Philox key word 0 (stream_key + r * 0x9E3779B9)
    int32_t warps_per_group;
    int32_t batches_per_group;  // reduced mode: 64-sample batches folded per group before a flush
};

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Warp totals of s and q with ONE butterfly: after the first exchange the lower half-warp carries s, the upper
// half-warp q.  Lanes 0-15 return the total of s, lanes 16-31 the total of q.
__device__ __forceinline__ double warp_sum2(double s, double q, int lane) {
    const bool upper = lane & 16;
    const double give = upper ? s : q, keep = upper ? q : s;
    double v = keep + __shfl_xor_sync(0xFFFFFFFFu, give, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Per-event statistics that are sums over the warp's samples (reduced modes).  `s` / `q`: this lane's sum of delays
// and of squared delays; `late_packed`: this lane's exceedance counts, 8 bits per threshold (a warp holds at most 128
// samples, so the fields of the warp total cannot carry into each other).  One butterfly, one REDUX, and the atomics
// of the thresholds issued by different lanes of one instruction.
template <typename P>
__device__ __forceinline__ void flush_sums(const P& p, uint32_t ev, int lane, double s, double q, uint32_t late_packed) {
    if (p.sum || p.sumsq) {
        const double t = warp_sum2(s, q, lane);
        if (lane == 0 && p.sum) atomicAdd(p.sum + ev, t);
        if (lane == 16 && p.sumsq) atomicAdd(p.sumsq + ev, t);
    }
    if (p.late) {
        const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, late_packed);
        if (lane < p.n_thresholds) {
            const uint32_t cnt = (tot >> (8 * lane)) & 0xFFu;
            if (cnt) atomicAdd(p.late + size_t(lane) * p.E + ev, (unsigned long long)cnt);
        }
    }
}

// std::min(a, b) of the reference build: (b < a) ? b : a  -- NOT fmin (NaN / signed-zero differ).
__device__ __forceinline__ double ref_min(double a, double b) { return (b < a) ? b : a; }

// MULTI (reduced mode only): a group folds several 64-sample batches per event before flushing.
template <int MODE, bool SMEM, bool MULTI = false>
__global__ void __launch_bounds__(MCDP_MAX_THREADS, MCDP_MIN_BLOCKS) sweep_kernel(const __grid_constant__ SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the log table of mcdp_math.cuh opens the dynamic shared memory
    uint32_t smem_base = uint32_t(__cvta_generic_to_shared(smem_raw));
    asm volatile("" : "+r"(smem_base));
    for (int i = threadIdx.x; i < kLogTabEntries; i += blockDim.x)
        reinterpret_cast<int4*>(smem_raw)[i] = __ldg(reinterpret_cast<const int4*>(p.log_tab) + i);
    const uint32_t log_tab = smem_base;
    size_t smem_used = kLogTabBytes;
    typename Mem<SMEM>::ptr dists, tab;
    if constexpr (SMEM) {
        // stage distribution records + guide / inverse-CDF tables once per CTA
        int4* s_dists = reinterpret_cast<int4*>(smem_raw + kLogTabBytes);
        double* s_tab = reinterpret_cast<double*>(smem_raw + p.smem_tab_off);
        const int n16 = int(sizeof(DistRec) / 16) * p.n_dists;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) s_dists[i] = __ldg(reinterpret_cast<const int4*>(p.dists) + i);
        for (int i = threadIdx.x; i < p.tab_pool_len; i += blockDim.x) s_tab[i] = __ldg(p.tab_pool + i);
        dists = smem_base + uint32_t(kLogTabBytes);
        tab = smem_base + p.smem_tab_off;
        smem_used = kLogTabBytes + sizeof(DistRec) * p.n_dists + sizeof(double) * p.tab_pool_len;
    } else {
        dists = reinterpret_cast<const char*>(p.dists);
        tab = reinterpret_cast<const char*>(p.tab_pool);
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // reduced mode: one private histogram per warp, flushed once per event
    uint32_t* w_hist = nullptr;
    if constexpr (MODE == kModeReduced) {
        if (MULTI && p.hist) {
            uint32_t* all = reinterpret_cast<uint32_t*>(smem_raw + ((smem_used + 15) & ~size_t(15)));
            for (int i = threadIdx.x; i < int(blockDim.x >> 5) * p.n_bins; i += blockDim.x) all[i] = 0u;
            w_hist = all + warp * p.n_bins;
        }
    }
    __syncthreads();

    const int wpg = p.warps_per_group;
    const int group_in_cta = warp / wpg;
    const int wsub = warp - group_in_cta * wpg;
    const int groups_per_cta = (blockDim.x >> 5) / wpg;
    // A group owns `batches` consecutive 64-sample batches (1 except in reduced mode, where a warp
    // folds the statistics of all its batches before touching the global accumulators).
    const int batches = (MODE == kModeReduced && MULTI) ? p.batches_per_group : 1;
    const int64_t batch0 = (int64_t(blockIdx.x) * groups_per_cta + group_in_cta) * batches;
    if (batch0 * 64 >= p.n) return;  // whole group (all its warps) out of range
    const PhiloxKeys& key0 = p.keys;
    const uint32_t ldb8 = uint32_t(p.ld) * 8u, ldb4 = uint32_t(p.ld) * 4u;

    // One event for the two samples (columns s0, s0 + 1) a lane owns in one batch.
    // ld is a multiple of 64, so every lane of a launched batch owns two in-bounds columns; columns
    // >= n are padding (computed and written like the others, never read back by the host side).
    // r_lane / d_lane / i_lane / c_lane: per-lane column bases of realized / durations / injected
    // durations / cause; a row is reached with one 32x32->64 multiply-add (ld * 8 < 2^32).
    auto event_body = [&](const int4& e0, const int4& e1, char* r_lane, char* d_lane, const char* i_lane, char* c_lane,
                          uint32_t seed_a, uint32_t seed_b, bool paired, double& ra, double& rb) {
        const uint32_t row = uint32_t(e0.x), fan_in = uint32_t(e0.w);
        const PredRec* pr = p.preds + uint32_t(e0.z);
        const double earliest = __hiloint2double(e1.y, e1.x);
        const double ub = __dadd_rn(earliest, p.max_delay);  // _core.cpp:334
        // _core.cpp:336-337
        double lat_a = earliest, lat_b = earliest;
        int cause_a = -1, cause_b = -1;
        // The predecessor row of the NEXT entry is requested before the current entry's delay is
        // drawn (its source row is carried by the event record / the current entry record), so the
        // HBM latency of the gather hides behind the sampling arithmetic.
        double2 nrs = make_double2(0.0, 0.0);
        if (fan_in) nrs = __ldcg(reinterpret_cast<const double2*>(r_lane + size_t(uint32_t(e1.z)) * ldb8));
        for (uint32_t k = 0; k < fan_in; ++k, ++pr) {
            const int4 q0 = __ldg(reinterpret_cast<const int4*>(pr));
            const int4 q1 = __ldg(reinterpret_cast<const int4*>(pr) + 1);
            const double2 rs = nrs;
            if (k + 1 < fan_in) nrs = __ldcg(reinterpret_cast<const double2*>(r_lane + size_t(uint32_t(q1.z)) * ldb8));
            const uint32_t act = uint32_t(q0.y);
            const double base = __hiloint2double(q0.w, q0.z);
            const uint32_t meta = uint32_t(q1.x);
            const int src_event = q0.x;  // full / injected mode: rows are event ids
            double da, db;
            if constexpr (MODE == kModeInjected) {
                double2 dd = make_double2(0.0, 0.0);
                if (act != kNoAct) dd = __ldcs(reinterpret_cast<const double2*>(i_lane + size_t(act) * ldb8));
                da = dd.x;
                db = dd.y;
            } else {
                if ((meta >> 29) == kKindNone) {
                    da = db = base;  // _core.cpp:304-305,325
                } else {
                    double ea, eb;
                    sample_extra2<SMEM>(meta, uint32_t(q1.y), dists, uint32_t(q1.w), tab, base, act, seed_a, seed_b, paired,
                                        key0, log_tab, ea, eb);
                    da = __dadd_rn(base, ea);  // _core.cpp:328
                    db = __dadd_rn(base, eb);
                }
                if constexpr (MODE == kModeFull) {
                    if (act != kNoAct) __stcs(reinterpret_cast<double2*>(d_lane + size_t(act) * ldb8), make_double2(da, db));
                }
            }
            // _core.cpp:341-346
            const double ta = ref_min(__dadd_rn(rs.x, da), ub);
            const double tb = ref_min(__dadd_rn(rs.y, db), ub);
            if (ta >= lat_a) {
                lat_a = ta;
                cause_a = src_event;
            }
            if (tb >= lat_b) {
                lat_b = tb;
                cause_b = src_event;
            }
        }
        // _core.cpp:348-349
        ra = ref_min(lat_a, ub);
        rb = ref_min(lat_b, ub);
        // realized rows are gathered once per consumer, by other warps: keep them out of L1 (L2 only) so
        // the warp-uniform record lines stay resident there
        __stcg(reinterpret_cast<double2*>(r_lane + size_t(row) * ldb8), make_double2(ra, rb));
        if constexpr (MODE != kModeReduced)
            __stcs(reinterpret_cast<int2*>(c_lane + size_t(row) * ldb4), make_int2(cause_a, cause_b));
    };
    auto seeds_of = [&](int64_t s0, uint32_t& seed_a, uint32_t& seed_b, bool& paired) {
        seed_a = 0u;
        seed_b = 1u;
        if constexpr (MODE != kModeInjected) {
            if (p.seeds) {
                seed_a = s0 < p.n ? uint32_t(__ldg(p.seeds + s0)) : 0u;
                seed_b = s0 + 1 < p.n ? uint32_t(__ldg(p.seeds + s0 + 1)) : seed_a + 1u;
            } else {
                seed_a = uint32_t(p.seed0) + uint32_t(s0);
                seed_b = seed_a + 1u;
            }
        }
        paired = ((seed_a & 1u) == 0u) && (seed_b == seed_a + 1u);
    };

    // full / injected mode: the lane's two columns and seeds are fixed for the whole sweep
    const int64_t s0_fixed = batch0 * 64 + 2 * lane;
    uint32_t seed_a0, seed_b0;
    bool paired0;
    seeds_of(s0_fixed, seed_a0, seed_b0, paired0);
    char* const r_lane0 = reinterpret_cast<char*>(p.realized) + s0_fixed * 8;
    char* const d_lane0 = reinterpret_cast<char*>(p.durations) + s0_fixed * 8;
    const char* const i_lane0 = reinterpret_cast<const char*>(p.inj) + s0_fixed * 8;
    char* const c_lane0 = reinterpret_cast<char*>(p.cause) + s0_fixed * 4;

    // Level scheduling.  One warp per group: events in stream order, no synchronisation.  Several
    // warps per group: the warps of a group pull event positions from a shared-memory counter (one
    // per level parity; the idle one is re-armed for the next level while the current level runs),
    // so uneven fan-in or memory latency does not leave warps waiting at the level barrier.
    __shared__ int s_cursor[16][2];
    const bool dyn = wpg > 1;
    if (dyn) {
        if (wsub == 0 && lane == 0) s_cursor[group_in_cta][0] = 0;
        group_barrier(1 + group_in_cta, wpg * 32);
    }
    auto grab = [&](int parity) -> int {
        int v = 0;
        if (lane == 0) v = atomicAdd(&s_cursor[group_in_cta][parity], 1);
        return __shfl_sync(0xFFFFFFFFu, v, 0);
    };

    for (int lvl = 0; lvl < p.n_levels; ++lvl) {
        const int lb = __ldg(p.level_begin + lvl), le = __ldg(p.level_begin + lvl + 1);
        const int par = lvl & 1;
        int i_next;
        if (dyn) {
            if (wsub == 0 && lane == 0) s_cursor[group_in_cta][par ^ 1] = le;  // next level starts at le
            i_next = grab(par);
        } else {
            i_next = lb;
        }
        while (i_next < le) {
            const int i = i_next;
            i_next = dyn ? grab(par) : i + 1;
            const int4 e0 = __ldg(reinterpret_cast<const int4*>(p.events + i));
            const int4 e1 = __ldg(reinterpret_cast<const int4*>(p.events + i) + 1);
            if (dyn) {
                // this warp's next event is already known: pull its record, and (estimating four entries
                // per event in between) its first entry records, into this SM's L1 while event i runs
                const int in = min(i_next, p.E - 1);
                prefetch_l1(p.events + in);
                prefetch_l1(p.preds + min(uint32_t(e0.z) + uint32_t(e0.w) + 4u * uint32_t(in - i - 1), p.last_pred));
            }
            if constexpr (MODE != kModeReduced) {
                double ra, rb;
                event_body(e0, e1, r_lane0, d_lane0, i_lane0, c_lane0, seed_a0, seed_b0, paired0, ra, rb);
            } else {
                // fold the delay statistics of all batches of this group, then one flush per event
                const double earliest = __hiloint2double(e1.y, e1.x);
                const uint32_t ev = uint32_t(e0.y);
                double acc = 0.0, acc2 = 0.0;
                int late[MCDP_MAX_THRESHOLDS] = {0, 0, 0, 0};
                for (int b = 0; b < batches; ++b) {
                    const int64_t s0 = (batch0 + b) * 64 + 2 * lane;
                    if ((batch0 + b) * 64 >= p.n) break;  // warp-uniform
                    uint32_t seed_a, seed_b;
                    bool paired;
                    seeds_of(s0, seed_a, seed_b, paired);
                    double ra, rb;
                    event_body(e0, e1, reinterpret_cast<char*>(p.realized) + s0 * 8, nullptr, nullptr, nullptr, seed_a, seed_b,
                               paired, ra, rb);
                    const bool valid_a = s0 < p.n, valid_b = s0 + 1 < p.n;
                    const double xa = valid_a ? ra - earliest : 0.0;
                    const double xb = valid_b ? rb - earliest : 0.0;
                    acc += xa + xb;
                    acc2 += xa * xa + xb * xb;
#pragma unroll
                    for (int t = 0; t < MCDP_MAX_THRESHOLDS; ++t)
                        if (t < p.n_thresholds) late[t] += int(valid_a && xa > p.thresholds[t]) + int(valid_b && xb > p.thresholds[t]);
                    if (p.hist) {
                        const int nb = p.n_bins;
                        int ba = min(max(int(floor((xa - p.hist_lo) * p.hist_scale)), 0), nb - 1);
                        int bb = min(max(int(floor((xb - p.hist_lo) * p.hist_scale)), 0), nb - 1);
                        if (w_hist) {  // several batches per group: private shared-memory histogram
                            if (valid_a) atomicAdd(w_hist + ba, 1u);
                            if (valid_b) atomicAdd(w_hist + bb, 1u);
                        } else {  // single batch: aggregate equal bins across the warp, one atomic per distinct bin
                            if (!valid_a) ba = -1 - lane;  // unique keys: match groups of size 1, skipped below
                            if (!valid_b) bb = -1 - lane;
                            uint32_t* h = p.hist + size_t(ev) * nb;
                            const unsigned ga = __match_any_sync(0xFFFFFFFFu, ba);
                            if (ba >= 0 && lane == __ffs(ga) - 1) atomicAdd(h + ba, uint32_t(__popc(ga)));
                            const unsigned gb = __match_any_sync(0xFFFFFFFFu, bb);
                            if (bb >= 0 && lane == __ffs(gb) - 1) atomicAdd(h + bb, uint32_t(__popc(gb)));
                        }
                    }
                }
                if (p.sum) {
                    const double s = warp_sum(acc);
                    if (lane == 0) atomicAdd(p.sum + ev, s);
                }
                if (p.sumsq) {
                    const double s = warp_sum(acc2);
                    if (lane == 0) atomicAdd(p.sumsq + ev, s);
                }
                if (p.late) {
#pragma unroll
                    for (int t = 0; t < MCDP_MAX_THRESHOLDS; ++t) {
                        if (t < p.n_thresholds) {
                            const int c = __reduce_add_sync(0xFFFFFFFFu, late[t]);
                            if (lane == 0 && c) atomicAdd(p.late + size_t(t) * p.E + ev, (unsigned long long)c);
                        }
                    }
                }
                if (w_hist) {
                    __syncwarp();
                    uint32_t* h = p.hist + size_t(ev) * p.n_bins;
                    for (int b = lane; b < p.n_bins; b += 32) {
                        const uint32_t v = w_hist[b];
                        if (v) {
                            atomicAdd(h + b, v);
                            w_hist[b] = 0u;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (dyn) group_barrier(1 + group_in_cta, wpg * 32);
    }

    if constexpr (MODE == kModeFull) {
        // activities no precedence entry references still get their sampled duration (_core.cpp:323-329)
        char* const d_lane = d_lane0;
        for (int i = wsub; i < p.n_orphans; i += wpg) {
            const int4 q0 = __ldg(reinterpret_cast<const int4*>(p.orphans + i));
            const int4 q1 = __ldg(reinterpret_cast<const int4*>(p.orphans + i) + 1);
            const uint32_t act = uint32_t(q0.y), meta = uint32_t(q1.x);
            const double base = __hiloint2double(q0.w, q0.z);
            double da = base, db = base;
            if ((meta >> 29) != kKindNone) {
                double ea, eb;
                sample_extra2<SMEM>(meta, uint32_t(q1.y), dists, uint32_t(q1.w), tab, base, act, seed_a0, seed_b0, paired0,
                                    key0, log_tab, ea, eb);
                da = __dadd_rn(base, ea);
                db = __dadd_rn(base, eb);
            }
            __stcs(reinterpret_cast<double2*>(d_lane + size_t(act) * ldb8), make_double2(da, db));
        }
    }
}

// out[c][r] = in[r][c]  (in: rows x cols with row stride in_ld; out: cols x rows, stride out_ld).
// 1-D grid of 32x32 tiles (either extent can exceed the 65535 limit of grid.y).
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, int64_t in_ld, int64_t rows,
                                                        int64_t cols, T* __restrict__ out, int64_t out_ld,
                                                        int64_t tiles_c) {
    __shared__ T tile[32][33];
    const int64_t tile_r = int64_t(blockIdx.x) / tiles_c, tile_c = int64_t(blockIdx.x) - tile_r * tiles_c;
    const int64_t c0 = tile_c * 32, r0 = tile_r * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t r = r0 + ty + j, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + j][tx] = in[r * in_ld + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t c = c0 + ty + j, r = r0 + tx;
        if (r < rows && c < cols) out[c * out_ld + r] = tile[tx][ty + j];
    }
}

}  // namespace mcdp
