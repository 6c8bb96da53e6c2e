// mcdp_quad_sweep.cuh -- the chunk-stream sweep of mcdp_chunk_sweep.cuh with FOUR samples per lane, sm_100a.
//
// Replaces Simulator::run (reference _core.cpp:312-353) for a block of seeds, like the pair kernel.
// Same chunk stream, same ring, same level cursor; what differs:
//   * a warp owns 128 adjacent samples, four per lane, so every predecessor read and every
//     realized / duration write is ONE 256-bit access per lane (ld/st.global.v4.f64, sm_100) and a
//     1 KB contiguous row segment per warp; cause_event rows are 128-bit stores;
//   * the warp-uniform work of a stream unit -- record words, kind dispatch, distribution-record
//     loads, row address arithmetic, loop control -- is paid once per 128 edge-samples instead of
//     once per 64, and four independent dependency chains per lane are in flight;
//   * up to 20 warps per SM at 96 registers (one CTA of 20 warps, or two of 10): the register file no
//     longer shapes the loop body (the pair kernel sits at the 64-register spill edge).
// Results are bit-identical to the pair kernel (same generator contract, same recurrence).
#pragma once
#include "mcdp_chunk_sweep.cuh"
#include "mcdp_sampling4.cuh"

namespace mcdp {

struct D4 {
    double x, y, z, w;
};
// 256-bit global accesses (PTX ISA 8.8, sm_100+): 32-byte aligned addresses
__device__ __forceinline__ D4 ldcg_d4(const void* ptr) {  // L2 only: rows written by other warps
    D4 v;
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(ptr));
    return v;
}
__device__ __forceinline__ D4 ldcs_d4(const void* ptr) {  // streamed once
    D4 v;
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(ptr));
    return v;
}
__device__ __forceinline__ void stcg_d4(void* ptr, const D4& v) {
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void stcs_d4(void* ptr, const D4& v) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

constexpr int kQuadSamples = 128;  // samples per group

// Reduced modes: per-warp staging of the per-event statistics in shared memory (behind the chunk rings).  A warp parks
// the per-lane partial sums of up to kStatSlots events (and their histogram counts, as 16-bit pairs) and folds them
// in one transposed pass -- lane -> (event, lane block) -- instead of running a five-stage shuffle butterfly and four
// match.any rounds per event.  Layout per warp: [kStatSlots][32] {sum, sum of squares} f64 pairs, then
// [kStatSlots][32] u32 histogram words (two bins per word, so up to kStatMaxBins bins).
#ifndef MCDP_RELAX_LITERAL
#define MCDP_RELAX_LITERAL 1  // 1: the reference's literal clamp-then-compare form of the recurrence; 0: relax() of
                              // mcdp_sweep.cuh (two compares, no clamp).  Measured on one B200, C3 / C3-MT / C2 full, ms:
                              // literal 25.57 / 29.3 / 33.5, relax() 26.20 / 30.1 / 34.4 -- fewer instructions, worse schedule.
#endif
#ifndef MCDP_STAGE_SUMS
#define MCDP_STAGE_SUMS 1  // 0: shuffle butterfly per event (flush_sums)
#endif
#ifndef MCDP_STAGE_HIST
#define MCDP_STAGE_HIST 0  // 0: match.any aggregation per event; N >= 1: shared-memory counts, N replicas per warp (lane % N).
                           // Measured on B200 (C3 / C5 / C4 reduced, ms): match.any 32.1 / 181 / 621, shared counts 33.2 / 203 /
                           // 663 (four replicas 32.8 / 208 / 673): concentrated bins serialise the shared-memory reductions.
#endif
constexpr int kStatSlots = 4;
constexpr int kStatMaxBins = MCDP_STAGE_HIST ? 64 : 0;
constexpr int kStatHistReplicas = MCDP_STAGE_HIST ? MCDP_STAGE_HIST : 1;
constexpr int kStatSumBytes = kStatSlots * 32 * 16;
constexpr int kStatHistBytes = kStatSlots * 32 * 4 * kStatHistReplicas;
constexpr int kStatStride = kStatSumBytes + kStatHistBytes;
__host__ __device__ inline size_t quad_stat_bytes(int n_warps) { return size_t(n_warps) * size_t(kStatStride); }

// c += K when x > th: one compare and one predicated add (the C form costs a select and an add per term)
__device__ __forceinline__ void add_if_gt(uint32_t& c, double x, double th, uint32_t k) {
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(c) : "d"(x), "d"(th), "r"(k));
}
// Thread-block cluster (small launches): the CTAs of a cluster share ONE sample group and split every level between
// them, so a launch with fewer groups than SMs still spreads over the machine.  Rows written by one CTA are read by
// the others through L2 (st.cg / ld.cg); the per-level cluster barrier orders them (release / acquire).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void atoms_add_u32(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// CTA size limit of the quad kernel: 640 threads = 20 resident warps per SM at 96 registers.  Measured on B200
// (C3, profiles/r01_quad_ab.txt): 16 warps at 128 registers 77 % of the measured HBM peak, 20 warps at 96
// registers 80.5 %, 24 warps at 80 registers 62 % (spills).
#ifndef MCDP_QUAD_MAX_THREADS
#define MCDP_QUAD_MAX_THREADS 640
#endif

// Unroll factor of the unit loop.  1 ships.  2 lets the two register sets of the one-unit-ahead gather alternate
// instead of being copied (8 moves per unit) at twice the loop's code size -- measured in round 2 on one B200 (ms,
// unroll 1 / 2): C3 full 25.5 / 31.5, C3-MT 29.3 / 37.9, C3 reduced 28.2 / 41.1, C4 reduced 540 / 827: the doubled
// body falls out of the instruction caches.
#ifndef MCDP_QUAD_UNIT_UNROLL
#define MCDP_QUAD_UNIT_UNROLL 1
#endif
constexpr int kQuadUnitUnroll = MCDP_QUAD_UNIT_UNROLL;

// what MCDP_OPT_SAMPLES_PER_LANE = 0 (auto) selects
#ifndef MCDP_AUTO_SPL
#define MCDP_AUTO_SPL 4
#endif

template <int MODE, bool SMEM, bool DYN>
__global__ void __launch_bounds__(MCDP_QUAD_MAX_THREADS, 1) quad_sweep_kernel(const __grid_constant__ SweepParams p) {
    constexpr bool kReduced = MODE == kModeReduced || MODE == kModeAttr;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    // dynamic shared memory as in chunk_sweep_kernel: [log table][DistRec[] + table pool (SMEM)][per-warp chunk rings]
    uint32_t smem_base = smem_u32(smem_dyn);
    asm volatile("" : "+r"(smem_base));
    for (int i = threadIdx.x; i < kLogTabEntries; i += blockDim.x)
        reinterpret_cast<int4*>(smem_dyn)[i] = __ldg(reinterpret_cast<const int4*>(p.log_tab) + i);
    const uint32_t log_tab = smem_base;
    typename Mem<SMEM>::ptr dists, tab;
    if constexpr (SMEM) {
        int4* s_dists = reinterpret_cast<int4*>(smem_dyn + kLogTabBytes);
        double* s_tab = reinterpret_cast<double*>(smem_dyn + p.smem_tab_off);
        const int n16 = int(sizeof(DistRec) / 16) * p.n_dists;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) s_dists[i] = __ldg(reinterpret_cast<const int4*>(p.dists) + i);
        for (int i = threadIdx.x; i < p.tab_pool_len; i += blockDim.x) s_tab[i] = __ldg(p.tab_pool + i);
        dists = smem_base + uint32_t(kLogTabBytes);
        tab = smem_base + p.smem_tab_off;
    } else {
        dists = reinterpret_cast<const char*>(p.dists);
        tab = reinterpret_cast<const char*>(p.tab_pool);
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    const uint32_t ring0 = smem_base + p.smem_ring_off + uint32_t(warp) * uint32_t(kRingStride);
    const uint32_t bar0 = ring0 + 2u * uint32_t(kChunkBytes);
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_fence_init();
    }
    // one END unit (closes the open event of a warp whose level has run out of chunks)
    __shared__ __align__(16) int4 s_end_unit[2];
    if (threadIdx.x == 0) {
        s_end_unit[0] = make_int4(0, 0, 0, 0);
        s_end_unit[1] = make_int4(int(kKindEnd << 29), 0, int(kNoRow), 0);
    }
    const uint32_t end_buf = smem_u32(s_end_unit);
    // reduced modes: this warp's statistics staging area; the histogram words start at zero
    const uint32_t stat0 = smem_base + p.smem_stat_off + uint32_t(warp) * uint32_t(kStatStride);
    const bool stage_hist = MCDP_STAGE_HIST != 0 && kReduced && p.hist != nullptr && p.n_bins <= kStatMaxBins;
    if constexpr (kReduced) {
#pragma unroll
        for (int k = 0; k < kStatSlots; ++k)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(stat0 + uint32_t(kStatSumBytes) + uint32_t(k * 32 + lane) * 4u), "r"(0u) : "memory");
    }
    __syncthreads();

    const int wpg = DYN ? p.warps_per_group : 1;
    const int group_in_cta = warp / wpg;
    const int wsub = warp - group_in_cta * wpg;
    const int groups_per_cta = n_warps / wpg;
    // cluster launch (DYN, one group per CTA): the group is the cluster, this CTA takes every csize-th chunk of a level
    const int csize = DYN ? p.cluster_size : 1;
    const int crank = (DYN && csize > 1) ? int(cluster_ctarank()) : 0;
    const int64_t batch0 = csize > 1 ? int64_t(blockIdx.x) / csize : int64_t(blockIdx.x) * groups_per_cta + group_in_cta;
    if (batch0 * kQuadSamples >= p.n) return;  // whole group (all its warps, all CTAs of its cluster) out of range
    const PhiloxKeys& key0 = p.keys;

    // The lane's four sample columns.  ld is a multiple of 64, not necessarily of 128: in the last group the
    // lanes of the upper half may then lie past the row end.  They shadow the lower half (same columns, same
    // seeds, hence the same values written twice) and stay out of the statistics.
    const int64_t seg0 = batch0 * kQuadSamples;
    const bool half_row = seg0 + 64 >= p.ld;
    int64_t s0 = seg0 + 4 * lane;
    const bool shadow = s0 >= p.ld;
    if (shadow) s0 -= 64;
    Seeds4 sd;
#pragma unroll
    for (int i = 0; i < 4; ++i) sd.s[i] = uint32_t(i);
    if constexpr (MODE != kModeInjected) {
        if (p.seeds) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                sd.s[i] = s0 + i < p.n ? uint32_t(__ldg(p.seeds + s0 + i)) : (i ? sd.s[i - 1] + 1u : 0u);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) sd.s[i] = uint32_t(p.seed0) + uint32_t(s0) + uint32_t(i);
        }
    }
    sd.quad = ((sd.s[0] & 1u) == 0u) && sd.s[1] == sd.s[0] + 1u && sd.s[2] == sd.s[0] + 2u && sd.s[3] == sd.s[0] + 3u;
    sd.quad4 = sd.quad && (sd.s[0] & 3u) == 0u;
    const uint32_t lane_off = uint32_t(s0) * 8u;
    uint64_t lane_off64 = lane_off;  // kept as a register pair: the 64-bit addend of the row multiply-add
    asm volatile("" : "+l"(lane_off64));
    auto f64_row = [&](const void* base, uint32_t row) -> char* {
        return const_cast<char*>(reinterpret_cast<const char*>(base)) + (uint64_t(row) * p.ldb8 + lane_off64);
    };
    auto i32_row = [&](const void* base, uint32_t row) -> char* {
        return const_cast<char*>(reinterpret_cast<const char*>(base)) + (uint64_t(row) * p.ldb4 + (lane_off >> 1));
    };
    const void* const dur_base = MODE == kModeInjected ? static_cast<const void*>(p.inj) : static_cast<const void*>(p.durations);
    // L2 prefetch of gather rows: lane l covers one 128-byte line of the warp's row segment (8 lines, 4 when
    // only half of the segment lies inside the row)
    const uint32_t pf_off = uint32_t(seg0) * 8u + (half_row ? uint32_t(lane & 3) : uint32_t(lane & 7)) * 128u;
    const uint32_t pf_unit = half_row ? uint32_t(lane >> 2) : uint32_t(lane >> 3);
    const uint32_t pf_step = half_row ? 8u : 4u;

    // ---- chunk pipeline ----
    uint32_t buf_sel = 0u, phase = 0u;
    auto issue = [&](int c, uint32_t b) {
        if (lane == 0) {
            mbar_expect_tx(bar0 + b * 8u, uint32_t(kChunkBytes));
            bulk_copy_g2s(ring0 + b * uint32_t(kChunkBytes), p.chunks + size_t(c) * kChunkUnits, uint32_t(kChunkBytes),
                          bar0 + b * 8u);
        }
    };

    // ---- level cursor (DYN): counts the chunks this CTA has taken in the current level; chunk = lb + k * csize + crank ----
    __shared__ int s_cursor[16][2];
    if constexpr (DYN) {
        if (wsub == 0 && lane == 0) s_cursor[group_in_cta][0] = 0;
        group_barrier(1 + group_in_cta, wpg * 32);
    }
    int seq = 0, lb = 0;
    auto grab = [&](int parity) -> int {
        if constexpr (!DYN) {
            return seq++;
        } else {
            int v = 0;
            if (lane == 0) v = atomicAdd(&s_cursor[group_in_cta][parity], 1);
            return lb + __shfl_sync(0xFFFFFFFFu, v, 0) * csize + crank;
        }
    };

    // ---- the running event of this warp ----
    bool open = false;
    uint32_t row = 0u;
    double lat[4] = {0.0, 0.0, 0.0, 0.0}, ub = 0.0;  // lat: the running UNCLAMPED arrival of the open event (see relax)
    int cause[4] = {-1, -1, -1, -1};
    D4 nrs{0.0, 0.0, 0.0, 0.0};
    uint32_t ev = 0u;
    double ev_earliest = 0.0;
    bool valid[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) valid[i] = !shadow && s0 + i < p.n;
    // ---- reduced modes: statistics staged per warp, folded every kStatSlots events ----
    int n_staged = 0;      // events parked in the staging area (warp-uniform)
    uint32_t my_ev = 0u;   // lane k < n_staged: event id of slot k
    const bool all_valid = valid[0] && valid[1] && valid[2] && valid[3];
    auto flush_stats = [&]() {
        __syncwarp();
        // Sums: lane -> (slot e = lane & 3, lane block = lane >> 2): four parked {sum, sumsq} pairs each, then three
        // shuffle stages over the eight blocks.  The element order inside a block is rotated by e so that the eight
        // lanes of a quarter warp read eight different 16-byte bank groups.
#if MCDP_STAGE_SUMS
        const int e = lane & 3, blk = lane >> 2;
        double as = 0.0, aq = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double a, b;
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];"
                         : "=d"(a), "=d"(b)
                         : "r"(stat0 + uint32_t(e * 32 + blk * 4 + ((k + e) & 3)) * 16u));
            as += a;
            aq += b;
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            as += __shfl_xor_sync(0xFFFFFFFFu, as, o);
            aq += __shfl_xor_sync(0xFFFFFFFFu, aq, o);
        }
        if (lane < n_staged) {  // lane == e here: its own slot's event
            if (p.sum) atomicAdd(p.sum + my_ev, as);
            if (p.sumsq) atomicAdd(p.sumsq + my_ev, aq);
        }
#endif
        if (stage_hist) {
            const int nb = p.n_bins;
#pragma unroll
            for (int k = 0; k < kStatSlots; ++k) {
                const uint32_t evk = __shfl_sync(0xFFFFFFFFu, my_ev, k);
                if (k < n_staged) {
                    const uint32_t wa = stat0 + uint32_t(kStatSumBytes) + uint32_t(k * kStatHistReplicas * 32 + lane) * 4u;
                    uint32_t w = 0u;
#pragma unroll
                    for (int rep = 0; rep < kStatHistReplicas; ++rep) {
                        uint32_t wr;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wr) : "r"(wa + uint32_t(rep) * 128u));
                        w += wr;
                    }
                    if (w) {
#pragma unroll
                        for (int rep = 0; rep < kStatHistReplicas; ++rep)
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(wa + uint32_t(rep) * 128u), "r"(0u) : "memory");
                        uint32_t* h = p.hist + size_t(evk) * nb + 2 * lane;
                        if (w & 0xFFFFu) atomicAdd(h, w & 0xFFFFu);
                        if (w >> 16) atomicAdd(h + 1, w >> 16);
                    }
                }
            }
        }
        __syncwarp();
        n_staged = 0;
    };
    auto finalize = [&]() {
        // _core.cpp:348-349
        double r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = ref_min(lat[i], ub);
        stcg_d4(f64_row(p.realized, row), D4{r[0], r[1], r[2], r[3]});
        open = false;
        if constexpr (!DYN) {
#pragma unroll
            for (int i = 0; i < 4; ++i) lat[i] = r[i];
        }
        if constexpr (!kReduced) {
            __stcs(reinterpret_cast<int4*>(i32_row(p.cause, row)), make_int4(cause[0], cause[1], cause[2], cause[3]));
        } else {
            double x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = valid[i] ? r[i] - ev_earliest : 0.0;
            // sum and sum of squares of this lane's samples: parked, folded across lanes in flush_stats
            {
                const double sl = (x[0] + x[1]) + (x[2] + x[3]);
                const double ql = (x[0] * x[0] + x[1] * x[1]) + (x[2] * x[2] + x[3] * x[3]);
#if MCDP_STAGE_SUMS
                asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(stat0 + uint32_t(n_staged * 32 + lane) * 16u), "d"(sl), "d"(ql)
                             : "memory");
#else
                flush_sums(p, ev, lane, sl, ql, 0u);
#endif
            }
            if (p.late) {
                // exceedance counts, 8 bits per threshold (at most 128 samples per warp): one REDUX, the atomics of the
                // thresholds issued by different lanes of one instruction.  Straight-line over the lane's four samples;
                // only a lane of the ragged last group (some of its columns are padding) takes the checked path.
                uint32_t late_packed = 0u;
                if (all_valid) {
#pragma unroll
                    for (int t = 0; t < MCDP_MAX_THRESHOLDS; ++t) {
                        if (t < p.n_thresholds) {
                            const double th = p.thresholds[t];
#pragma unroll
                            for (int i = 0; i < 4; ++i) add_if_gt(late_packed, x[i], th, 1u << (8 * t));
                        }
                    }
                } else {
                    for (int t = 0; t < p.n_thresholds; ++t)
#pragma unroll
                        for (int i = 0; i < 4; ++i) late_packed += uint32_t(valid[i] && x[i] > p.thresholds[t]) << (8 * t);
                }
                const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, late_packed);
                if (lane < p.n_thresholds) {
                    const uint32_t cnt = (tot >> (8 * lane)) & 0xFFu;
                    if (cnt) atomicAdd(p.late + size_t(lane) * p.E + ev, (unsigned long long)cnt);
                }
            }
            if (p.hist) {
                const int nb = p.n_bins;
                int b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) b[i] = min(max(int(floor((x[i] - p.hist_lo) * p.hist_scale)), 0), nb - 1);
                if (stage_hist) {
                    // counts into this warp's private 16-bit-pair histogram of the slot (shared-memory reductions)
                    const uint32_t h0 = stat0 + uint32_t(kStatSumBytes) + uint32_t(n_staged * kStatHistReplicas + (lane % kStatHistReplicas)) * 128u;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        atoms_add_u32(h0 + uint32_t(b[i] >> 1) * 4u, valid[i] ? 1u << ((b[i] & 1) * 16) : 0u);
                } else {
                    uint32_t* h = p.hist + size_t(ev) * nb;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int key = valid[i] ? b[i] : -1 - lane;  // unique keys: match groups of size 1, skipped below
                        const unsigned g = __match_any_sync(0xFFFFFFFFu, key);
                        if (key >= 0 && lane == __ffs(g) - 1) atomicAdd(h + key, uint32_t(__popc(g)));
                    }
                }
            }
            if (lane == n_staged) my_ev = ev;
            if (++n_staged == kStatSlots) flush_stats();
            if constexpr (MODE == kModeAttr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int w = valid[i] ? cause[i] : -4 - lane;
                    const unsigned g = __match_any_sync(0xFFFFFFFFu, w);
                    if (lane == __ffs(g) - 1) {
                        if (w >= 0) atomicAdd(p.cause_act + w, (unsigned long long)__popc(g));
                        else if (w == -1) atomicAdd(p.cause_none + ev, (unsigned long long)__popc(g));
                    }
                }
            }
        }
    };
    // An event stays open until the next header or END unit this warp meets -- the first unit of its next chunk, or
    // the END unit it is handed when its level has no more chunks for it (below) -- so the close (finalize: stores,
    // statistics) is inlined ONCE: two copies of it pushed the reduced kernels out of the instruction caches.
    auto process = [&](uint32_t buf, int u0) {
#pragma unroll kQuadUnitUnroll
        for (int u = u0; u < kChunkUnits; ++u) {
            const int4 q0 = lds128(buf + uint32_t(u) * 32u);
            const int4 q1 = lds128(buf + uint32_t(u) * 32u + 16u);
            const uint32_t meta = uint32_t(q1.x);
            const uint32_t kind = meta >> 29;
            // the realized row of the NEXT entry unit is requested before this unit's delays are drawn
            const D4 rs = nrs;
            if (DYN || kind < kKindEvent) {
                if (uint32_t(q1.z) != kNoRow) nrs = ldcg_d4(f64_row(p.realized, uint32_t(q1.z)));
            }
            if (kind >= kKindEvent) {
                if (open) finalize();
                if (kind == kKindEnd) break;
                if constexpr (!DYN) {
                    if (q1.w != 0) nrs = D4{lat[0], lat[1], lat[2], lat[3]};
                    else if (uint32_t(q1.z) != kNoRow) nrs = ldcg_d4(f64_row(p.realized, uint32_t(q1.z)));
                }
                // _core.cpp:333-337
                row = uint32_t(q0.x);
                const double earliest = __hiloint2double(q0.w, q0.z);
                if constexpr (kReduced) {
                    ev = uint32_t(q0.y);
                    ev_earliest = earliest;
                }
                ub = __dadd_rn(earliest, p.max_delay);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    lat[i] = earliest;
                    cause[i] = -1;
                }
                open = true;
                continue;
            }
            const uint32_t act = uint32_t(q0.y);
            const double base = __hiloint2double(q0.w, q0.z);
            // duration known: stream it out (full mode) and apply the recurrence, _core.cpp:341-346
            auto apply = [&](const double (&d)[4]) {
                if constexpr (MODE == kModeFull) {
                    if (act != kNoAct) stcs_d4(f64_row(dur_base, act), D4{d[0], d[1], d[2], d[3]});
                }
                const int src_event = MODE == kModeAttr ? (act == kNoAct ? -3 : int(act)) : q0.x;
                const double rsv[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#if MCDP_RELAX_LITERAL
                    const double t = ref_min(__dadd_rn(rsv[i], d[i]), ub);
                    if (t >= lat[i]) {
                        lat[i] = t;
                        cause[i] = src_event;
                    }
#else
                    relax(__dadd_rn(rsv[i], d[i]), ub, lat[i], cause[i], src_event);
#endif
                }
            };
            if constexpr (MODE == kModeInjected) {
                D4 dd{0.0, 0.0, 0.0, 0.0};
                if (act != kNoAct) dd = ldcs_d4(f64_row(dur_base, act));
                const double d[4] = {dd.x, dd.y, dd.z, dd.w};
                apply(d);
            } else if (kind == kKindNone) {
                // _core.cpp:304-305,325.  A path of its own down to the recurrence: as a two-way merge in front of a
                // shared tail, the four register copies of `base` were hoisted above the kind dispatch and cost every
                // sampled entry 8 instructions.
                const double d[4] = {base, base, base, base};
                apply(d);
            } else {
                double e[4], d[4];
                sample_extra4<SMEM, kReduced>(meta, uint32_t(q1.y), dists, uint32_t(q1.w), tab, base, act, sd, key0, log_tab, e);
#pragma unroll
                for (int i = 0; i < 4; ++i) d[i] = __dadd_rn(base, e[i]);  // _core.cpp:328
                apply(d);
            }
        }
    };

    // ---- main loop ----
    const int n_rounds = DYN ? p.n_levels : (p.n_chunks > 0 ? 1 : 0);
    for (int lvl = 0; lvl < n_rounds; ++lvl) {
        const int le = DYN ? __ldg(p.chunk_level_begin + lvl + 1) : p.n_chunks;
        const int par = lvl & 1;
        if constexpr (DYN) {
            if (wsub == 0 && lane == 0) s_cursor[group_in_cta][par ^ 1] = 0;  // re-armed for the next level
        }
        int c = grab(par);
        bool from_cursor = true;
        if (c < le) issue(c, buf_sel);
        for (;;) {
            const bool tail = c >= le;  // no chunk left for this warp in this level: close what is open, through the END unit
            if (tail && !open) break;
            uint32_t buf = end_buf;
            bool is_cont = false, skip = false;
            int cn = c;
            if (!tail) {
                buf = ring0 + buf_sel * uint32_t(kChunkBytes);
                mbar_wait(bar0 + buf_sel * 8u, (phase >> buf_sel) & 1u);
                phase ^= 1u << buf_sel;
                buf_sel ^= 1u;
                const int4 h1 = lds128(buf + 16u);
                is_cont = (uint32_t(h1.x) >> 29) == kKindEnd;
                skip = DYN && from_cursor && is_cont;
                const uint32_t remaining = skip ? 0u : uint32_t(h1.y);
                from_cursor = remaining == 0u;
                if (from_cursor) {
                    cn = grab(par);
                } else {
                    cn = c + 1;
                    if constexpr (!DYN) ++seq;
                }
                if (cn < le) issue(cn, buf_sel);
                if (!skip) {
                    for (uint32_t uu = pf_unit; uu < uint32_t(kChunkUnits); uu += pf_step) {
                        const uint32_t ua = buf + uu * 32u;
                        uint32_t src, meta;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(src) : "r"(ua));
                        asm volatile("ld.shared.u32 %0, [%1+16];" : "=r"(meta) : "r"(ua));
                        if ((meta >> 29) < kKindEvent)
                            prefetch_l2(reinterpret_cast<const char*>(p.realized) + (uint64_t(src) * p.ldb8 + pf_off));
                    }
                }
            }
            if (!skip) process(buf, is_cont ? 1 : 0);
            __syncwarp();
            if (tail) break;
            c = cn;
        }
        if constexpr (DYN) {
            lb = le;
            if (csize > 1) cluster_barrier();
            else group_barrier(1 + group_in_cta, wpg * 32);
        }
    }
    if constexpr (kReduced) {
        if (n_staged) flush_stats();
    }

    if constexpr (MODE == kModeFull) {
        // activities no precedence entry references still get their sampled duration (_core.cpp:323-329)
        for (int i = wsub + crank * wpg; i < p.n_orphans; i += wpg * csize) {
            const int4 q0 = __ldg(reinterpret_cast<const int4*>(p.orphans + i));
            const int4 q1 = __ldg(reinterpret_cast<const int4*>(p.orphans + i) + 1);
            const uint32_t act = uint32_t(q0.y), meta = uint32_t(q1.x);
            const double base = __hiloint2double(q0.w, q0.z);
            double d[4] = {base, base, base, base};
            if ((meta >> 29) != kKindNone) {
                double e[4];
                sample_extra4<SMEM, kReduced>(meta, uint32_t(q1.y), dists, uint32_t(q1.w), tab, base, act, sd, key0, log_tab, e);
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = __dadd_rn(base, e[k]);
            }
            stcs_d4(f64_row(dur_base, act), D4{d[0], d[1], d[2], d[3]});
        }
    }
}

}  // namespace mcdp
