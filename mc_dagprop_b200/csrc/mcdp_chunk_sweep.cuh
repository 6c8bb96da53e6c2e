// mcdp_chunk_sweep.cuh -- fused sample + max-plus sweep over the chunk stream (full, duration-injection
// and single-batch reduced-statistics modes), sm_100a.
//
// Replaces Simulator::run (reference _core.cpp:312-353) for a block of seeds.  Same data layout
// and work split as mcdp_sweep.cuh (a warp owns 64 adjacent samples, two per lane;
// `warps_per_group` warps share them and split every topological level), but the plan arrives as
// the chunk stream of mcdp_records.h:
//   * a warp pulls 512-byte chunks (whole events with their precedence entries) and has the NEXT
//     chunk copied into its private shared-memory ring by the bulk-copy engine (cp.async.bulk +
//     mbarrier complete_tx) while it works on the current one, so the warp-uniform record words
//     are shared-memory broadcasts instead of dependent global loads;
//   * the units of a chunk are walked by one flat loop -- event headers open / close the running
//     max, entry units draw their delay and apply the recurrence -- and every entry unit names the
//     source row of the next one (across events), so the gather of realized[pred] is always one
//     unit ahead of its use;
//   * warps that split a level take chunks from a shared-memory cursor and meet at a named barrier
//     per level; the chunks of a level shrink towards its end so the warps arrive close together.
#pragma once
#include "mcdp_sweep.cuh"

namespace mcdp {

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return uint32_t(__cvta_generic_to_shared(ptr)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine, completion counted on `bar`
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ int4 lds128(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
// base + row * stride as ONE 32x32+64 multiply-add (the lane's column base stays in registers)
__device__ __forceinline__ char* row_ptr(char* base, uint32_t row, uint32_t stride) {
    char* r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(row), "r"(stride), "l"(base));
    return r;
}

// dynamic shared memory behind the staged tables: per warp a two-chunk ring followed by its two mbarriers
constexpr int kRingStride = 2 * kChunkBytes + 16;
__host__ __device__ inline size_t chunk_ring_bytes(int n_warps) { return size_t(n_warps) * size_t(kRingStride); }

template <int MODE, bool SMEM, bool DYN>
__global__ void __launch_bounds__(MCDP_MAX_THREADS, MCDP_MIN_BLOCKS)
    chunk_sweep_kernel(const __grid_constant__ SweepParams p) {
    constexpr bool kReduced = MODE == kModeReduced || MODE == kModeAttr;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    // Dynamic shared memory: [log table][DistRec[] + table pool (SMEM)][per-warp chunk rings].  Everything
    // is addressed through ONE shared-window base register (kept opaque so that it is not re-derived
    // from special registers at every use) plus offsets the host put into the parameter block.
    uint32_t smem_base = smem_u32(smem_dyn);
    asm volatile("" : "+r"(smem_base));
    for (int i = threadIdx.x; i < kLogTabEntries; i += blockDim.x)
        reinterpret_cast<int4*>(smem_dyn)[i] = __ldg(reinterpret_cast<const int4*>(p.log_tab) + i);
    const uint32_t log_tab = smem_base;
    typename Mem<SMEM>::ptr dists, tab;
    if constexpr (SMEM) {
        // stage distribution records + guide / inverse-CDF tables once per CTA
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
    // this warp's ring: two 512-byte chunk buffers followed by their two mbarriers
    const uint32_t ring0 = smem_base + p.smem_ring_off + uint32_t(warp) * uint32_t(kRingStride);
    const uint32_t bar0 = ring0 + 2u * uint32_t(kChunkBytes);
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int wpg = DYN ? p.warps_per_group : 1;
    const int group_in_cta = warp / wpg;
    const int wsub = warp - group_in_cta * wpg;
    const int groups_per_cta = n_warps / wpg;
    const int64_t batch0 = int64_t(blockIdx.x) * groups_per_cta + group_in_cta;
    if (batch0 * 64 >= p.n) return;  // whole group (all its warps) out of range
    const PhiloxKeys& key0 = p.keys;

    // the lane's two sample columns (ld is a multiple of 64: both are in-bounds padding at worst)
    const int64_t s0 = batch0 * 64 + 2 * lane;
    uint32_t seed_a = 0u, seed_b = 1u;
    if constexpr (MODE != kModeInjected) {
        if (p.seeds) {
            seed_a = s0 < p.n ? uint32_t(__ldg(p.seeds + s0)) : 0u;
            seed_b = s0 + 1 < p.n ? uint32_t(__ldg(p.seeds + s0 + 1)) : seed_a + 1u;
        } else {
            seed_a = uint32_t(p.seed0) + uint32_t(s0);
            seed_b = seed_a + 1u;
        }
    }
    const bool paired = ((seed_a & 1u) == 0u) && (seed_b == seed_a + 1u);
    // Row addressing: array base (uniform, from the parameter block) + row * ld * 8 + the lane's column
    // offset, which fits 32 bits (ld < 2^29): one register of per-lane addressing state for all arrays.
    const uint32_t lane_off = uint32_t(s0) * 8u;
    auto f64_row = [&](const void* base, uint32_t row) -> char* {
        return const_cast<char*>(reinterpret_cast<const char*>(base)) + (uint64_t(row) * p.ldb8 + lane_off);
    };
    auto i32_row = [&](const void* base, uint32_t row) -> char* {
        return const_cast<char*>(reinterpret_cast<const char*>(base)) + (uint64_t(row) * p.ldb4 + (lane_off >> 1));
    };
    const void* const dur_base = MODE == kModeInjected ? static_cast<const void*>(p.inj) : static_cast<const void*>(p.durations);

    // ---- chunk pipeline: the ring's buffers alternate; `phase` holds the mbarrier parity of each ----
    uint32_t buf_sel = 0u, phase = 0u;
    auto issue = [&](int c, uint32_t b) {  // bulk-copy chunk c into buffer b
        if (lane == 0) {
            mbar_expect_tx(bar0 + b * 8u, uint32_t(kChunkBytes));
            bulk_copy_g2s(ring0 + b * uint32_t(kChunkBytes), p.chunks + size_t(c) * kChunkUnits, uint32_t(kChunkBytes),
                          bar0 + b * 8u);
        }
    };

    // ---- level cursor (DYN): one shared-memory counter per level parity ---------------------------
    __shared__ int s_cursor[16][2];
    if constexpr (DYN) {
        if (wsub == 0 && lane == 0) s_cursor[group_in_cta][0] = 0;
        group_barrier(1 + group_in_cta, wpg * 32);
    }
    int seq = 0;  // static mode: chunks in stream order
    auto grab = [&](int parity) -> int {
        if constexpr (!DYN) {
            return seq++;
        } else {
            int v = 0;
            if (lane == 0) v = atomicAdd(&s_cursor[group_in_cta][parity], 1);
            return __shfl_sync(0xFFFFFFFFu, v, 0);
        }
    };

    // ---- the running event of this warp ---------------------------------------------------------
    bool open = false;
    uint32_t row = 0u;
    double lat_a = 0.0, lat_b = 0.0, ub = 0.0;
    int cause_a = -1, cause_b = -1;
    double2 nrs = make_double2(0.0, 0.0);
    uint32_t ev = 0u;        // reduced mode: event id (row is a recycled scratch slot there)
    double ev_earliest = 0.0;
    const bool valid_a = s0 < p.n, valid_b = s0 + 1 < p.n;
    auto finalize = [&]() {
        // _core.cpp:348-349.  realized rows are gathered by other warps: L2 only (st.cg / ld.cg)
        const double ra = ref_min(lat_a, ub), rb = ref_min(lat_b, ub);
        __stcg(reinterpret_cast<double2*>(f64_row(p.realized, row)), make_double2(ra, rb));
        open = false;
        if constexpr (!DYN) {  // a following header may take the closed event's value over (forward)
            lat_a = ra;
            lat_b = rb;
        }
        if constexpr (!kReduced) {
            __stcs(reinterpret_cast<int2*>(i32_row(p.cause, row)), make_int2(cause_a, cause_b));
        } else {
            // per-event statistics of the delay realized - earliest over this warp's 64 samples: one
            // atomic per statistic (and per distinct histogram bin) per warp
            const double xa = valid_a ? ra - ev_earliest : 0.0;
            const double xb = valid_b ? rb - ev_earliest : 0.0;
            uint32_t late_packed = 0u;
            if (p.late) {
#pragma unroll
                for (int t = 0; t < MCDP_MAX_THRESHOLDS; ++t) {
                    if (t < p.n_thresholds)
                        late_packed |= (uint32_t(valid_a && xa > p.thresholds[t]) + uint32_t(valid_b && xb > p.thresholds[t])) << (8 * t);
                }
            }
            flush_sums(p, ev, lane, xa + xb, xa * xa + xb * xb, late_packed);
            if (p.hist) {
                const int nb = p.n_bins;
                int ba = min(max(int(floor((xa - p.hist_lo) * p.hist_scale)), 0), nb - 1);
                int bb = min(max(int(floor((xb - p.hist_lo) * p.hist_scale)), 0), nb - 1);
                if (!valid_a) ba = -1 - lane;  // unique keys: match groups of size 1, skipped below
                if (!valid_b) bb = -1 - lane;
                uint32_t* h = p.hist + size_t(ev) * nb;
                const unsigned ga = __match_any_sync(0xFFFFFFFFu, ba);
                if (ba >= 0 && lane == __ffs(ga) - 1) atomicAdd(h + ba, uint32_t(__popc(ga)));
                const unsigned gb = __match_any_sync(0xFFFFFFFFu, bb);
                if (bb >= 0 && lane == __ffs(gb) - 1) atomicAdd(h + bb, uint32_t(__popc(gb)));
            }
            if constexpr (MODE == kModeAttr) {
                // delay-cause attribution: cause_* hold the activity index of the entry that decided the event
                // (-1 none, -3 an entry without a duration row); equal winners across the warp share one atomic
                const int wa = valid_a ? cause_a : -4 - lane, wb = valid_b ? cause_b : -4 - lane;
                const unsigned ga = __match_any_sync(0xFFFFFFFFu, wa);
                if (lane == __ffs(ga) - 1) {
                    if (wa >= 0) atomicAdd(p.cause_act + wa, (unsigned long long)__popc(ga));
                    else if (wa == -1) atomicAdd(p.cause_none + ev, (unsigned long long)__popc(ga));
                }
                const unsigned gb = __match_any_sync(0xFFFFFFFFu, wb);
                if (lane == __ffs(gb) - 1) {
                    if (wb >= 0) atomicAdd(p.cause_act + wb, (unsigned long long)__popc(gb));
                    else if (wb == -1) atomicAdd(p.cause_none + ev, (unsigned long long)__popc(gb));
                }
            }
        }
    };
    auto process = [&](uint32_t buf, int u0, uint32_t remaining) {
#pragma unroll 1
        for (int u = u0; u < kChunkUnits; ++u) {
            const int4 q0 = lds128(buf + uint32_t(u) * 32u);
            const int4 q1 = lds128(buf + uint32_t(u) * 32u + 16u);
            const uint32_t meta = uint32_t(q1.x);
            const uint32_t kind = meta >> 29;
            // The realized row of the NEXT entry unit is requested before this unit's delay is drawn (q1.z:
            // PredRec::next_src_row / HeaderUnit::first_src_row).  In a dense stream (!DYN) a header may ask for
            // the row of the event that is closed at this very header: there the request follows the close.
            const double2 rs = nrs;
            if (DYN || kind < kKindEvent) {
                if (uint32_t(q1.z) != kNoRow)
                    nrs = __ldcg(reinterpret_cast<const double2*>(f64_row(p.realized, uint32_t(q1.z))));
            }
            if (kind >= kKindEvent) {
                if (open) finalize();
                if (kind == kKindEnd) break;
                if constexpr (!DYN) {
                    // the first entry's source: taken over from the event this warp has just closed
                    // (HeaderUnit::pad), or requested now
                    if (q1.w != 0) nrs = make_double2(lat_a, lat_b);
                    else if (uint32_t(q1.z) != kNoRow)
                        nrs = __ldcg(reinterpret_cast<const double2*>(f64_row(p.realized, uint32_t(q1.z))));
                }
                // _core.cpp:333-337
                row = uint32_t(q0.x);
                const double earliest = __hiloint2double(q0.w, q0.z);
                if constexpr (kReduced) {
                    ev = uint32_t(q0.y);
                    ev_earliest = earliest;
                }
                ub = __dadd_rn(earliest, p.max_delay);
                lat_a = lat_b = earliest;
                cause_a = cause_b = -1;
                open = true;
                continue;
            }
            const uint32_t act = uint32_t(q0.y);
            const double base = __hiloint2double(q0.w, q0.z);
            double da, db;
            if constexpr (MODE == kModeInjected) {
                double2 dd = make_double2(0.0, 0.0);
                if (act != kNoAct) dd = __ldcs(reinterpret_cast<const double2*>(f64_row(dur_base, act)));
                da = dd.x;
                db = dd.y;
            } else {
                if (kind == kKindNone) {
                    da = db = base;  // _core.cpp:304-305,325
                } else {
                    double ea, eb;
                    sample_extra2<SMEM>(meta, uint32_t(q1.y), dists, uint32_t(q1.w), tab, base, act, seed_a, seed_b, paired,
                                        key0, log_tab, ea, eb);
                    da = __dadd_rn(base, ea);  // _core.cpp:328
                    db = __dadd_rn(base, eb);
                }
                if constexpr (MODE == kModeFull) {
                    if (act != kNoAct) __stcs(reinterpret_cast<double2*>(f64_row(dur_base, act)), make_double2(da, db));
                }
            }
            // _core.cpp:341-346
            const double ta = ref_min(__dadd_rn(rs.x, da), ub);
            const double tb = ref_min(__dadd_rn(rs.y, db), ub);
            // what cause_event reports: the source event; in attribution mode the deciding entry's activity
            const int src_event = MODE == kModeAttr ? (act == kNoAct ? -3 : int(act)) : q0.x;
            if (ta >= lat_a) {
                lat_a = ta;
                cause_a = src_event;
            }
            if (tb >= lat_b) {
                lat_b = tb;
                cause_b = src_event;
            }
        }
        if (open && remaining == 0u) finalize();
    };

    // ---- main loop: level by level; inside a level the group's warps pull chunks from the cursor ----
    const int n_rounds = DYN ? p.n_levels : (p.n_chunks > 0 ? 1 : 0);
    for (int lvl = 0; lvl < n_rounds; ++lvl) {
        const int le = DYN ? __ldg(p.chunk_level_begin + lvl + 1) : p.n_chunks;
        const int par = lvl & 1;
        if constexpr (DYN) {
            if (wsub == 0 && lane == 0) s_cursor[group_in_cta][par ^ 1] = le;  // next level starts at le
        }
        int c = grab(par);
        bool from_cursor = true;
        if (c < le) issue(c, buf_sel);
        while (c < le) {
            const uint32_t buf = ring0 + buf_sel * uint32_t(kChunkBytes);
            mbar_wait(bar0 + buf_sel * 8u, (phase >> buf_sel) & 1u);
            phase ^= 1u << buf_sel;
            buf_sel ^= 1u;
            const int4 h1 = lds128(buf + 16u);
            const bool is_cont = (uint32_t(h1.x) >> 29) == kKindEnd;
            const bool skip = DYN && from_cursor && is_cont;  // somebody else's continuation chunk
            const uint32_t remaining = skip ? 0u : uint32_t(h1.y);
            // the next chunk: the continuation of this chunk's long event, else one from the cursor
            from_cursor = remaining == 0u;
            int cn;
            if (from_cursor) {
                cn = grab(par);
            } else {
                cn = c + 1;
                if constexpr (!DYN) ++seq;
            }
            if (cn < le) issue(cn, buf_sel);
            if (!skip) {
                // Pull the realized rows this chunk gathers towards L2 now (the gather proper runs only one
                // unit ahead of its use).  Lane l covers the 128-byte line (l & 3) of this warp's 512-byte
                // row segment for the entry units (l >> 2), 8 + (l >> 2), ...
#pragma unroll
                for (int h = 0; h < kChunkUnits / 8; ++h) {
                    const uint32_t ua = buf + (uint32_t(lane >> 2) + 8u * uint32_t(h)) * 32u;
                    uint32_t src, meta;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(src) : "r"(ua));
                    asm volatile("ld.shared.u32 %0, [%1+16];" : "=r"(meta) : "r"(ua));
                    if ((meta >> 29) < kKindEvent)
                        prefetch_l2(f64_row(p.realized, src) - lane * 16 + (lane & 3) * 128);
                }
                process(buf, is_cont ? 1 : 0, remaining);
            }
            __syncwarp();  // every lane is done with `buf` before a later bulk copy may overwrite it
            c = cn;
        }
        if constexpr (DYN) group_barrier(1 + group_in_cta, wpg * 32);
    }

    if constexpr (MODE == kModeFull) {
        // activities no precedence entry references still get their sampled duration (_core.cpp:323-329)
        for (int i = wsub; i < p.n_orphans; i += wpg) {
            const int4 q0 = __ldg(reinterpret_cast<const int4*>(p.orphans + i));
            const int4 q1 = __ldg(reinterpret_cast<const int4*>(p.orphans + i) + 1);
            const uint32_t act = uint32_t(q0.y), meta = uint32_t(q1.x);
            const double base = __hiloint2double(q0.w, q0.z);
            double da = base, db = base;
            if ((meta >> 29) != kKindNone) {
                double ea, eb;
                sample_extra2<SMEM>(meta, uint32_t(q1.y), dists, uint32_t(q1.w), tab, base, act, seed_a, seed_b, paired, key0,
                                    log_tab, ea, eb);
                da = __dadd_rn(base, ea);
                db = __dadd_rn(base, eb);
            }
            __stcs(reinterpret_cast<double2*>(f64_row(dur_base, act)), make_double2(da, db));
        }
    }
}

}  // namespace mcdp
