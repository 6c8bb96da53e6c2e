// mcdp_capi.cu -- the C ABI of include/mcdp_b200.h: plan upload, kernel launches, and the
// chunked host-buffer calls.  No torch types, no exceptions across the boundary, no CPU fallback.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <system_error>
#include <unordered_map>
#include <vector>

#include "../../include/mcdp_b200.h"
#include "mcdp_plan.hpp"
#include "mcdp_compat.cuh"
#include "mcdp_chunk_sweep.cuh"
#include "mcdp_quad_sweep.cuh"
#include "mcdp_small_sweep.cuh"

using namespace mcdp;

static_assert(MCDP_CHUNK_UNITS == kChunkUnits, "header and records disagree on the chunk size");

namespace {

thread_local std::string g_err;

int32_t fail(int32_t code, const std::string& msg) {
    g_err = msg;
    return code;
}

}  // namespace

// the same thread-local message for the other translation units of the library (mcdp_analytic.cu)
int32_t mcdp_set_error(int32_t code, const std::string& msg) { return fail(code, msg); }

namespace {

#define MCDP_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(MCDP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
    } while (0)

// like MCDP_CUDA inside a chunk loop: record the failure and leave the loop, so that the code behind the loop
// still drains the streams (copies into caller memory may be in flight)
#define MCDP_CUDA_BRK(expr)                                                                         \
    {                                                                                               \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            rc = fail(MCDP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
            break;                                                                                  \
        }                                                                                           \
    }

// NVTX range (plan compile / H2D / kernel / D2H / reduce); a no-op unless a tool is attached
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (dev < 0) return;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Buffers of one in-flight chunk of a *_host call.
struct HostSlot {
    cudaStream_t stream = nullptr;
    DevBuf<int32_t> seeds;
    DevBuf<double> realized, durations;  // event-major
    DevBuf<int32_t> cause;
    DevBuf<double> t_realized, t_durations;  // sample-major staging
    DevBuf<int32_t> t_cause;
    void release() {
        seeds.release();
        realized.release();
        durations.release();
        cause.release();
        t_realized.release();
        t_durations.release();
        t_cause.release();
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

}  // namespace

struct mcdp_plan {
    // the compiled plan on the host: shared by the plans of a plan set (one compile, one copy per device)
    std::shared_ptr<HostPlan> host_sp;
    HostPlan& host;
    explicit mcdp_plan(std::shared_ptr<HostPlan> h) : host_sp(std::move(h)), host(*host_sp) {}
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    // device-resident stream
    // event + precedence records live in ONE allocation so that a single L2 access-policy window
    // (persisting) covers the whole stream: 128 GB of streaming output per launch would otherwise
    // keep evicting the records every warp re-reads
    // chunk streams [rows: 0 event ids (full / injected), 1 scratch slots (reduced)][0 level-aligned, 1 dense]
    struct ChunkStreamDev {
        DevBuf<ChunkUnit> units;
        DevBuf<int32_t> level_begin;
        int32_t n_chunks = 0;
        bool ready = false;
    } streams[2][2];
    std::vector<EventRec> ev_red;         // reduced modes: host copies of the slot-row records (built once)
    std::vector<PredRec> pr_red;
    size_t l2_persist_bytes = 0, l2_window_max = 0;
    DevBuf<PredRec> d_orphans;
    DevBuf<DistRec> d_dists;
    DevBuf<double> d_tab;
    DevBuf<double> d_log_tab;
    // reduced-mode variant of the stream (rows = recycled scratch slots), built on first use
    DevBuf<double> d_scratch;
    bool red_ready = false;
    // options
    uint32_t stream_key = 0;
    int warps_per_group = 0, groups_per_cta = 0;
    int samples_per_lane = 0;  // 0 auto, 2 pair kernel (mcdp_chunk_sweep.cuh), 4 quad kernel (mcdp_quad_sweep.cuh)
    int cluster_size = 0;      // 0 auto, 1 never, 2 / 4 / 8: CTAs per cluster that share one sample group (quad kernel)
    int64_t host_chunk = 0;
    int rng_stream = 0;  // 0 Philox contract, 1 reference-compatible Xoshiro stream
    DevBuf<ActRec> d_acts;
    DevBuf<double> d_norm_cache;
    // calls of a handful of samples (mcdp_small_sweep.cuh): the evaluation-ordered records themselves and one sampling
    // record per activity index, uploaded on first use
    struct SmallDev {
        DevBuf<EventRec> events;
        DevBuf<PredRec> items;
        DevBuf<uint32_t> pred_src, pred_act;
        DevBuf<int4> tiles;
        int32_t n_items = 0, n_tiles = 0;
        cudaMemPool_t pool = nullptr;  // stream-ordered scratch of the calls ([samples][precedence entries] durations)
        bool ready = false, unsupported = false;
    } small;
    int64_t small_max = -1;  // calls of at most this many samples take that path; -1 auto, 0 never
    // host-call workspaces
    HostSlot slots[2];
    DevBuf<double> d_stat_f64;
    DevBuf<unsigned long long> d_stat_u64;
    DevBuf<uint32_t> d_stat_u32;
    std::mutex stream_mu;  // lazy construction of the chunk streams
    std::mutex mu;  // instances are not re-entrant in the reference either; serialise instead of corrupting scratch
    // largest dynamic shared memory size each kernel has been opted in for on this plan's device
    std::mutex attr_mu;
    std::unordered_map<const void*, size_t> smem_attr;
    bool holds_l2_carveout = false;

    ~mcdp_plan();
};

namespace {

// The persisting-L2 carve-out is a device-wide limit: plans on one device share it.  It only ever grows while plans
// are alive (the largest record stream decides) and the device's original limit comes back with the last plan.
struct L2Carveout {
    std::mutex mu;
    struct PerDevice {
        int plans = 0;
        size_t original = 0, current = 0;
    };
    std::unordered_map<int, PerDevice> dev;
    size_t acquire(int device, size_t want) {
        std::lock_guard<std::mutex> lock(mu);
        PerDevice& d = dev[device];
        if (d.plans == 0) {
            if (cudaDeviceGetLimit(&d.original, cudaLimitPersistingL2CacheSize) != cudaSuccess) {
                cudaGetLastError();
                d.original = 0;
            }
            d.current = d.original;
        }
        ++d.plans;
        if (want > d.current) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
                d.current = want;
            else
                cudaGetLastError();
        }
        return d.current;
    }
    void release(int device) {
        std::lock_guard<std::mutex> lock(mu);
        auto it = dev.find(device);
        if (it == dev.end() || it->second.plans == 0) return;
        if (--it->second.plans == 0) {
            cudaCtxResetPersistingL2Cache();
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, it->second.original);
            cudaGetLastError();
            dev.erase(it);
        }
    }
};
L2Carveout g_l2;

}  // namespace

mcdp_plan::~mcdp_plan() {
    {
        DeviceGuard g(device);
        if (holds_l2_carveout) g_l2.release(device);
        for (auto& s : slots) s.release();
        for (auto& a : streams)
            for (auto& st : a) {
                st.units.release();
                st.level_begin.release();
            }
        d_orphans.release();
        d_dists.release();
        d_tab.release();
        d_log_tab.release();
        d_scratch.release();
        d_acts.release();
        d_norm_cache.release();
        small.events.release();
        small.items.release();
        small.pred_src.release();
        small.pred_act.release();
        small.tiles.release();
        if (small.pool) cudaMemPoolDestroy(small.pool);
        d_stat_f64.release();
        d_stat_u64.release();
        d_stat_u32.release();
    }
}

namespace {

template <typename T>
int32_t upload(DevBuf<T>& buf, const std::vector<T>& v) {
    MCDP_CUDA(buf.ensure(v.size()));
    if (!v.empty()) MCDP_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return MCDP_OK;
}

struct LaunchShape {
    int wpg, gpc, threads, batches;
    unsigned grid;
    size_t smem;
    bool smem_tables;
    int spl;  // samples per lane: 2 = pair kernel (64-sample groups), 4 = quad kernel (128-sample groups)
    int cluster = 1;  // quad kernel, one group per CTA: CTAs per thread-block cluster splitting the group's levels
    // what the shape was chosen for (launch_sweep re-chooses with spl = 2 when the buffers are not 32-byte aligned)
    bool reduced, single_batch;
    int n_bins;
};

// How many warps split a level and how many sample groups share a CTA for the kernel with `spl` samples per lane;
// `eff_out` = share of the SMs' warp slots the launch keeps busy over whole waves.
LaunchShape choose_shape_spl(const mcdp_plan* plan, int64_t n, bool reduced, int n_bins, bool single_batch, int spl,
                             double* eff_out) {
    const HostPlan& h = plan->host;
    LaunchShape s{};
    s.reduced = reduced;
    s.single_batch = single_batch;
    s.n_bins = n_bins;
    int64_t n_groups = (n + 63) / 64;
    const int batches = 1;  // every reduced launch is single-batch since the quad kernel stages its statistics
    s.spl = spl;
    const int sm_warps = spl == 4 ? MCDP_QUAD_MAX_THREADS / 32 : 32;  // resident warps per SM: 96 vs 64 registers per thread
    if (spl == 4) n_groups = (n + kQuadSamples - 1) / kQuadSamples;
    int wpg = plan->warps_per_group;
    const int kMaxWarps = (spl == 4 ? MCDP_QUAD_MAX_THREADS : MCDP_MAX_THREADS) / 32;
    int gpc = plan->groups_per_cta;
    const int gpc_warps = spl == 4 ? 10 : 8;  // default CTA: about this many warps (two CTAs per SM / four)
    if (wpg <= 0) {
        // Candidates: power-of-two warps per group, bounded by what the levels can feed (>= 4 events
        // per warp per level on average; 8 in the quad kernel, whose single CTA per SM has no second
        // group to run while one waits at a level barrier).  Pick the shape that keeps the most warp
        // slots busy over whole waves; ties go to more warps per group: fewer resident samples per
        // SM shorten the reuse distance of realized rows in L2.
        const int64_t avg_width = h.n_levels > 0 ? h.E / h.n_levels : 1;
        const int64_t by_width = std::max<int64_t>(1, avg_width / (spl == 4 ? 8 : 4));
        double best = -1.0;
        // pair kernel: powers of two (32 warp slots per SM); quad kernel: also the divisors of its 20 slots
        static const int kCandPair[] = {1, 2, 4, 8, 16}, kCandQuad[] = {1, 2, 4, 5, 8, 10, 16, 20};
        const int* cands = spl == 4 ? kCandQuad : kCandPair;
        const int n_cands = spl == 4 ? 8 : 5;
        for (int ci = 0; ci < n_cands; ++ci) {
            const int cand = cands[ci];
            if (cand > kMaxWarps || cand > by_width) break;
            const int g = gpc > 0 ? std::min(gpc, kMaxWarps / cand) : std::max(1, gpc_warps / cand);
            const int64_t ctas = (n_groups + g - 1) / g;
            const int64_t per_sm = std::max(1, sm_warps / (cand * g));
            const int64_t slots = int64_t(plan->sm_count) * per_sm;
            const int64_t waves = (ctas + slots - 1) / slots;
            const double eff = double(n_groups * cand) / double(waves * plan->sm_count * sm_warps);
            if (eff >= best - 1e-9) {
                best = eff;
                wpg = cand;
            }
        }
        wpg = std::max(wpg, 1);
    }
    wpg = std::max(1, std::min(wpg, kMaxWarps));
    if (gpc <= 0) gpc = std::max(1, gpc_warps / wpg);
    gpc = std::max(1, std::min({gpc, 15, kMaxWarps / wpg}));
    s.wpg = wpg;
    s.gpc = gpc;
    s.batches = batches;
    s.threads = 32 * wpg * gpc;
    s.grid = unsigned((n_groups + gpc - 1) / gpc);
    // Small launches of the quad kernel: fewer groups than SMs.  Spread each group over a thread-block cluster whose
    // CTAs split the levels (mcdp_quad_sweep.cuh), as far as the levels can feed the warps (about one chunk per warp
    // per level) and the machine has SMs to give.
    s.cluster = 1;
    if (spl == 4 && wpg > 1 && gpc == 1 && plan->cluster_size != 1 && n_groups > 0) {
        int c = 1;
        if (plan->cluster_size > 1) {
            c = plan->cluster_size;
        } else {
            const int64_t chunks_per_level = h.n_levels > 0 ? int64_t(h.units.size() / size_t(kChunkUnits)) / h.n_levels : 0;
            // (clusters of 8 CTAs of this size pack 16 to the machine, not 18: stay clear of a second wave)
            while (c < 8 && int64_t(2 * c) * n_groups * 4 <= int64_t(plan->sm_count) * 3 && int64_t(2 * c) * wpg <= chunks_per_level)
                c *= 2;
        }
        s.cluster = c;
        s.grid *= unsigned(c);
    }
    if (eff_out) {
        const int64_t per_sm = std::max(1, sm_warps / (wpg * gpc));
        const int64_t slots = int64_t(plan->sm_count) * per_sm;
        const int64_t waves = std::max<int64_t>(1, (int64_t(s.grid) + slots - 1) / slots);
        *eff_out = double(n_groups * wpg * s.cluster) / double(waves * plan->sm_count * sm_warps);
    }
    const size_t need = sizeof(DistRec) * h.dists.size() + sizeof(double) * h.tab_pool.size();
    // keep several CTAs per SM resident: stage only when the tables are a modest share of shared memory
    s.smem_tables = need > 0 && need <= std::min<size_t>(plan->smem_optin, 64 * 1024);
    s.smem = size_t(kLogTabBytes) + (s.smem_tables ? need : 0);  // the log table of mcdp_math.cuh is always staged
    return s;
}

// Which kernel "auto" means.  On B200 the quad kernel is 10-15 % faster on every named workload once its launch
// fills the machine (profiles/r01_quad_ab.txt), but its groups are twice as large: a launch too small to give every
// SM a quad group runs faster as pair groups spread over twice as many SMs (C3 at 9 472 samples: 21.8 vs 28.4 ms).
constexpr int kAutoSamplesPerLane = MCDP_AUTO_SPL;
constexpr double kQuadMinFill = 0.8;

LaunchShape choose_shape(const mcdp_plan* plan, int64_t n, bool reduced = false, int n_bins = 0, bool single_batch = false,
                         int force_spl = 0) {
    const int want = force_spl ? force_spl : plan->samples_per_lane;
    if (want) return choose_shape_spl(plan, n, reduced, n_bins, single_batch, want, nullptr);
    double e2 = 0.0, e4 = 0.0;
    const LaunchShape s2 = choose_shape_spl(plan, n, reduced, n_bins, single_batch, 2, &e2);
    if (kAutoSamplesPerLane != 4) return s2;
    const LaunchShape s4 = choose_shape_spl(plan, n, reduced, n_bins, single_batch, 4, &e4);
    return (s4.spl == 4 && (e4 >= kQuadMinFill || e4 > e2 + 0.1 || (s4.cluster > 1 && e4 > e2))) ? s4 : s2;
}

template <typename K>
int32_t launch_kernel(mcdp_plan* plan, K k, const SweepParams& p, const LaunchShape& s, size_t smem, const void* stream_base,
                      size_t stream_bytes, cudaStream_t stream) {
    if (smem > 48 * 1024) {
        // opt in once per kernel and size (not on every launch)
        std::lock_guard<std::mutex> lock(plan->attr_mu);
        size_t& have = plan->smem_attr[reinterpret_cast<const void*>(k)];
        if (smem > have) {
            MCDP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            have = smem;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(s.grid);
    cfg.blockDim = dim3(unsigned(s.threads));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n_attr = 0;
    if (s.cluster > 1) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = unsigned(s.cluster);
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    if (plan->l2_persist_bytes > 0 && plan->l2_window_max > 0 && stream_bytes > 0) {
        // the plan stream is re-read by every group while >100 GB of outputs stream through L2: pin it with a
        // persisting access-policy window
        const size_t win = std::min(stream_bytes, plan->l2_window_max);
        attr[n_attr].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[n_attr].val.accessPolicyWindow.base_ptr = const_cast<void*>(stream_base);
        attr[n_attr].val.accessPolicyWindow.num_bytes = win;
        attr[n_attr].val.accessPolicyWindow.hitRatio = float(std::min(1.0, double(plan->l2_persist_bytes) / double(win)));
        attr[n_attr].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[n_attr].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    MCDP_CUDA(cudaLaunchKernelEx(&cfg, k, p));
    return MCDP_OK;
}

int32_t ensure_chunk_stream(mcdp_plan* plan, bool reduced, bool dense, SweepParams& p);
int64_t reduced_chunk(const mcdp_plan* plan, int64_t n, int n_bins, bool attr);

// the launch-shape dependent fields of the parameter block
void apply_shape(SweepParams& p, const LaunchShape& s) {
    p.smem_ring_off = uint32_t((s.smem + 127) & ~size_t(127));
    p.smem_stat_off = p.smem_ring_off + uint32_t(chunk_ring_bytes(s.threads / 32));
    p.warps_per_group = s.wpg;
    p.batches_per_group = s.batches;
    p.cluster_size = s.cluster;
}

bool aligned32(const void* a) { return (reinterpret_cast<uintptr_t>(a) & 31) == 0; }

template <int MODE>
int32_t launch_sweep(mcdp_plan* plan, const SweepParams& p_in, const LaunchShape& s_in, cudaStream_t stream) {
    if (p_in.n <= 0) return MCDP_OK;
    SweepParams p = p_in;
    LaunchShape s = s_in;
    if (s.spl == 4 && !(aligned32(p.realized) && aligned32(p.durations) && aligned32(p.inj))) {
        // the quad kernel moves 256-bit row segments: caller buffers that are only 16-byte aligned take the pair kernel
        s = choose_shape(plan, p.n, s.reduced, s.n_bins, s.single_batch, 2);
        apply_shape(p, s);
    }
    {
        // chunk stream (mcdp_chunk_sweep.cuh): tables (128-byte rounded) + per-warp chunk ring.  One warp per
        // group walks the dense stream, several warps per group split the level-aligned one.
        const int32_t rc = ensure_chunk_stream(plan, MODE == kModeReduced || MODE == kModeAttr, s.wpg == 1, p);
        if (rc) return rc;
        const size_t bytes = size_t(p.n_chunks) * size_t(kChunkBytes);
        size_t smem = ((s.smem + 127) & ~size_t(127)) + chunk_ring_bytes(s.threads / 32);
        if (s.spl == 4 && (MODE == kModeReduced || MODE == kModeAttr)) smem += quad_stat_bytes(s.threads / 32);
        if (s.spl == 4) {
            if (s.wpg > 1)
                return s.smem_tables ? launch_kernel(plan, quad_sweep_kernel<MODE, true, true>, p, s, smem, p.chunks, bytes, stream)
                                     : launch_kernel(plan, quad_sweep_kernel<MODE, false, true>, p, s, smem, p.chunks, bytes, stream);
            return s.smem_tables ? launch_kernel(plan, quad_sweep_kernel<MODE, true, false>, p, s, smem, p.chunks, bytes, stream)
                                 : launch_kernel(plan, quad_sweep_kernel<MODE, false, false>, p, s, smem, p.chunks, bytes, stream);
        }
        if (s.wpg > 1)
            return s.smem_tables ? launch_kernel(plan, chunk_sweep_kernel<MODE, true, true>, p, s, smem, p.chunks, bytes, stream)
                                 : launch_kernel(plan, chunk_sweep_kernel<MODE, false, true>, p, s, smem, p.chunks, bytes, stream);
        return s.smem_tables ? launch_kernel(plan, chunk_sweep_kernel<MODE, true, false>, p, s, smem, p.chunks, bytes, stream)
                             : launch_kernel(plan, chunk_sweep_kernel<MODE, false, false>, p, s, smem, p.chunks, bytes, stream);
    }
}

SweepParams base_params(const mcdp_plan* plan, const LaunchShape& s, int64_t n, int64_t ld) {
    SweepParams p{};
    const HostPlan& h = plan->host;
    p.orphans = plan->d_orphans.p;
    p.dists = plan->d_dists.p;
    p.tab_pool = plan->d_tab.p;
    p.log_tab = plan->d_log_tab.p;
    p.n = n;
    p.ld = ld;
    p.ldb8 = uint32_t(ld) * 8u;
    p.ldb4 = uint32_t(ld) * 4u;
    p.smem_tab_off = uint32_t(kLogTabBytes + sizeof(DistRec) * h.dists.size());
    p.n_levels = h.n_levels;
    p.n_orphans = int32_t(h.orphans.size());
    p.n_dists = int32_t(h.dists.size());
    p.tab_pool_len = int32_t(h.tab_pool.size());
    p.E = h.E;
    p.max_delay = h.max_delay;
    for (int r = 0; r < 10; ++r) p.keys.k[r] = plan->stream_key + uint32_t(r) * 0x9E3779B9u;
    apply_shape(p, s);
    return p;
}

int32_t check_layout(const void* a, int64_t n, int64_t ld) {
    if (n < 0) return fail(MCDP_ERR_ARG, "n must be non-negative");
    if (ld < n || (ld & 63) || ld >= (int64_t(1) << 29))
        return fail(MCDP_ERR_ARG, "ld must be a multiple of 64, >= n and < 2^29");
    if (reinterpret_cast<uintptr_t>(a) & 15) return fail(MCDP_ERR_ARG, "device buffers must be 16-byte aligned");
    return MCDP_OK;
}

template <typename T>
int32_t launch_transpose(const T* in, int64_t in_ld, int64_t rows, int64_t cols, T* out, int64_t out_ld,
                         cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return MCDP_OK;
    const int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
    if (tiles_c * tiles_r > int64_t(0x7FFFFFFF)) return fail(MCDP_ERR_ARG, "transpose too large");
    transpose_kernel<T><<<unsigned(tiles_c * tiles_r), 256, 0, stream>>>(in, in_ld, rows, cols, out, out_ld, tiles_c);
    MCDP_CUDA(cudaGetLastError());
    return MCDP_OK;
}

int32_t ensure_reduced_stream(mcdp_plan* plan) {
    if (plan->red_ready) return MCDP_OK;
    const HostPlan& h = plan->host;
    std::vector<EventRec>& ev = plan->ev_red;
    std::vector<PredRec>& pr = plan->pr_red;
    ev = h.events;
    pr = h.preds;
    for (auto& q : pr) q.src_row = h.slot_of_event[q.src_row];  // host stream rows are event ids
    for (auto& e : ev) {
        e.row = h.slot_of_event[e.event];
        if (e.fan_in) e.first_src_row = pr[e.pred_begin].src_row;
        for (uint32_t k = 0; k < e.fan_in; ++k)
            pr[e.pred_begin + k].next_src_row = k + 1 < e.fan_in ? pr[e.pred_begin + k + 1].src_row : 0u;
    }
    plan->red_ready = true;
    return MCDP_OK;
}

// the chunk stream a launch needs: built and uploaded on first use
int32_t ensure_chunk_stream(mcdp_plan* plan, bool reduced, bool dense, SweepParams& p) {
    std::lock_guard<std::mutex> lock(plan->stream_mu);
    mcdp_plan::ChunkStreamDev& st = plan->streams[reduced ? 1 : 0][dense ? 1 : 0];
    if (!st.ready) {
        const HostPlan& h = plan->host;
        if (reduced) {
            int32_t rc = ensure_reduced_stream(plan);
            if (rc) return rc;
        }
        std::vector<ChunkUnit> units;
        std::vector<int32_t> clb;
        const std::vector<ChunkUnit>* src = &units;
        const std::vector<int32_t>* src_clb = &clb;
        if (!reduced && !dense) {
            src = &h.units;
            src_clb = &h.chunk_level_begin;
        } else {
            build_chunk_stream(reduced ? plan->ev_red : h.events, reduced ? plan->pr_red : h.preds, h.level_begin, h.n_levels,
                               dense, units, clb);
        }
        int32_t rc = upload(st.units, *src);
        if (!rc) rc = upload(st.level_begin, *src_clb);
        if (rc) return rc;
        st.n_chunks = int32_t(src->size() / size_t(kChunkUnits));
        st.ready = true;
    }
    p.chunks = st.units.p;
    p.chunk_level_begin = st.level_begin.p;
    p.n_chunks = st.n_chunks;
    return MCDP_OK;
}

int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- calls of a handful of samples (mcdp_small_sweep.cuh) ----
// Measured (profiles/r02_small_calls.txt): one CTA per sample, so up to a sample per SM the time grows slowly with the
// number of samples -- 100k-event DAG 0.66 ms for one sample, 1.3 ms for 64, against 4.4 ms of the quad kernel's
// cluster launch for anything up to 128; 1M-event DAG 5.8 ms against 35 ms; 50k-event chain 46 against 86 ms.  The
// sweep kernels' throughput wins from 100-250 samples on, the rule stays well below that.
constexpr int64_t kSmallAutoMax = 64;
constexpr int64_t kSmallMidPreds = int64_t(1) << 20;  // precedence entries up to which 128 samples still take that path
constexpr int64_t kSmallScratchMax = int64_t(1) << 30;  // bytes of [samples][precedence entries] durations per call
bool use_small(const mcdp_plan* plan, int64_t n) {
    if (n <= 0 || n > 65535 || plan->small.unsupported || plan->host.max_fan_in > kSmallTilePreds) return false;
    if ((plan->host.P + plan->host.E) * n * 8 > kSmallScratchMax && plan->small_max < 0) return false;
    if (plan->small_max >= 0) return n <= plan->small_max;
    // an explicit kernel choice (samples per lane, cluster size, warps per group) is a request for the sweep kernels
    if (plan->samples_per_lane || plan->cluster_size || plan->warps_per_group || plan->groups_per_cta) return false;
    // up to 64 samples on any DAG; up to 128 (still one CTA per SM) where a sample's pass is short
    const int64_t limit = plan->host.P <= kSmallMidPreds ? 2 * kSmallAutoMax : kSmallAutoMax;
    return n <= std::min<int64_t>(limit, plan->sm_count);
}

int32_t ensure_small(mcdp_plan* plan) {
    std::lock_guard<std::mutex> lock(plan->stream_mu);
    if (plan->small.ready) return MCDP_OK;
    const HostPlan& h = plan->host;
    // sampling records: every precedence entry (with its position) and every orphan activity, by sampler class --
    // the gamma samplers vote across the warp, so Marsaglia-Tsang and the exact transformation get warps of their own
    std::vector<PredRec> by_class[3];
    auto add = [&](PredRec r, uint32_t pos) {
        r.next_src_row = pos;
        int cls = 0;
        if ((r.meta >> 29) == uint32_t(MCDP_DIST_GAMMA)) cls = (h.dists[r.dist].flags & 8) ? 2 : 1;
        by_class[cls].push_back(r);
    };
    std::vector<uint32_t> src(h.preds.size()), act(h.preds.size());
    for (size_t j = 0; j < h.preds.size(); ++j) {
        add(h.preds[j], uint32_t(j));
        src[j] = h.preds[j].src_row;
        act[j] = h.preds[j].act;
    }
    for (const PredRec& r : h.orphans) add(r, kSmallNoPos);
    std::vector<PredRec> items;
    for (auto& v : by_class) {
        while (v.size() % 32) v.push_back(v.back());  // whole warps per class; the copy writes the same values again
        items.insert(items.end(), v.begin(), v.end());
    }
    // tiles: events of one level, bounded in events and in precedence entries
    std::vector<int4> tiles;
    for (int32_t l = 0; l < h.n_levels; ++l) {
        int32_t pos = h.level_begin[l];
        const int32_t end = h.level_begin[l + 1];
        while (pos < end) {
            int4 t = make_int4(pos, 0, int(h.events[size_t(pos)].pred_begin), 0);
            while (pos < end && t.y < kSmallTileEvents && t.w + int(h.events[size_t(pos)].fan_in) <= kSmallTilePreds) {
                t.w += int(h.events[size_t(pos)].fan_in);
                ++t.y;
                ++pos;
            }
            if (t.y == 0) {  // one event with more entries than a tile holds: this plan stays with the sweep kernels
                plan->small.unsupported = true;
                return MCDP_OK;
            }
            tiles.push_back(t);
        }
    }
    int32_t rc = upload(plan->small.items, items);
    if (!rc) rc = upload(plan->small.events, h.events);
    if (!rc) rc = upload(plan->small.pred_src, src);
    if (!rc) rc = upload(plan->small.pred_act, act);
    if (!rc) rc = upload(plan->small.tiles, tiles);
    if (rc) return rc;
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = plan->device;
    MCDP_CUDA(cudaMemPoolCreate(&plan->small.pool, &props));
    uint64_t keep = UINT64_MAX;  // freed scratch stays with the pool: a loop of run(seed) calls allocates once
    MCDP_CUDA(cudaMemPoolSetAttribute(plan->small.pool, cudaMemPoolAttrReleaseThreshold, &keep));
    MCDP_CUDA(cudaFuncSetAttribute(small_propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallStages * kSmallTilePreds * 12));
    plan->small.n_items = int32_t(items.size());
    plan->small.n_tiles = int32_t(tiles.size());
    plan->small.ready = true;
    return MCDP_OK;
}

// MODE kModeFull: sample every activity, then propagate; kModeInjected: propagate p.inj
template <int MODE>
int32_t launch_small(mcdp_plan* plan, const SweepParams& p, cudaStream_t stream) {
    if (p.n <= 0) return MCDP_OK;
    const HostPlan& h = plan->host;
    SmallParams sp{};
    sp.items = plan->small.items.p;
    sp.n_items = plan->small.n_items;
    sp.events = plan->small.events.p;
    sp.pred_src = plan->small.pred_src.p;
    sp.pred_act = plan->small.pred_act.p;
    sp.tiles = plan->small.tiles.p;
    sp.n_tiles = plan->small.n_tiles;
    sp.P = h.P;
    sp.dists = p.dists;
    sp.tab_pool = p.tab_pool;
    sp.log_tab = p.log_tab;
    sp.seeds = p.seeds;
    sp.seed0 = p.seed0;
    sp.n = p.n;
    sp.ld = p.ld;
    sp.realized = p.realized;
    sp.durations = p.durations;
    sp.durations_in = p.inj;
    sp.cause = p.cause;
    sp.max_delay = p.max_delay;
    sp.keys = p.keys;
    void* scratch = nullptr;
    MCDP_CUDA(cudaMallocFromPoolAsync(&scratch, size_t(std::max<int64_t>(h.P + h.E, 1)) * size_t(p.n) * 8, plan->small.pool, stream));
    sp.dur_by_pred = static_cast<double*>(scratch);
    sp.realized_by_sample = sp.dur_by_pred + h.P * p.n;
    sp.E = h.E;
    cudaError_t e = cudaSuccess;
    if (MODE == kModeFull) {
        if (sp.n_items > 0) {
            NvtxRange nvtx("mcdp:small sample");
            const dim3 grid(unsigned((sp.n_items + kSmallSampleThreads - 1) / kSmallSampleThreads), unsigned(p.n));
            small_sample_kernel<<<grid, kSmallSampleThreads, 0, stream>>>(sp);
            e = cudaGetLastError();
        }
    } else if (h.P > 0) {
        NvtxRange nvtx("mcdp:small gather");
        const dim3 grid(unsigned((h.P + kSmallSampleThreads - 1) / kSmallSampleThreads), unsigned(p.n));
        small_gather_kernel<<<grid, kSmallSampleThreads, 0, stream>>>(sp);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && h.E > 0) {
        NvtxRange nvtx("mcdp:small propagate");
        small_propagate_kernel<<<unsigned(p.n), kSmallSweepThreads, kSmallStages * kSmallTilePreds * 12, stream>>>(sp);
        e = cudaGetLastError();
    }
    const cudaError_t ef = cudaFreeAsync(scratch, stream);  // stream-ordered: after the kernels above
    if (e == cudaSuccess) e = ef;
    MCDP_CUDA(e);
    return MCDP_OK;
}

// the sweep of one launch: the small-call path or the sweep kernels
template <int MODE>
int32_t launch_any(mcdp_plan* plan, const SweepParams& p, const LaunchShape& s, cudaStream_t stream) {
    if (use_small(plan, p.n)) {
        const int32_t rc = ensure_small(plan);
        if (rc) return rc;
        if (!plan->small.unsupported) return launch_small<MODE>(plan, p, stream);
    }
    return launch_sweep<MODE>(plan, p, s, stream);
}

// reference-compatible stream: draw durations[A][ld] with the Xoshiro sampler; the caller then
// sweeps them in duration-injection mode
int32_t launch_compat_sampler(mcdp_plan* plan, const int32_t* d_seeds, int32_t seed0, int64_t n, double* d_durations,
                              int64_t ld, cudaStream_t stream) {
    const HostPlan& h = plan->host;
    if (h.A == 0 || n <= 0) return MCDP_OK;
    if (plan->d_acts.cap < size_t(h.A)) {
        std::vector<ActRec> acts(static_cast<size_t>(h.A));
        for (int32_t a = 0; a < h.A; ++a) acts[a] = ActRec{h.act_base[a], h.act_dist[a], 0u};
        int32_t rc = upload(plan->d_acts, acts);
        if (rc) return rc;
    }
    MCDP_CUDA(plan->d_norm_cache.ensure(std::max<size_t>(h.dists.size(), 1) * size_t(ld)));
    CompatParams cp{};
    cp.acts = plan->d_acts.p;
    cp.dists = plan->d_dists.p;
    cp.tab_pool = plan->d_tab.p;
    cp.seeds = d_seeds;
    cp.durations = d_durations;
    cp.norm_cache = plan->d_norm_cache.p;
    cp.n = n;
    cp.ld = ld;
    cp.A = h.A;
    cp.n_dists = int32_t(h.dists.size());
    cp.seed0 = seed0;
    const int64_t cols = round_up(n, 64);
    compat_sample_kernel<<<unsigned((cols + 127) / 128), 128, 0, stream>>>(cp);
    MCDP_CUDA(cudaGetLastError());
    return MCDP_OK;
}

}  // namespace

extern "C" {

const char* mcdp_last_error(void) { return g_err.c_str(); }
int32_t mcdp_abi_version(void) { return MCDP_ABI_VERSION; }

int32_t mcdp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

}  // extern "C"

namespace {
// a plan on `device` (or host-only) around an already compiled host plan
int32_t plan_from_host(std::shared_ptr<HostPlan> host, int32_t device, mcdp_plan** out) {
    *out = nullptr;
    std::unique_ptr<mcdp_plan> plan(new (std::nothrow) mcdp_plan(std::move(host)));
    if (!plan) return fail(MCDP_ERR_ARG, "out of host memory");
    if (device == MCDP_DEVICE_NONE) {
        // validation / introspection only: every run call on this plan fails (there is no CPU path)
        plan->device = MCDP_DEVICE_NONE;
        *out = plan.release();
        return MCDP_OK;
    }
    // No CPU fallback: a runnable plan lives on a CUDA device or does not exist.
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev <= 0) {
        cudaGetLastError();
        plan->device = MCDP_DEVICE_NONE;
        return fail(MCDP_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(ce));
    }
    if (device < 0 || device >= n_dev) {
        plan->device = MCDP_DEVICE_NONE;
        return fail(MCDP_ERR_ARG, "device ordinal out of range");
    }
    plan->device = device;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MCDP_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        plan->sm_count = prop.multiProcessorCount;
        plan->smem_optin = prop.sharedMemPerBlockOptin;
        plan->l2_window_max = size_t(std::max(prop.accessPolicyMaxWindowSize, 0));
        // set aside L2 for the record stream (a device-wide limit shared by the plans of this device: see L2Carveout)
        const size_t stream_bytes = plan->host.units.size() * sizeof(ChunkUnit);
        const size_t want = std::min<size_t>({stream_bytes + (stream_bytes >> 3) + (1u << 20),
                                             size_t(std::max(prop.persistingL2CacheMaxSize, 0)), size_t(64) << 20});
        if (want > 0) {
            plan->l2_persist_bytes = std::min(g_l2.acquire(device, want), want);
            plan->holds_l2_carveout = true;
        }
    }
    const HostPlan& h = plan->host;
    NvtxRange nvtx_upload("mcdp:plan H2D");
    int32_t rc = upload(plan->d_orphans, h.orphans);
    if (!rc) rc = upload(plan->d_dists, h.dists);
    if (!rc) rc = upload(plan->d_tab, h.tab_pool);
    if (!rc) {
        std::vector<double> lt(size_t(2 * kLogTabEntries));
        make_log_table(lt.data());
        rc = upload(plan->d_log_tab, lt);
    }
    if (rc) return rc;
    *out = plan.release();
    return MCDP_OK;
}

int32_t compile_host_plan(const mcdp_graph_desc* graph, const mcdp_dists_desc* dists, std::shared_ptr<HostPlan>* out) {
    auto host = std::make_shared<HostPlan>();
    std::string err;
    bool ok = false;
    NvtxRange nvtx_compile("mcdp:plan compile");
    try {
        ok = compile_plan(*graph, *dists, *host, err);
    } catch (const std::exception& e) {
        err = e.what();
    }
    if (!ok) return fail(MCDP_ERR_INVALID, err);
    *out = std::move(host);
    return MCDP_OK;
}
}  // namespace

extern "C" {

int32_t mcdp_plan_create(const mcdp_graph_desc* graph, const mcdp_dists_desc* dists, int32_t device, mcdp_plan** out) {
    if (!graph || !dists || !out) return fail(MCDP_ERR_ARG, "null argument");
    *out = nullptr;
    std::shared_ptr<HostPlan> host;
    const int32_t rc = compile_host_plan(graph, dists, &host);
    if (rc) return rc;
    return plan_from_host(std::move(host), device, out);
}

void mcdp_plan_destroy(mcdp_plan* plan) { delete plan; }

int32_t mcdp_plan_set_option(mcdp_plan* plan, int32_t option, int64_t value) {
    if (!plan) return fail(MCDP_ERR_ARG, "null plan");
    switch (option) {
        case MCDP_OPT_STREAM_KEY: plan->stream_key = uint32_t(value); break;
        case MCDP_OPT_WARPS_PER_GROUP:
            if (value < 0 || value > std::max(MCDP_MAX_THREADS, MCDP_QUAD_MAX_THREADS) / 32)
                return fail(MCDP_ERR_ARG, "warps per group out of range (0 = auto, at most 16 for 512-thread CTAs)");
            plan->warps_per_group = int(value);
            break;
        case MCDP_OPT_GROUPS_PER_CTA:
            if (value < 0 || value > 15) return fail(MCDP_ERR_ARG, "groups per CTA must be 0..15");
            plan->groups_per_cta = int(value);
            break;
        case MCDP_OPT_RNG_STREAM:
            if (value != 0 && value != 1) return fail(MCDP_ERR_ARG, "rng stream must be 0 (Philox) or 1 (reference-compatible)");
            if (value == 1 && plan->host.dists.size() > 64)
                return fail(MCDP_ERR_ARG, "the reference-compatible stream supports at most 64 distributions");
            plan->rng_stream = int(value);
            break;
        case MCDP_OPT_SAMPLES_PER_LANE:
            if (value != 0 && value != 2 && value != 4) return fail(MCDP_ERR_ARG, "samples per lane must be 0 (auto), 2 or 4");
            plan->samples_per_lane = int(value);
            break;
        case MCDP_OPT_CLUSTER_SIZE:
            if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8)
                return fail(MCDP_ERR_ARG, "cluster size must be 0 (auto), 1, 2, 4 or 8");
            plan->cluster_size = int(value);
            break;
        case MCDP_OPT_SMALL_CALL_MAX:
            if (value < -1 || value > 65535) return fail(MCDP_ERR_ARG, "small-call limit must be -1 (auto), 0 (never) or at most 65535 samples");
            plan->small_max = value;
            break;
        case MCDP_OPT_HOST_CHUNK:
            if (value < 0) return fail(MCDP_ERR_ARG, "host chunk must be non-negative");
            plan->host_chunk = value;
            break;
        default: return fail(MCDP_ERR_ARG, "unknown option");
    }
    return MCDP_OK;
}

int32_t mcdp_plan_node_count(const mcdp_plan* plan) { return plan ? plan->host.E : 0; }
int32_t mcdp_plan_activity_count(const mcdp_plan* plan) { return plan ? plan->host.A : 0; }
int64_t mcdp_plan_pred_count(const mcdp_plan* plan) { return plan ? plan->host.P : 0; }
int32_t mcdp_plan_level_count(const mcdp_plan* plan) { return plan ? plan->host.n_levels : 0; }
int32_t mcdp_plan_slot_count(const mcdp_plan* plan) { return plan ? plan->host.n_slots : 0; }
int32_t mcdp_plan_device(const mcdp_plan* plan) { return plan ? plan->device : -1; }

int32_t mcdp_plan_get_order(const mcdp_plan* plan, int32_t* order_out, int32_t* level_out) {
    if (!plan) return fail(MCDP_ERR_ARG, "null plan");
    const HostPlan& h = plan->host;
    if (order_out) std::copy(h.order.begin(), h.order.end(), order_out);
    if (level_out) std::copy(h.level_of_pos.begin(), h.level_of_pos.end(), level_out);
    return MCDP_OK;
}

int64_t mcdp_plan_get_cumulative(const mcdp_plan* plan, int32_t activity_type, double* cp_out, int64_t cap) {
    if (!plan) return -1;
    const HostPlan& h = plan->host;
    for (size_t i = 0; i < h.dists.size(); ++i) {
        if (h.dist_types[i] != activity_type) continue;
        const DistRec& d = h.dists[i];
        if (d.tab_off < 0) return -1;
        const size_t cp0 = size_t(d.tab_off) + guide_doubles(uint32_t(d.guide_log2));
        for (int64_t k = 0; k < d.tab_len && k < cap; ++k) cp_out[k] = h.tab_pool[cp0 + size_t(k)];
        return d.tab_len;
    }
    return -1;
}

int64_t mcdp_plan_get_chunks(const mcdp_plan* plan, int32_t rows, int32_t dense, void* units_out, int64_t cap_bytes,
                             int32_t* chunk_level_begin_out) {
    if (!plan) return -1;
    const HostPlan& h = plan->host;
    std::vector<EventRec> ev = h.events;
    std::vector<PredRec> pr = h.preds;
    if (rows) {  // scratch-slot rows, as ensure_reduced_stream builds them
        for (auto& q : pr) q.src_row = h.slot_of_event[q.src_row];
        for (auto& e : ev) e.row = h.slot_of_event[e.event];
    }
    std::vector<ChunkUnit> units;
    std::vector<int32_t> clb;
    build_chunk_stream(ev, pr, h.level_begin, h.n_levels, dense != 0, units, clb);
    const int64_t bytes = int64_t(units.size() * sizeof(ChunkUnit));
    if (units_out && cap_bytes > 0) std::memcpy(units_out, units.data(), size_t(std::min(bytes, cap_bytes)));
    if (chunk_level_begin_out) std::copy(clb.begin(), clb.end(), chunk_level_begin_out);
    return int64_t(units.size() / size_t(kChunkUnits));
}

int64_t mcdp_plan_reduced_chunk(mcdp_plan* plan, int64_t n, int32_t n_bins, int32_t attribution) {
    if (!plan || n < 0) return -1;
    std::lock_guard<std::mutex> lock(plan->mu);
    DeviceGuard guard(plan->device);
    return reduced_chunk(plan, n, n_bins, attribution != 0);
}

int32_t mcdp_plan_launch_shape(const mcdp_plan* plan, int64_t n, int32_t reduced, int32_t n_bins, int64_t* out8) {
    if (!plan || !out8) return fail(MCDP_ERR_ARG, "null argument");
    if (n < 0) return fail(MCDP_ERR_ARG, "n must be non-negative");
    if (!reduced && use_small(plan, n)) {  // one thread per activity, then per event: one sample per "lane"
        const int64_t shape[8] = {1, 0, 0, kSmallSweepThreads, n, 0, 1, 0};
        std::copy(shape, shape + 8, out8);
        return MCDP_OK;
    }
    const LaunchShape s = choose_shape(plan, n, reduced != 0, n_bins);
    out8[0] = s.spl;
    out8[1] = s.wpg;
    out8[2] = s.gpc;
    out8[3] = s.threads;
    out8[4] = int64_t(s.grid);
    out8[5] = int64_t(((s.smem + 127) & ~size_t(127)) + chunk_ring_bytes(s.threads / 32) +
                      (s.spl == 4 && reduced ? quad_stat_bytes(s.threads / 32) : 0));
    out8[6] = s.cluster;
    out8[7] = s.smem_tables ? 1 : 0;
    return MCDP_OK;
}

int32_t mcdp_run_full_device(mcdp_plan* plan, const int32_t* d_seeds, int32_t seed0, int64_t n, double* d_realized,
                             double* d_durations, int32_t* d_cause, int64_t ld, void* stream) {
    if (!plan) return fail(MCDP_ERR_ARG, "null plan");
    if (plan->device < 0) return fail(MCDP_ERR_CUDA, "plan was created host-only (MCDP_DEVICE_NONE): there is no CPU execution path");
    if (((!d_realized || !d_cause) && plan->host.E > 0) || (!d_durations && plan->host.A > 0))
        return fail(MCDP_ERR_ARG, "null output buffer");
    int32_t rc = check_layout(d_realized, n, ld);
    if (!rc) rc = check_layout(d_durations, n, ld);
    if (!rc) rc = check_layout(d_cause, n, ld);
    if (rc) return rc;
    DeviceGuard guard(plan->device);
    const LaunchShape s = choose_shape(plan, n);
    SweepParams p = base_params(plan, s, n, ld);
    p.seeds = d_seeds;
    p.seed0 = seed0;
    p.realized = d_realized;
    p.durations = d_durations;
    p.cause = d_cause;
    if (plan->rng_stream == 1) {
        std::lock_guard<std::mutex> lock(plan->mu);
        rc = launch_compat_sampler(plan, d_seeds, seed0, n, d_durations, ld, static_cast<cudaStream_t>(stream));
        if (rc) return rc;
        p.inj = d_durations;
        return launch_any<kModeInjected>(plan, p, s, static_cast<cudaStream_t>(stream));
    }
    return launch_any<kModeFull>(plan, p, s, static_cast<cudaStream_t>(stream));
}

int32_t mcdp_run_injected_device(mcdp_plan* plan, const double* d_durations, int64_t n, double* d_realized,
                                 int32_t* d_cause, int64_t ld, void* stream) {
    if (!plan) return fail(MCDP_ERR_ARG, "null plan");
    if (plan->device < 0) return fail(MCDP_ERR_CUDA, "plan was created host-only (MCDP_DEVICE_NONE): there is no CPU execution path");
    if (((!d_realized || !d_cause) && plan->host.E > 0) || (!d_durations && plan->host.A > 0))
        return fail(MCDP_ERR_ARG, "null buffer");
    int32_t rc = check_layout(d_realized, n, ld);
    if (!rc) rc = check_layout(d_durations, n, ld);
    if (!rc) rc = check_layout(d_cause, n, ld);
    if (rc) return rc;
    DeviceGuard guard(plan->device);
    const LaunchShape s = choose_shape(plan, n);
    SweepParams p = base_params(plan, s, n, ld);
    p.realized = d_realized;
    p.inj = d_durations;
    p.cause = d_cause;
    return launch_any<kModeInjected>(plan, p, s, static_cast<cudaStream_t>(stream));
}

int32_t mcdp_run_reduced_device(mcdp_plan* plan, const int32_t* d_seeds, int32_t seed0, int64_t n,
                                const mcdp_stats_desc* desc, double* d_sum, double* d_sumsq,
                                unsigned long long* d_late, uint32_t* d_hist, void* stream) {
    return mcdp_run_attribution_device(plan, d_seeds, seed0, n, desc, d_sum, d_sumsq, d_late, d_hist, nullptr, nullptr, stream);
}

}  // extern "C"

namespace {
// Samples per launch of a reduced call over n samples.  The realized scratch uses recycled slot rows; the samples
// are processed in chunks that bound it: what is already allocated, else 60 % of the free HBM (deep DAGs keep many
// rows live).
int64_t reduced_chunk(const mcdp_plan* plan, int64_t n, int n_bins, bool attr) {
    const HostPlan& h = plan->host;
    const int64_t bytes_per_sample = int64_t(std::max(h.n_slots, 1)) * 8;
    size_t free_b = 0, total_b = 0;
    if (plan->device < 0 || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        free_b = size_t(160) << 30;
    }
    const int64_t budget = std::max<int64_t>(int64_t(plan->d_scratch.cap) * 8, int64_t(double(free_b) * 0.6));
    int64_t chunk = std::max<int64_t>(64, budget / bytes_per_sample / 64 * 64);
    chunk = std::min<int64_t>(chunk, int64_t(1) << 22);
    chunk = std::min<int64_t>(chunk, round_up(std::max<int64_t>(n, 1), 64));
    if (n > chunk) {
        // several launches: whole waves of sample groups each (a launch of 0.9 waves idles a tenth of the machine for
        // its whole duration), and equal sizes so the last launch is not a sliver
        // (the wave of the kernel the launches will use: the quad kernel unless the plan is pinned to the pair kernel --
        // asking the auto rule about the not yet aligned size would answer for a launch of 1.x waves)
        const LaunchShape s = choose_shape(plan, chunk, true, n_bins, attr, plan->samples_per_lane == 2 ? 2 : 4);
        const int64_t sm_warps = s.spl == 4 ? MCDP_QUAD_MAX_THREADS / 32 : 32;
        const int64_t ctas_per_sm = std::max<int64_t>(1, sm_warps / (int64_t(s.wpg) * s.gpc));
        const int64_t wave = int64_t(plan->sm_count) * ctas_per_sm * s.gpc * (s.spl == 4 ? kQuadSamples : 64);
        if (chunk >= wave) {
            chunk = chunk / wave * wave;
            const int64_t launches = (n + chunk - 1) / chunk;
            chunk = std::min(chunk, round_up((n + launches - 1) / launches, wave));
        }
    }
    return chunk;
}

// plan->mu must be held by the caller (the device entry point and the host entry point both take it for the whole call)
int32_t run_attribution_locked(mcdp_plan* plan, const int32_t* d_seeds, int32_t seed0, int64_t n,
                               const mcdp_stats_desc* desc, double* d_sum, double* d_sumsq,
                               unsigned long long* d_late, uint32_t* d_hist, unsigned long long* d_cause_act,
                               unsigned long long* d_cause_none, void* stream) {
    if (!plan || !desc) return fail(MCDP_ERR_ARG, "null argument");
    const bool attr = d_cause_act != nullptr || d_cause_none != nullptr;
    if (attr && (!d_cause_act || !d_cause_none)) return fail(MCDP_ERR_ARG, "cause_act and cause_none go together");
    if (plan->device < 0) return fail(MCDP_ERR_CUDA, "plan was created host-only (MCDP_DEVICE_NONE): there is no CPU execution path");
    if (n < 0) return fail(MCDP_ERR_ARG, "n must be non-negative");
    if (plan->rng_stream == 1) return fail(MCDP_ERR_ARG, "the reference-compatible stream is available for full-output calls only");
    if (desc->n_thresholds < 0 || desc->n_thresholds > MCDP_MAX_THRESHOLDS) return fail(MCDP_ERR_ARG, "n_thresholds must be 0..4");
    if (desc->n_bins < 0 || desc->n_bins > 1024) return fail(MCDP_ERR_ARG, "n_bins must be 0..1024");
    if (d_hist && desc->n_bins > 0 && !(desc->hist_hi > desc->hist_lo)) return fail(MCDP_ERR_ARG, "hist_hi must exceed hist_lo");
    DeviceGuard guard(plan->device);
    NvtxRange nvtx("mcdp:reduced sweep");
    int32_t rc = ensure_reduced_stream(plan);
    if (rc) return rc;
    const HostPlan& h = plan->host;
    const int64_t chunk = reduced_chunk(plan, n, d_hist ? desc->n_bins : 0, attr);
    MCDP_CUDA(plan->d_scratch.ensure(size_t(std::max(h.n_slots, 1)) * size_t(chunk)));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = std::min(chunk, n - off);
        const LaunchShape s = choose_shape(plan, m, true, d_hist ? desc->n_bins : 0, attr);
        SweepParams p = base_params(plan, s, m, chunk);
        p.cause_act = d_cause_act;
        p.cause_none = d_cause_none;
        p.seeds = d_seeds ? d_seeds + off : nullptr;
        p.seed0 = int32_t(uint32_t(seed0) + uint32_t(off));
        p.realized = plan->d_scratch.p;
        p.sum = d_sum;
        p.sumsq = d_sumsq;
        p.late = desc->n_thresholds > 0 ? d_late : nullptr;
        p.hist = desc->n_bins > 0 ? d_hist : nullptr;
        p.n_thresholds = desc->n_thresholds;
        for (int i = 0; i < desc->n_thresholds; ++i) p.thresholds[i] = desc->thresholds[i];
        p.n_bins = desc->n_bins;
        p.hist_lo = desc->hist_lo;
        p.hist_scale = desc->n_bins > 0 ? double(desc->n_bins) / (desc->hist_hi - desc->hist_lo) : 0.0;
        rc = attr ? launch_sweep<kModeAttr>(plan, p, s, st) : launch_sweep<kModeReduced>(plan, p, s, st);
        if (rc) return rc;
    }
    return MCDP_OK;
}
}  // namespace

extern "C" {

int32_t mcdp_run_attribution_device(mcdp_plan* plan, const int32_t* d_seeds, int32_t seed0, int64_t n,
                                    const mcdp_stats_desc* desc, double* d_sum, double* d_sumsq,
                                    unsigned long long* d_late, uint32_t* d_hist, unsigned long long* d_cause_act,
                                    unsigned long long* d_cause_none, void* stream) {
    if (!plan || !desc) return fail(MCDP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(plan->mu);
    return run_attribution_locked(plan, d_seeds, seed0, n, desc, d_sum, d_sumsq, d_late, d_hist, d_cause_act, d_cause_none,
                                  stream);
}

int32_t mcdp_transpose_f64_device(const double* d_in, int64_t rows, int64_t n, int64_t ld, double* d_out, void* stream) {
    return launch_transpose<double>(d_in, ld, rows, n, d_out, rows, static_cast<cudaStream_t>(stream));
}
int32_t mcdp_transpose_i32_device(const int32_t* d_in, int64_t rows, int64_t n, int64_t ld, int32_t* d_out, void* stream) {
    return launch_transpose<int32_t>(d_in, ld, rows, n, d_out, rows, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// host-buffer calls
// ---------------------------------------------------------------------------------------------

static int64_t pick_host_chunk(mcdp_plan* plan, int64_t n, bool with_durations) {
    const HostPlan& h = plan->host;
    if (plan->host_chunk > 0) return round_up(std::min(plan->host_chunk, std::max<int64_t>(n, 1)), 64);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = size_t(8) << 30;
    // per sample: event-major + sample-major staging copies, two slots in flight
    const int64_t per_sample = 2 * 2 * (int64_t(h.E) * 12 + (with_durations ? int64_t(h.A) * 8 : 0)) + 64;
    int64_t chunk = int64_t(double(free_b) * 0.5) / per_sample / 64 * 64;
    chunk = std::max<int64_t>(64, std::min<int64_t>(chunk, 1 << 16));
    return std::min(chunk, round_up(std::max<int64_t>(n, 1), 64));
}

int32_t mcdp_run_many_host(mcdp_plan* plan, const int32_t* seeds, int64_t n, double* realized, double* durations,
                           int32_t* cause) {
    if (!plan) return fail(MCDP_ERR_ARG, "null plan");
    if (plan->device < 0) return fail(MCDP_ERR_CUDA, "plan was created host-only (MCDP_DEVICE_NONE): there is no CPU execution path");
    if (n < 0) return fail(MCDP_ERR_ARG, "n must be non-negative");
    if (n > 0 && !seeds) return fail(MCDP_ERR_ARG, "null seeds");
    if (n == 0) return MCDP_OK;
    std::lock_guard<std::mutex> lock(plan->mu);
    DeviceGuard guard(plan->device);
    const HostPlan& h = plan->host;
    const int64_t E = h.E, A = h.A;
    const int64_t chunk = pick_host_chunk(plan, n, true);
    int32_t rc = MCDP_OK;
    int64_t idx = 0;
    for (int64_t off = 0; off < n && !rc; off += chunk, ++idx) {
        HostSlot& sl = plan->slots[idx & 1];
        const int64_t m = std::min(chunk, n - off);
        if (!sl.stream) MCDP_CUDA_BRK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        // a slot is reused every other chunk: its previous copies must have drained
        MCDP_CUDA_BRK(cudaStreamSynchronize(sl.stream));
        MCDP_CUDA_BRK(sl.seeds.ensure(size_t(chunk)));
        MCDP_CUDA_BRK(sl.realized.ensure(size_t(E) * size_t(chunk)));
        MCDP_CUDA_BRK(sl.durations.ensure(size_t(A) * size_t(chunk)));
        MCDP_CUDA_BRK(sl.cause.ensure(size_t(E) * size_t(chunk)));
        {
            NvtxRange nvtx("mcdp:seeds H2D");
            MCDP_CUDA_BRK(cudaMemcpyAsync(sl.seeds.p, seeds + off, size_t(m) * 4, cudaMemcpyHostToDevice, sl.stream));
        }
        NvtxRange nvtx_chunk("mcdp:chunk sweep + transpose + D2H");
        const LaunchShape s = choose_shape(plan, m);
        SweepParams p = base_params(plan, s, m, chunk);
        p.seeds = sl.seeds.p;
        p.realized = sl.realized.p;
        p.durations = sl.durations.p;
        p.cause = sl.cause.p;
        if (plan->rng_stream == 1) {
            // the sampler's normal cache is plan-wide scratch: chunks of the compatible stream run one after another
            cudaError_t ce = cudaSuccess;
            for (auto& other : plan->slots)
                if (other.stream && other.stream != sl.stream && ce == cudaSuccess) ce = cudaStreamSynchronize(other.stream);
            MCDP_CUDA_BRK(ce);
            rc = launch_compat_sampler(plan, sl.seeds.p, 0, m, sl.durations.p, chunk, sl.stream);
            if (rc) break;
            p.inj = sl.durations.p;
            rc = launch_any<kModeInjected>(plan, p, s, sl.stream);
        } else {
            rc = launch_any<kModeFull>(plan, p, s, sl.stream);
        }
        if (rc) break;
        if (realized && E) {
            MCDP_CUDA_BRK(sl.t_realized.ensure(size_t(E) * size_t(chunk)));
            rc = launch_transpose<double>(sl.realized.p, chunk, E, m, sl.t_realized.p, E, sl.stream);
            if (rc) break;
            MCDP_CUDA_BRK(cudaMemcpyAsync(realized + off * E, sl.t_realized.p, size_t(m) * E * 8, cudaMemcpyDeviceToHost, sl.stream));
        }
        if (durations && A) {
            MCDP_CUDA_BRK(sl.t_durations.ensure(size_t(A) * size_t(chunk)));
            rc = launch_transpose<double>(sl.durations.p, chunk, A, m, sl.t_durations.p, A, sl.stream);
            if (rc) break;
            MCDP_CUDA_BRK(cudaMemcpyAsync(durations + off * A, sl.t_durations.p, size_t(m) * A * 8, cudaMemcpyDeviceToHost, sl.stream));
        }
        if (cause && E) {
            MCDP_CUDA_BRK(sl.t_cause.ensure(size_t(E) * size_t(chunk)));
            rc = launch_transpose<int32_t>(sl.cause.p, chunk, E, m, sl.t_cause.p, E, sl.stream);
            if (rc) break;
            MCDP_CUDA_BRK(cudaMemcpyAsync(cause + off * E, sl.t_cause.p, size_t(m) * E * 4, cudaMemcpyDeviceToHost, sl.stream));
        }
    }
    for (auto& sl : plan->slots)
        if (sl.stream) {
            cudaError_t e = cudaStreamSynchronize(sl.stream);
            if (e != cudaSuccess && !rc) rc = fail(MCDP_ERR_CUDA, std::string("stream sync: ") + cudaGetErrorString(e));
        }
    return rc;
}

int32_t mcdp_run_injected_host(mcdp_plan* plan, const double* durations, int64_t n, double* realized, int32_t* cause) {
    if (!plan) return fail(MCDP_ERR_ARG, "null plan");
    if (plan->device < 0) return fail(MCDP_ERR_CUDA, "plan was created host-only (MCDP_DEVICE_NONE): there is no CPU execution path");
    if (n < 0) return fail(MCDP_ERR_ARG, "n must be non-negative");
    if (n == 0) return MCDP_OK;
    const HostPlan& h = plan->host;
    if (!durations && h.A > 0) return fail(MCDP_ERR_ARG, "null durations");
    std::lock_guard<std::mutex> lock(plan->mu);
    DeviceGuard guard(plan->device);
    const int64_t E = h.E, A = h.A;
    const int64_t chunk = pick_host_chunk(plan, n, true);
    int32_t rc = MCDP_OK;
    int64_t idx = 0;
    for (int64_t off = 0; off < n && !rc; off += chunk, ++idx) {
        HostSlot& sl = plan->slots[idx & 1];
        const int64_t m = std::min(chunk, n - off);
        if (!sl.stream) MCDP_CUDA_BRK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        MCDP_CUDA_BRK(cudaStreamSynchronize(sl.stream));
        MCDP_CUDA_BRK(sl.realized.ensure(size_t(E) * size_t(chunk)));
        MCDP_CUDA_BRK(sl.durations.ensure(size_t(A) * size_t(chunk)));
        MCDP_CUDA_BRK(sl.t_durations.ensure(size_t(A) * size_t(chunk)));
        MCDP_CUDA_BRK(sl.cause.ensure(size_t(E) * size_t(chunk)));
        if (A) {
            MCDP_CUDA_BRK(cudaMemcpyAsync(sl.t_durations.p, durations + off * A, size_t(m) * A * 8, cudaMemcpyHostToDevice, sl.stream));
            // [m][A] sample-major -> [A][chunk] event-major
            rc = launch_transpose<double>(sl.t_durations.p, A, m, A, sl.durations.p, chunk, sl.stream);
            if (rc) break;
        }
        const LaunchShape s = choose_shape(plan, m);
        SweepParams p = base_params(plan, s, m, chunk);
        p.realized = sl.realized.p;
        p.inj = sl.durations.p;
        p.cause = sl.cause.p;
        rc = launch_any<kModeInjected>(plan, p, s, sl.stream);
        if (rc) break;
        if (realized && E) {
            MCDP_CUDA_BRK(sl.t_realized.ensure(size_t(E) * size_t(chunk)));
            rc = launch_transpose<double>(sl.realized.p, chunk, E, m, sl.t_realized.p, E, sl.stream);
            if (rc) break;
            MCDP_CUDA_BRK(cudaMemcpyAsync(realized + off * E, sl.t_realized.p, size_t(m) * E * 8, cudaMemcpyDeviceToHost, sl.stream));
        }
        if (cause && E) {
            MCDP_CUDA_BRK(sl.t_cause.ensure(size_t(E) * size_t(chunk)));
            rc = launch_transpose<int32_t>(sl.cause.p, chunk, E, m, sl.t_cause.p, E, sl.stream);
            if (rc) break;
            MCDP_CUDA_BRK(cudaMemcpyAsync(cause + off * E, sl.t_cause.p, size_t(m) * E * 4, cudaMemcpyDeviceToHost, sl.stream));
        }
    }
    for (auto& sl : plan->slots)
        if (sl.stream) {
            cudaError_t e = cudaStreamSynchronize(sl.stream);
            if (e != cudaSuccess && !rc) rc = fail(MCDP_ERR_CUDA, std::string("stream sync: ") + cudaGetErrorString(e));
        }
    return rc;
}

int32_t mcdp_run_reduced_host(mcdp_plan* plan, const int32_t* seeds, int64_t n, const mcdp_stats_desc* desc,
                              double* sum, double* sumsq, unsigned long long* late, uint32_t* hist) {
    return mcdp_run_attribution_host(plan, seeds, n, desc, sum, sumsq, late, hist, nullptr, nullptr);
}

int32_t mcdp_run_attribution_host(mcdp_plan* plan, const int32_t* seeds, int64_t n, const mcdp_stats_desc* desc,
                                  double* sum, double* sumsq, unsigned long long* late, uint32_t* hist,
                                  unsigned long long* cause_act, unsigned long long* cause_none) {
    if (!plan || !desc) return fail(MCDP_ERR_ARG, "null argument");
    if (plan->device < 0) return fail(MCDP_ERR_CUDA, "plan was created host-only (MCDP_DEVICE_NONE): there is no CPU execution path");
    if (n > 0 && !seeds) return fail(MCDP_ERR_ARG, "null seeds");
    const bool attr = cause_act != nullptr || cause_none != nullptr;
    if (attr && (!cause_act || !cause_none)) return fail(MCDP_ERR_ARG, "cause_act and cause_none go together");
    const int64_t E = plan->host.E, A = plan->host.A;
    const int64_t nt = std::max(desc->n_thresholds, 0), nb = std::max(desc->n_bins, 0);
    // One call at a time per plan, from the sizing of the shared accumulators to the last copy-back: two threads on
    // one propagator would otherwise sum into each other's statistics (the binding releases the GIL here).
    std::lock_guard<std::mutex> lock(plan->mu);
    DeviceGuard guard(plan->device);
    HostSlot& sl = plan->slots[0];
    if (!sl.stream) MCDP_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    cudaStream_t st = sl.stream;
    MCDP_CUDA(plan->d_stat_f64.ensure(size_t(2 * E)));
    MCDP_CUDA(plan->d_stat_u64.ensure(size_t(nt * E) + (attr ? size_t(A + E) : 0)));
    MCDP_CUDA(plan->d_stat_u32.ensure(size_t(nb * E)));
    MCDP_CUDA(sl.seeds.ensure(size_t(std::max<int64_t>(n, 1))));
    double* d_sum = plan->d_stat_f64.p;
    double* d_sumsq = d_sum + E;
    unsigned long long* d_late = plan->d_stat_u64.p;
    uint32_t* d_hist = plan->d_stat_u32.p;
    unsigned long long* d_cause_act = nullptr;
    unsigned long long* d_cause_none = nullptr;
    int32_t rc = MCDP_OK;
    do {
        {
            NvtxRange nvtx("mcdp:reduced H2D + clear");
            if (E) MCDP_CUDA_BRK(cudaMemsetAsync(d_sum, 0, size_t(2 * E) * 8, st));
            if (nt && E) MCDP_CUDA_BRK(cudaMemsetAsync(d_late, 0, size_t(nt * E) * 8, st));
            if (nb && E) MCDP_CUDA_BRK(cudaMemsetAsync(d_hist, 0, size_t(nb * E) * 4, st));
            if (attr) {
                d_cause_act = d_late + nt * E;
                d_cause_none = d_cause_act + A;
                if (A + E) MCDP_CUDA_BRK(cudaMemsetAsync(d_cause_act, 0, size_t(A + E) * 8, st));
            }
            if (n) MCDP_CUDA_BRK(cudaMemcpyAsync(sl.seeds.p, seeds, size_t(n) * 4, cudaMemcpyHostToDevice, st));
        }
        rc = run_attribution_locked(plan, sl.seeds.p, 0, n, desc, sum ? d_sum : nullptr, sumsq ? d_sumsq : nullptr,
                                    late ? d_late : nullptr, hist ? d_hist : nullptr, d_cause_act, d_cause_none, st);
        if (rc) break;
        NvtxRange nvtx("mcdp:reduced D2H");
        if (sum && E) MCDP_CUDA_BRK(cudaMemcpyAsync(sum, d_sum, size_t(E) * 8, cudaMemcpyDeviceToHost, st));
        if (sumsq && E) MCDP_CUDA_BRK(cudaMemcpyAsync(sumsq, d_sumsq, size_t(E) * 8, cudaMemcpyDeviceToHost, st));
        if (late && nt && E) MCDP_CUDA_BRK(cudaMemcpyAsync(late, d_late, size_t(nt * E) * 8, cudaMemcpyDeviceToHost, st));
        if (hist && nb && E) MCDP_CUDA_BRK(cudaMemcpyAsync(hist, d_hist, size_t(nb * E) * 4, cudaMemcpyDeviceToHost, st));
        if (attr) {
            if (A) MCDP_CUDA_BRK(cudaMemcpyAsync(cause_act, d_cause_act, size_t(A) * 8, cudaMemcpyDeviceToHost, st));
            if (E) MCDP_CUDA_BRK(cudaMemcpyAsync(cause_none, d_cause_none, size_t(E) * 8, cudaMemcpyDeviceToHost, st));
        }
    } while (false);
    // always drain: copies into caller memory may be in flight even when a later step failed
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess && !rc) rc = fail(MCDP_ERR_CUDA, std::string("stream sync: ") + cudaGetErrorString(e));
    return rc;
}

void* mcdp_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, std::max<size_t>(bytes, 1), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void mcdp_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"

#include "mcdp_multi.inl"
