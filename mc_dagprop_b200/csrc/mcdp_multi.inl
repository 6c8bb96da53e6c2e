// mcdp_multi.inl -- several devices behind one call (included at the end of mcdp_capi.cu).
//
// The reference scales out through run_many alone (_core.cpp:355-361: a loop over independent seeds, GIL released)
// plus one Simulator per thread (test/test_simulator.py:201-215).  Here one mcdp_planset holds the compiled plan
// once on the host and once per device; every *_multi call shards its seeds into contiguous blocks, one per
// device, and returns exactly what a single device returns:
//   * full outputs: no exchange at all -- each device copies its block of rows straight into the caller's arrays;
//   * reduced statistics: the one exchange step of the system.  Device d sums slice d of every accumulator over
//     all devices by reading the peers' buffers directly (peer_sum_kernel: loads over NVLink / NVSwitch peer
//     mappings, fixed summation order 0..G-1, so the f64 sums do not depend on timing) and copies that slice to
//     the caller: a reduce-scatter whose scatter target is host memory, one PCIe link per device.

struct mcdp_planset {
    std::vector<mcdp_plan*> plans;  // owned
    bool peer_ok = true;            // every pair of distinct devices can address each other's memory
    std::string peer_msg;
    std::vector<cudaEvent_t> done;  // per plan: statistics of the shard are complete
    std::mutex mu;                  // one multi call at a time per set
    ~mcdp_planset() {
        for (size_t i = 0; i < plans.size(); ++i) {
            if (i < done.size() && done[i]) {
                DeviceGuard g(plans[i]->device);
                cudaEventDestroy(done[i]);
            }
            delete plans[i];
        }
    }
};

namespace {

constexpr int kMaxSetDevices = 16;

template <typename T>
struct PeerPtrs {
    const T* src[kMaxSetDevices];
    int n;
};

// dst[i] = sum over devices of src[k][i], i in [begin, end): the reduce step of the reduced mode.  dst may alias
// one of the sources (each element is read by exactly the thread that writes it).
template <typename T>
__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerPtrs<T> srcs, T* dst, int64_t begin, int64_t end) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = begin + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < end; i += stride) {
        T acc = srcs.src[0][i];
        for (int k = 1; k < srcs.n; ++k) acc += srcs.src[k][i];
        dst[i] = acc;
    }
}

template <typename T>
int32_t launch_peer_sum(const PeerPtrs<T>& srcs, T* dst, int64_t begin, int64_t end, int sm_count, cudaStream_t st) {
    if (end <= begin) return MCDP_OK;
    const int64_t blocks = std::min<int64_t>((end - begin + 255) / 256, int64_t(sm_count) * 8);
    peer_sum_kernel<T><<<unsigned(blocks), 256, 0, st>>>(srcs, dst, begin, end);
    MCDP_CUDA(cudaGetLastError());
    return MCDP_OK;
}

// contiguous block of [0, n) owned by shard i of g: whole 128-sample groups except the last block
void shard_block(int64_t n, int i, int g, int64_t* lo, int64_t* hi) {
    const int64_t groups = (n + 127) / 128, per = groups / g, extra = groups % g;
    const int64_t g0 = int64_t(i) * per + std::min<int64_t>(i, extra);
    const int64_t g1 = g0 + per + (i < extra ? 1 : 0);
    *lo = std::min(g0 * 128, n);
    *hi = std::min(g1 * 128, n);
}

// one host thread per device around a synchronous single-device call; the first failure is reported
template <typename F>
int32_t for_each_shard(mcdp_planset* set, int64_t n, F&& call) {
    const int g = int(set->plans.size());
    std::vector<int32_t> rcs(size_t(g), MCDP_OK);
    std::vector<std::string> msgs(static_cast<size_t>(g));
    auto work = [&](int i) {
        int64_t lo, hi;
        shard_block(n, i, g, &lo, &hi);
        rcs[size_t(i)] = call(set->plans[size_t(i)], lo, hi - lo);
        if (rcs[size_t(i)]) msgs[size_t(i)] = g_err;  // the message is thread-local: hand it to the caller's thread
    };
    std::vector<std::thread> threads;
    for (int i = 1; i < g; ++i) {
        int64_t lo, hi;
        shard_block(n, i, g, &lo, &hi);
        if (hi <= lo) continue;  // a call smaller than the set (run(seed)): no thread for an empty block
        try {
            threads.emplace_back(work, i);
        } catch (const std::system_error&) {
            work(i);  // no thread to be had: this block runs on the caller's thread
        }
    }
    work(0);
    for (auto& t : threads) t.join();
    for (int i = 0; i < g; ++i)
        if (rcs[size_t(i)]) return fail(rcs[size_t(i)], "device " + std::to_string(set->plans[size_t(i)]->device) + ": " + msgs[size_t(i)]);
    return MCDP_OK;
}

}  // namespace

extern "C" {

int32_t mcdp_planset_create(const mcdp_graph_desc* graph, const mcdp_dists_desc* dists, const int32_t* devices,
                            int32_t n_devices, mcdp_planset** out) {
    if (!graph || !dists || !devices || !out) return fail(MCDP_ERR_ARG, "null argument");
    *out = nullptr;
    if (n_devices < 1 || n_devices > kMaxSetDevices) return fail(MCDP_ERR_ARG, "a plan set holds 1..16 devices");
    std::shared_ptr<HostPlan> host;
    {
        const int32_t rc = compile_host_plan(graph, dists, &host);
        if (rc) return rc;
    }
    std::unique_ptr<mcdp_planset> set(new (std::nothrow) mcdp_planset());
    if (!set) return fail(MCDP_ERR_ARG, "out of host memory");
    for (int i = 0; i < n_devices; ++i) {
        mcdp_plan* plan = nullptr;
        const int32_t rc = plan_from_host(host, devices[i], &plan);
        if (rc) return rc;
        set->plans.push_back(plan);
    }
    set->done.assign(size_t(n_devices), nullptr);
    for (int i = 0; i < n_devices; ++i) {
        DeviceGuard g(devices[i]);
        if (devices[i] >= 0) MCDP_CUDA(cudaEventCreateWithFlags(&set->done[size_t(i)], cudaEventDisableTiming));
        for (int j = 0; j < n_devices; ++j) {
            if (devices[i] < 0 || devices[j] < 0 || devices[i] == devices[j]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) {
                cudaGetLastError();
                set->peer_ok = false;
                set->peer_msg = "devices " + std::to_string(devices[i]) + " and " + std::to_string(devices[j]) +
                                " cannot address each other's memory";
                continue;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                set->peer_ok = false;
                set->peer_msg = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
            }
            cudaGetLastError();
        }
    }
    *out = set.release();
    return MCDP_OK;
}

void mcdp_planset_destroy(mcdp_planset* set) { delete set; }
int32_t mcdp_planset_size(const mcdp_planset* set) { return set ? int32_t(set->plans.size()) : 0; }
mcdp_plan* mcdp_planset_plan(mcdp_planset* set, int32_t i) {
    return set && i >= 0 && size_t(i) < set->plans.size() ? set->plans[size_t(i)] : nullptr;
}
int32_t mcdp_planset_set_option(mcdp_planset* set, int32_t option, int64_t value) {
    if (!set) return fail(MCDP_ERR_ARG, "null plan set");
    for (mcdp_plan* p : set->plans) {
        const int32_t rc = mcdp_plan_set_option(p, option, value);
        if (rc) return rc;
    }
    return MCDP_OK;
}

int32_t mcdp_run_many_host_multi(mcdp_planset* set, const int32_t* seeds, int64_t n, double* realized, double* durations,
                                 int32_t* cause) {
    if (!set) return fail(MCDP_ERR_ARG, "null plan set");
    if (n < 0) return fail(MCDP_ERR_ARG, "n must be non-negative");
    if (n > 0 && !seeds) return fail(MCDP_ERR_ARG, "null seeds");
    if (set->plans.size() == 1) return mcdp_run_many_host(set->plans[0], seeds, n, realized, durations, cause);
    std::lock_guard<std::mutex> lock(set->mu);
    const int64_t E = set->plans[0]->host.E, A = set->plans[0]->host.A;
    return for_each_shard(set, n, [&](mcdp_plan* plan, int64_t lo, int64_t cnt) {
        return mcdp_run_many_host(plan, seeds + lo, cnt, realized ? realized + lo * E : nullptr,
                                  durations ? durations + lo * A : nullptr, cause ? cause + lo * E : nullptr);
    });
}

int32_t mcdp_run_injected_host_multi(mcdp_planset* set, const double* durations, int64_t n, double* realized, int32_t* cause) {
    if (!set) return fail(MCDP_ERR_ARG, "null plan set");
    if (n < 0) return fail(MCDP_ERR_ARG, "n must be non-negative");
    if (set->plans.size() == 1) return mcdp_run_injected_host(set->plans[0], durations, n, realized, cause);
    std::lock_guard<std::mutex> lock(set->mu);
    const int64_t E = set->plans[0]->host.E, A = set->plans[0]->host.A;
    if (!durations && A > 0 && n > 0) return fail(MCDP_ERR_ARG, "null durations");
    return for_each_shard(set, n, [&](mcdp_plan* plan, int64_t lo, int64_t cnt) {
        return mcdp_run_injected_host(plan, durations ? durations + lo * A : nullptr, cnt, realized ? realized + lo * E : nullptr,
                                      cause ? cause + lo * E : nullptr);
    });
}

int32_t mcdp_run_reduced_host_multi(mcdp_planset* set, const int32_t* seeds, int64_t n, const mcdp_stats_desc* desc,
                                    double* sum, double* sumsq, unsigned long long* late, uint32_t* hist) {
    return mcdp_run_attribution_host_multi(set, seeds, n, desc, sum, sumsq, late, hist, nullptr, nullptr);
}

int32_t mcdp_run_attribution_host_multi(mcdp_planset* set, const int32_t* seeds, int64_t n, const mcdp_stats_desc* desc,
                                        double* sum, double* sumsq, unsigned long long* late, uint32_t* hist,
                                        unsigned long long* cause_act, unsigned long long* cause_none) {
    if (!set || !desc) return fail(MCDP_ERR_ARG, "null argument");
    if (set->plans.size() == 1)
        return mcdp_run_attribution_host(set->plans[0], seeds, n, desc, sum, sumsq, late, hist, cause_act, cause_none);
    if (n < 0) return fail(MCDP_ERR_ARG, "n must be non-negative");
    if (n > 0 && !seeds) return fail(MCDP_ERR_ARG, "null seeds");
    const bool attr = cause_act != nullptr || cause_none != nullptr;
    if (attr && (!cause_act || !cause_none)) return fail(MCDP_ERR_ARG, "cause_act and cause_none go together");
    if (!set->peer_ok) return fail(MCDP_ERR_CUDA, "reduced statistics over several devices need peer access: " + set->peer_msg);
    const int g = int(set->plans.size());
    for (mcdp_plan* p : set->plans)
        if (p->device < 0) return fail(MCDP_ERR_CUDA, "plan set was created host-only: there is no CPU execution path");
    std::lock_guard<std::mutex> set_lock(set->mu);
    // every plan of the set is busy for the whole call (always taken in set order)
    std::vector<std::unique_lock<std::mutex>> locks;
    for (mcdp_plan* p : set->plans) locks.emplace_back(p->mu);
    const int64_t E = set->plans[0]->host.E, A = set->plans[0]->host.A;
    const int64_t nt = std::max(desc->n_thresholds, 0), nb = std::max(desc->n_bins, 0);
    const int64_t n_f64 = 2 * E, n_u64 = nt * E + (attr ? A + E : 0), n_u32 = nb * E;
    int32_t rc = MCDP_OK;
    // ---- phase A: every device clears its accumulators and sweeps its block of seeds (all asynchronous) ----
    for (int i = 0; i < g && !rc; ++i) {
        mcdp_plan* plan = set->plans[size_t(i)];
        DeviceGuard guard(plan->device);
        HostSlot& sl = plan->slots[0];
        int64_t lo, hi;
        shard_block(n, i, g, &lo, &hi);
        do {
            if (!sl.stream) MCDP_CUDA_BRK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            MCDP_CUDA_BRK(plan->d_stat_f64.ensure(size_t(n_f64)));
            MCDP_CUDA_BRK(plan->d_stat_u64.ensure(size_t(n_u64)));
            MCDP_CUDA_BRK(plan->d_stat_u32.ensure(size_t(n_u32)));
            MCDP_CUDA_BRK(sl.seeds.ensure(size_t(std::max<int64_t>(hi - lo, 1))));
            if (n_f64) MCDP_CUDA_BRK(cudaMemsetAsync(plan->d_stat_f64.p, 0, size_t(n_f64) * 8, sl.stream));
            if (n_u64) MCDP_CUDA_BRK(cudaMemsetAsync(plan->d_stat_u64.p, 0, size_t(n_u64) * 8, sl.stream));
            if (n_u32) MCDP_CUDA_BRK(cudaMemsetAsync(plan->d_stat_u32.p, 0, size_t(n_u32) * 4, sl.stream));
            if (hi > lo) MCDP_CUDA_BRK(cudaMemcpyAsync(sl.seeds.p, seeds + lo, size_t(hi - lo) * 4, cudaMemcpyHostToDevice, sl.stream));
            unsigned long long* d_late = plan->d_stat_u64.p;
            rc = run_attribution_locked(plan, sl.seeds.p, 0, hi - lo, desc, sum ? plan->d_stat_f64.p : nullptr,
                                        sumsq ? plan->d_stat_f64.p + E : nullptr, late ? d_late : nullptr,
                                        hist ? plan->d_stat_u32.p : nullptr, attr ? d_late + nt * E : nullptr,
                                        attr ? d_late + nt * E + A : nullptr, sl.stream);
            if (rc) break;
            MCDP_CUDA_BRK(cudaEventRecord(set->done[size_t(i)], sl.stream));
        } while (false);
    }
    // ---- phase B: device d folds slice d of every accumulator over all devices and hands it to the caller ----
    for (int i = 0; i < g && !rc; ++i) {
        mcdp_plan* plan = set->plans[size_t(i)];
        DeviceGuard guard(plan->device);
        cudaStream_t st = plan->slots[0].stream;
        NvtxRange nvtx("mcdp:peer reduce + D2H");
        do {
            for (int j = 0; j < g; ++j)
                if (j != i) MCDP_CUDA_BRK(cudaStreamWaitEvent(st, set->done[size_t(j)], 0));
            if (rc) break;
            PeerPtrs<double> pf{};
            PeerPtrs<unsigned long long> pu{};
            PeerPtrs<uint32_t> ph{};
            pf.n = pu.n = ph.n = g;
            for (int j = 0; j < g; ++j) {
                pf.src[j] = set->plans[size_t(j)]->d_stat_f64.p;
                pu.src[j] = set->plans[size_t(j)]->d_stat_u64.p;
                ph.src[j] = set->plans[size_t(j)]->d_stat_u32.p;
            }
            auto slice = [&](int64_t len, int64_t* b, int64_t* e) {
                *b = len * i / g;
                *e = len * (i + 1) / g;
            };
            // device array [begin, end) -> the caller's arrays laid out back to back in `segs`
            auto copy_out = [&](const char* dev_base, size_t elem, int64_t b, int64_t e,
                                std::initializer_list<std::pair<void*, int64_t>> segs) -> int32_t {
                int64_t off = 0;
                for (const auto& sg : segs) {
                    const int64_t s0 = std::max(b, off), s1 = std::min(e, off + sg.second);
                    if (sg.first && s1 > s0)
                        MCDP_CUDA(cudaMemcpyAsync(static_cast<char*>(sg.first) + size_t(s0 - off) * elem, dev_base + size_t(s0) * elem,
                                                  size_t(s1 - s0) * elem, cudaMemcpyDeviceToHost, st));
                    off += sg.second;
                }
                return MCDP_OK;
            };
            int64_t b, e;
            slice(n_f64, &b, &e);
            rc = launch_peer_sum<double>(pf, plan->d_stat_f64.p, b, e, plan->sm_count, st);
            if (!rc) rc = copy_out(reinterpret_cast<const char*>(plan->d_stat_f64.p), 8, b, e, {{sum, E}, {sumsq, E}});
            if (rc) break;
            slice(n_u64, &b, &e);
            rc = launch_peer_sum<unsigned long long>(pu, plan->d_stat_u64.p, b, e, plan->sm_count, st);
            if (!rc)
                rc = copy_out(reinterpret_cast<const char*>(plan->d_stat_u64.p), 8, b, e,
                              {{late, nt * E}, {attr ? cause_act : nullptr, attr ? A : 0}, {attr ? cause_none : nullptr, attr ? E : 0}});
            if (rc) break;
            slice(n_u32, &b, &e);
            rc = launch_peer_sum<uint32_t>(ph, plan->d_stat_u32.p, b, e, plan->sm_count, st);
            if (!rc) rc = copy_out(reinterpret_cast<const char*>(plan->d_stat_u32.p), 4, b, e, {{hist, nb * E}});
        } while (false);
    }
    // ---- always drain every stream that may have work or copies in flight ----
    for (mcdp_plan* plan : set->plans) {
        if (!plan->slots[0].stream) continue;
        DeviceGuard guard(plan->device);
        const cudaError_t e = cudaStreamSynchronize(plan->slots[0].stream);
        if (e != cudaSuccess && !rc) rc = fail(MCDP_ERR_CUDA, std::string("stream sync: ") + cudaGetErrorString(e));
    }
    return rc;
}

}  // extern "C"
