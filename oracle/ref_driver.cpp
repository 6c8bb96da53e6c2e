// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A flat-array C ABI around the UNMODIFIED reference engine.  The reference
// translation unit is compiled *where it lies* (the Makefile passes
// -DMCDP_REF_CORE_CPP="/root/reference/src/mc_dagprop/monte_carlo/_core.cpp"), no
// reference source is copied into this repository.  What this file adds is only
// the marshalling from flat arrays to the reference's own containers
// (DagContext / GenericDelayGenerator, _core.cpp:52-62,146-159) and a loop over
// Simulator::run (_core.cpp:312-353) -- the very loop run_many performs
// (_core.cpp:355-361) minus the pybind11 GIL release, so that ctypes callers
// may run one Simulator per host thread (the reference's only parallel
// pattern, test/test_simulator.py:201-215).
//
// Output lands in oracle/_ref/libmcdp_ref.so (git-ignored, travels to the GPU
// box).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
// may load it.

#include MCDP_REF_CORE_CPP

#include <cstdint>
#include <cstring>
#include <exception>
#include <string>

namespace {
thread_local std::string g_ref_error;

struct RefHandle {
    Simulator* sim;
};
}  // namespace

extern "C" {

const char* mcdp_ref_last_error() { return g_ref_error.c_str(); }

// dist_kind: 0 constant (p0=factor), 1 exponential (p0=lambda, p1=max_scale),
// 2 gamma (p0=shape, p1=scale, p2=max_scale), 3 empirical absolute, 4 empirical
// relative (table = tab_values/tab_weights[tab_off[t] .. tab_off[t+1])).
void* mcdp_ref_create(int32_t n_events, const double* earliest,
                      int32_t n_act_entries, const int32_t* act_idx, const double* act_base,
                      const int32_t* act_type,
                      int32_t n_prec_entries, const int32_t* prec_target, const int64_t* prec_off,
                      const int32_t* pred_src, const int32_t* pred_act, double max_delay,
                      int32_t n_dists, const int32_t* dist_type, const int32_t* dist_kind,
                      const double* p0, const double* p1, const double* p2,
                      const int64_t* tab_off, const double* tab_values, const double* tab_weights) {
    try {
        std::vector<Event> events;
        events.reserve(n_events);
        for (int i = 0; i < n_events; ++i) {
            events.push_back(Event{std::to_string(i), EventTimestamp{earliest[i], earliest[i], earliest[i]}});
        }
        std::unordered_map<std::pair<EventIndex, EventIndex>, Activity> amap;
        amap.reserve(n_act_entries);
        for (int i = 0; i < n_act_entries; ++i) {
            // keys are ignored by the engine (_core.cpp:214-229 iterates values only)
            amap[std::make_pair(i, -1 - i)] = Activity{act_idx[i], act_base[i], act_type[i]};
        }
        std::vector<std::pair<EventIndex, Preds>> plist;
        plist.reserve(n_prec_entries);
        for (int i = 0; i < n_prec_entries; ++i) {
            Preds pr;
            for (int64_t k = prec_off[i]; k < prec_off[i + 1]; ++k) pr.emplace_back(pred_src[k], pred_act[k]);
            plist.emplace_back(prec_target[i], std::move(pr));
        }
        DagContext ctx(std::move(events), std::move(amap), std::move(plist), max_delay);
        GenericDelayGenerator gen;
        for (int t = 0; t < n_dists; ++t) {
            switch (dist_kind[t]) {
                case 0: gen.add_constant(dist_type[t], p0[t]); break;
                case 1: gen.add_exponential(dist_type[t], p0[t], p1[t]); break;
                case 2: gen.add_gamma(dist_type[t], p0[t], p1[t], p2[t]); break;
                case 3:
                case 4: {
                    std::vector<double> v(tab_values + tab_off[t], tab_values + tab_off[t + 1]);
                    std::vector<double> w(tab_weights + tab_off[t], tab_weights + tab_off[t + 1]);
                    if (dist_kind[t] == 3)
                        gen.dist_map_[dist_type[t]] = EmpiricalAbsoluteDist{std::move(v), std::move(w)};
                    else
                        gen.dist_map_[dist_type[t]] = EmpiricalRelativeDist{std::move(v), std::move(w)};
                    break;
                }
                default: throw std::runtime_error("ref_driver: unknown dist kind");
            }
        }
        auto* h = new RefHandle{new Simulator(std::move(ctx), std::move(gen))};
        return h;
    } catch (const std::exception& e) {
        g_ref_error = e.what();
        return nullptr;
    }
}

void mcdp_ref_destroy(void* handle) {
    auto* h = static_cast<RefHandle*>(handle);
    if (!h) return;
    delete h->sim;
    delete h;
}

int32_t mcdp_ref_node_count(void* handle) { return static_cast<RefHandle*>(handle)->sim->node_count(); }
int32_t mcdp_ref_activity_count(void* handle) { return static_cast<RefHandle*>(handle)->sim->activity_count(); }

// Sample-major outputs: realized[n][E], durations[n][A], cause[n][E]; any may be NULL.
int32_t mcdp_ref_run_many(void* handle, const int32_t* seeds, int64_t n, double* realized, double* durations,
                          int32_t* cause) {
    try {
        Simulator* sim = static_cast<RefHandle*>(handle)->sim;
        const size_t E = sim->node_count(), A = sim->activity_count();
        for (int64_t s = 0; s < n; ++s) {
            SimResult r = sim->run(seeds[s]);
            if (realized) std::memcpy(realized + s * E, r.realized.data(), E * sizeof(double));
            if (durations) std::memcpy(durations + s * A, r.durations.data(), A * sizeof(double));
            if (cause) std::memcpy(cause + s * E, r.cause_event.data(), E * sizeof(int32_t));
        }
        return 0;
    } catch (const std::exception& e) {
        g_ref_error = e.what();
        return 1;
    }
}

}  // extern "C"
