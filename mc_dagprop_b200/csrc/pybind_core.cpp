// pybind_core.cpp -- Python module `mc_dagprop_b200.monte_carlo._core` (re-exported as
// `mc_dagprop.monte_carlo._core` by the alias package): the reference's Python surface
// (reference src/mc_dagprop/monte_carlo/_core.cpp:366-552, typed by _core.pyi) on top of the
// C ABI of include/mcdp_b200.h.  Same class names, constructor/keyword names, frozen-dataclass
// treatment, numpy-view result properties and RuntimeError behaviour; the engine underneath is
// the sm_100a library.  The data-model structs live in a namespace so the module can coexist in
// one process with the reference's own _core (whose types are global-namespace).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <unistd.h>

#include <cstdint>
#include <cstdlib>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "mcdp_b200.h"

namespace py = pybind11;

namespace mcdp_py {

using EventIndex = int;
using ActivityIndex = int;
using ActivityType = int;
using Preds = std::vector<std::pair<EventIndex, ActivityIndex>>;

struct PairHash {
    size_t operator()(const std::pair<int, int>& p) const noexcept {
        return std::hash<int>{}(p.first) ^ (std::hash<int>{}(p.second) << 1);
    }
};

// ---- data model (reference _core.cpp:36-62) ----
struct EventTimestamp {
    double earliest, latest, actual;
};
struct Event {
    std::string event_id;
    EventTimestamp ts;
};
struct Activity {
    ActivityIndex idx;
    double duration;
    ActivityType activity_type;
};
struct DagContext {
    std::vector<Event> events;
    std::unordered_map<std::pair<EventIndex, EventIndex>, Activity, PairHash> activity_map;
    std::vector<std::pair<EventIndex, Preds>> precedence_list;
    double max_delay;
    DagContext(std::vector<Event> ev, std::unordered_map<std::pair<EventIndex, EventIndex>, Activity, PairHash> am,
               std::vector<std::pair<EventIndex, Preds>> pl, double md)
        : events(std::move(ev)), activity_map(std::move(am)), precedence_list(std::move(pl)), max_delay(md) {}
};

// ---- result memory: pinned host blocks, recycled between calls -----------------------------------------
// Results are written by DMA (mcdp_run_many_host).  Into pageable memory the driver stages every copy through
// its own bounce buffers at a fraction of the PCIe rate, and a fresh std::vector adds a value-initialising pass
// plus one page fault per 4 KB.  Blocks of >= 1 MiB therefore come from cudaHostAlloc (mcdp_host_alloc) and go
// back to a size-bucketed free list when the last SimResult / array that views them dies, so a loop of
// run_many calls pins its memory once.  Small blocks (run(seed) on a small DAG) are plain malloc.
class HostPool {
   public:
    struct Block {
        void* p = nullptr;
        size_t cap = 0;
        bool pinned = false;
    };
    static HostPool& instance() {
        static HostPool* pool = new HostPool();  // leaked on purpose: blocks may outlive static destruction
        return *pool;
    }
    Block take(size_t bytes) {
        Block b;
        if (bytes < kPinThreshold) {
            b.p = std::malloc(bytes ? bytes : 1);
            b.cap = bytes;
            if (!b.p) throw std::bad_alloc();
            return b;
        }
        const size_t cap = bucket(bytes);
        {
            std::lock_guard<std::mutex> lock(mu_);
            auto it = free_.find(cap);
            if (it != free_.end()) {
                b = Block{it->second, cap, true};
                free_.erase(it);
                cached_ -= cap;
                return b;
            }
        }
        b.p = mcdp_host_alloc(cap);
        b.cap = cap;
        b.pinned = b.p != nullptr;
        if (!b.p) {  // pinning refused (limits, no device): pageable memory still works, only slower
            b.p = std::malloc(cap);
            if (!b.p) throw std::bad_alloc();
        }
        return b;
    }
    void give(const Block& b) {
        if (!b.p) return;
        if (!b.pinned) {
            std::free(b.p);
            return;
        }
        {
            std::lock_guard<std::mutex> lock(mu_);
            if (cached_ + b.cap <= limit_) {
                free_.emplace(b.cap, b.p);
                cached_ += b.cap;
                return;
            }
        }
        mcdp_host_free(b.p);
    }
    void set_limit(size_t bytes) {
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> lock(mu_);
            limit_ = bytes;
            while (cached_ > limit_ && !free_.empty()) {
                auto it = std::prev(free_.end());
                drop.push_back(it->second);
                cached_ -= it->first;
                free_.erase(it);
            }
        }
        for (void* p : drop) mcdp_host_free(p);
    }
    size_t limit() const { return limit_; }
    size_t cached() const { return cached_; }

   private:
    static constexpr size_t kPinThreshold = size_t(1) << 20;
    // sizes round up to 1/8 steps of their power of two: a slightly different seed count reuses the block
    static size_t bucket(size_t bytes) {
        size_t p2 = size_t(1) << 20;
        while (p2 < bytes) p2 <<= 1;
        const size_t step = p2 >> 4;
        return (bytes + step - 1) / step * step;
    }
    HostPool() {
        const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
        const size_t ram = pages > 0 && psz > 0 ? size_t(pages) * size_t(psz) : size_t(64) << 30;
        limit_ = std::min<size_t>(ram / 4, size_t(64) << 30);
    }
    std::mutex mu_;
    std::multimap<size_t, void*> free_;
    size_t cached_ = 0, limit_ = 0;
};

// ---- results: one shared batch, per-sample views (reference SimResult, _core.cpp:65-69) ----
// All SimResult objects of one run_many call view ONE batch (a single DMA target): the batch lives until the last
// of them is released.  A caller that keeps one result of a large batch for long should copy its arrays.
struct SimBatch {
    size_t n = 0, E = 0, A = 0;
    HostPool::Block block;
    double* realized = nullptr;
    double* durations = nullptr;
    int32_t* cause = nullptr;
    SimBatch(size_t n_, size_t E_, size_t A_) : n(n_), E(E_), A(A_) {
        auto up = [](size_t b) { return (b + 63) & ~size_t(63); };
        const size_t br = up(n * E * 8), bd = up(n * A * 8), bc = up(n * E * 4);
        block = HostPool::instance().take(br + bd + bc);
        char* base = static_cast<char*>(block.p);
        realized = reinterpret_cast<double*>(base);
        durations = reinterpret_cast<double*>(base + br);
        cause = reinterpret_cast<int32_t*>(base + br + bd);
    }
    ~SimBatch() { HostPool::instance().give(block); }
    SimBatch(const SimBatch&) = delete;
    SimBatch& operator=(const SimBatch&) = delete;
};
struct SimResult {
    std::shared_ptr<SimBatch> batch;
    size_t row = 0;
    double* realized() const { return batch->realized + row * batch->E; }
    double* durations() const { return batch->durations + row * batch->A; }
    int32_t* cause() const { return batch->cause + row * batch->E; }
};

// a numpy array over a pooled block; the block returns to the pool when the array dies
template <typename T>
py::array_t<T> pooled_array(std::vector<py::ssize_t> shape) {
    size_t count = 1;
    for (auto d : shape) count *= size_t(d);
    auto* blk = new HostPool::Block(HostPool::instance().take(count * sizeof(T)));
    py::capsule owner(blk, [](void* q) {
        auto* b = static_cast<HostPool::Block*>(q);
        HostPool::instance().give(*b);
        delete b;
    });
    return py::array_t<T>(shape, static_cast<T*>(blk->p), owner);
}

// ---- generator: parameter tables only; sampling happens on the device ----
struct DistSpec {
    int kind = 0;
    double p0 = 0, p1 = 0, p2 = 0;
    std::vector<double> values, weights;
};
class GenericDelayGenerator {
   public:
    std::map<ActivityType, DistSpec> dist_map_;
    int seed_ = 0;
    // reference _core.cpp:153: seeds an RNG the simulator never reads -- kept, and equally unobservable
    void set_seed(int s) { seed_ = s; }
    void add_constant(ActivityType t, double f) { dist_map_[t] = DistSpec{MCDP_DIST_CONSTANT, f, 0, 0, {}, {}}; }
    void add_exponential(ActivityType t, double lam, double mx) {
        dist_map_[t] = DistSpec{MCDP_DIST_EXPONENTIAL, lam, mx, 0, {}, {}};
    }
    void add_gamma(ActivityType t, double k, double s, double m) { dist_map_[t] = DistSpec{MCDP_DIST_GAMMA, k, s, m, {}, {}}; }
    void add_table(int kind, ActivityType t, std::vector<double> v, std::vector<double> w) {
        if (v.size() != w.size())  // reference _core.cpp:119-120,134-135
            throw std::runtime_error(kind == MCDP_DIST_EMP_ABS
                                         ? "EmpiricalAbsoluteDist: values and weights must have same length"
                                         : "EmpiricalRelativeDist: factors and weights must have same length");
        dist_map_[t] = DistSpec{kind, 0, 0, 0, std::move(v), std::move(w)};
    }
};

[[noreturn]] void throw_last() { throw std::runtime_error(mcdp_last_error()); }

struct FlatDists {
    std::vector<int32_t> type, kind;
    std::vector<double> p0, p1, p2, values, weights;
    std::vector<int64_t> off{0};
    mcdp_dists_desc desc{};
    explicit FlatDists(const GenericDelayGenerator& g) {
        for (const auto& kv : g.dist_map_) {
            type.push_back(kv.first);
            kind.push_back(kv.second.kind);
            p0.push_back(kv.second.p0);
            p1.push_back(kv.second.p1);
            p2.push_back(kv.second.p2);
            values.insert(values.end(), kv.second.values.begin(), kv.second.values.end());
            weights.insert(weights.end(), kv.second.weights.begin(), kv.second.weights.end());
            off.push_back(int64_t(values.size()));
        }
        desc.n_dists = int32_t(type.size());
        desc.dist_type = type.data();
        desc.kind = kind.data();
        desc.p0 = p0.data();
        desc.p1 = p1.data();
        desc.p2 = p2.data();
        desc.tab_off = off.data();
        desc.tab_values = values.data();
        desc.tab_weights = weights.data();
    }
};

template <typename T>
const T* arr_ptr(const py::array_t<T, py::array::c_style | py::array::forcecast>& a) {
    return a.data();
}

// ---- the propagator (reference Simulator, _core.cpp:162-362) ----
class MonteCarloPropagator {
    // one compiled plan per device; every call shards its seeds over the set (a set of one is a plain plan)
    mcdp_planset* set_ = nullptr;
    std::vector<int32_t> devices_;

   public:
    MonteCarloPropagator(const DagContext& ctx, const GenericDelayGenerator& gen, int device, const py::object& devices) {
        // flatten the Python-side containers (the by-value conversion the reference also pays, _core.cpp:425-428)
        std::vector<double> earliest;
        earliest.reserve(ctx.events.size());
        for (const auto& e : ctx.events) earliest.push_back(e.ts.earliest);
        std::vector<int32_t> a_idx, a_type;
        std::vector<double> a_base;
        a_idx.reserve(ctx.activity_map.size());
        for (const auto& kv : ctx.activity_map) {
            a_idx.push_back(kv.second.idx);
            a_base.push_back(kv.second.duration);
            a_type.push_back(kv.second.activity_type);
        }
        std::vector<int32_t> tgt, src, act;
        std::vector<int64_t> off{0};
        for (const auto& entry : ctx.precedence_list) {
            tgt.push_back(entry.first);
            for (const auto& pr : entry.second) {
                src.push_back(pr.first);
                act.push_back(pr.second);
            }
            off.push_back(int64_t(src.size()));
        }
        mcdp_graph_desc g{};
        g.n_events = int32_t(earliest.size());
        g.earliest = earliest.data();
        g.n_act_entries = int32_t(a_idx.size());
        g.act_idx = a_idx.data();
        g.act_base = a_base.data();
        g.act_type = a_type.data();
        g.n_prec_entries = int32_t(tgt.size());
        g.prec_target = tgt.data();
        g.prec_off = off.data();
        g.pred_src = src.data();
        g.pred_act = act.data();
        g.max_delay = ctx.max_delay;
        create(g, gen, device, devices);
    }

    // additive: array ingest without per-object conversion
    MonteCarloPropagator(py::array_t<double, py::array::c_style | py::array::forcecast> earliest,
                         py::array_t<int32_t, py::array::c_style | py::array::forcecast> act_idx,
                         py::array_t<double, py::array::c_style | py::array::forcecast> act_base,
                         py::array_t<int32_t, py::array::c_style | py::array::forcecast> act_type,
                         py::array_t<int32_t, py::array::c_style | py::array::forcecast> prec_target,
                         py::array_t<int64_t, py::array::c_style | py::array::forcecast> prec_off,
                         py::array_t<int32_t, py::array::c_style | py::array::forcecast> pred_src,
                         py::array_t<int32_t, py::array::c_style | py::array::forcecast> pred_act, double max_delay,
                         const GenericDelayGenerator& gen, int device, const py::object& devices) {
        if (act_idx.size() != act_base.size() || act_idx.size() != act_type.size())
            throw std::runtime_error("from_arrays: activity arrays must have the same length");
        if (pred_src.size() != pred_act.size()) throw std::runtime_error("from_arrays: pred_src and pred_act differ in length");
        if (prec_off.size() != prec_target.size() + 1) throw std::runtime_error("from_arrays: prec_off must have len(prec_target)+1 entries");
        if (prec_off.size() && (prec_off.data()[0] != 0 || prec_off.data()[prec_off.size() - 1] != pred_src.size()))
            throw std::runtime_error("from_arrays: prec_off must start at 0 and end at len(pred_src)");
        mcdp_graph_desc g{};
        g.n_events = int32_t(earliest.size());
        g.earliest = earliest.data();
        g.n_act_entries = int32_t(act_idx.size());
        g.act_idx = act_idx.data();
        g.act_base = act_base.data();
        g.act_type = act_type.data();
        g.n_prec_entries = int32_t(prec_target.size());
        g.prec_target = prec_target.data();
        g.prec_off = prec_off.data();
        g.pred_src = pred_src.data();
        g.pred_act = pred_act.data();
        g.max_delay = max_delay;
        create(g, gen, device, devices);
    }

    void create(const mcdp_graph_desc& g, const GenericDelayGenerator& gen, int device, const py::object& devices) {
        if (devices.is_none()) {
            devices_ = {device};
        } else {
            for (const auto& d : devices) devices_.push_back(d.cast<int32_t>());
            if (devices_.empty()) throw std::runtime_error("devices must name at least one CUDA device");
        }
        FlatDists fd(gen);
        int32_t rc;
        {
            py::gil_scoped_release release;
            rc = mcdp_planset_create(&g, &fd.desc, devices_.data(), int32_t(devices_.size()), &set_);
        }
        if (rc != MCDP_OK) throw_last();
    }

    ~MonteCarloPropagator() { mcdp_planset_destroy(set_); }
    MonteCarloPropagator(const MonteCarloPropagator&) = delete;
    MonteCarloPropagator& operator=(const MonteCarloPropagator&) = delete;

    mcdp_plan* plan0() const { return mcdp_planset_plan(set_, 0); }
    int node_count() const { return mcdp_plan_node_count(plan0()); }
    int activity_count() const { return mcdp_plan_activity_count(plan0()); }
    int level_count() const { return mcdp_plan_level_count(plan0()); }
    int device() const { return mcdp_plan_device(plan0()); }
    std::vector<int32_t> devices() const { return devices_; }
    void set_option(int option, int64_t value) {
        if (mcdp_planset_set_option(set_, option, value) != MCDP_OK) throw_last();
    }

    std::shared_ptr<SimBatch> run_batch(const std::vector<int>& seeds) {
        auto b = std::make_shared<SimBatch>(seeds.size(), size_t(node_count()), size_t(activity_count()));
        static_assert(sizeof(int) == sizeof(int32_t), "seeds are C ints");
        int32_t rc = mcdp_run_many_host_multi(set_, reinterpret_cast<const int32_t*>(seeds.data()), int64_t(b->n), b->realized,
                                              b->durations, b->cause);
        if (rc != MCDP_OK) throw_last();
        return b;
    }

    SimResult run(int seed) {  // reference _core.cpp:312-353
        std::shared_ptr<SimBatch> b;
        {
            py::gil_scoped_release release;
            b = run_batch(std::vector<int>{seed});
        }
        return SimResult{b, 0};
    }

    std::vector<SimResult> run_many(const std::vector<int>& seeds) {  // reference _core.cpp:355-361
        std::shared_ptr<SimBatch> b;
        {
            py::gil_scoped_release release;
            b = run_batch(seeds);
        }
        std::vector<SimResult> out;
        out.reserve(seeds.size());
        for (size_t i = 0; i < seeds.size(); ++i) out.push_back(SimResult{b, i});
        return out;
    }

    // additive: one [n,E] / [n,A] / [n,E] array triple instead of n SimResult objects
    py::tuple run_many_arrays(py::array_t<int32_t, py::array::c_style | py::array::forcecast> seeds) {
        const int64_t n = seeds.size(), E = node_count(), A = activity_count();
        py::array_t<double> r = pooled_array<double>({n, E}), d = pooled_array<double>({n, A});
        py::array_t<int32_t> c = pooled_array<int32_t>({n, E});
        int32_t rc;
        {
            py::gil_scoped_release release;
            rc = mcdp_run_many_host_multi(set_, seeds.data(), n, r.mutable_data(), d.mutable_data(), c.mutable_data());
        }
        if (rc != MCDP_OK) throw_last();
        return py::make_tuple(r, d, c);
    }

    // additive: duration injection (propagation phase only, _core.cpp:332-350)
    py::tuple run_with_durations(py::array_t<double, py::array::c_style | py::array::forcecast> durations) {
        const int64_t E = node_count(), A = activity_count();
        if (durations.ndim() != 2 || durations.shape(1) != A)
            throw std::runtime_error("run_with_durations: durations must have shape [n, activity_count()]");
        const int64_t n = durations.shape(0);
        py::array_t<double> r({n, E});
        py::array_t<int32_t> c({n, E});
        int32_t rc;
        {
            py::gil_scoped_release release;
            rc = mcdp_run_injected_host_multi(set_, durations.data(), n, r.mutable_data(), c.mutable_data());
        }
        if (rc != MCDP_OK) throw_last();
        return py::make_tuple(r, c);
    }

    // additive: fused per-event statistics of realized - earliest
    py::dict run_many_reduced(py::array_t<int32_t, py::array::c_style | py::array::forcecast> seeds,
                              std::vector<double> thresholds, int n_bins, double hist_lo, double hist_hi,
                              bool cause_counts) {
        const int64_t n = seeds.size(), E = node_count(), A = activity_count();
        if (thresholds.size() > MCDP_MAX_THRESHOLDS) throw std::runtime_error("at most 4 thresholds");
        mcdp_stats_desc desc{};
        desc.n_thresholds = int32_t(thresholds.size());
        for (size_t i = 0; i < thresholds.size(); ++i) desc.thresholds[i] = thresholds[i];
        desc.n_bins = n_bins;
        desc.hist_lo = hist_lo;
        desc.hist_hi = hist_hi;
        py::array_t<double> sum(E), sumsq(E);
        py::array_t<unsigned long long> late({int64_t(thresholds.size()), E});
        py::array_t<uint32_t> hist({E, int64_t(n_bins)});
        // delay-cause attribution: how often each activity was the binding predecessor of its target event
        py::array_t<unsigned long long> cause_act(cause_counts ? A : 0), cause_none(cause_counts ? E : 0);
        int32_t rc;
        {
            py::gil_scoped_release release;
            rc = mcdp_run_attribution_host_multi(set_, seeds.data(), n, &desc, sum.mutable_data(), sumsq.mutable_data(),
                                           late.size() ? late.mutable_data() : nullptr,
                                           hist.size() ? hist.mutable_data() : nullptr,
                                           cause_counts ? cause_act.mutable_data() : nullptr,
                                           cause_counts ? cause_none.mutable_data() : nullptr);
        }
        if (rc != MCDP_OK) throw_last();
        py::dict out;
        out["n"] = n;
        out["sum"] = sum;
        out["sumsq"] = sumsq;
        out["late"] = late;
        out["hist"] = hist;
        if (cause_counts) {
            out["cause_activity"] = cause_act;
            out["cause_none"] = cause_none;
        }
        return out;
    }
};

}  // namespace mcdp_py

using namespace mcdp_py;

PYBIND11_MODULE(_core, m) {
    m.doc() = "Core Monte-Carlo DAG-propagation simulator (B200 / sm_100a engine)";

    m.def(
        "set_pinned_cache_limit", [](size_t bytes) { HostPool::instance().set_limit(bytes); }, py::arg("bytes"),
        "Upper bound on pinned result memory kept for reuse between calls (default: min(RAM / 4, 64 GiB)); 0 frees the cache");
    m.def("pinned_cache_bytes", []() { return HostPool::instance().cached(); }, "Pinned result memory currently cached for reuse");

    py::class_<EventTimestamp> ts_cls(m, "EventTimestamp");
    ts_cls
        .def(py::init<double, double, double>(), py::arg("earliest"), py::arg("latest"), py::arg("actual"),
             "Create an event timestamp (earliest, latest, actual).")
        .def_readwrite("earliest", &EventTimestamp::earliest, "Earliest bound")
        .def_readwrite("latest", &EventTimestamp::latest, "Latest bound")
        .def_readwrite("actual", &EventTimestamp::actual, "Scheduled time")
        .def("__repr__", [](const EventTimestamp& ts) {
            return py::str("EventTimestamp(earliest={}, latest={}, actual={})").format(ts.earliest, ts.latest, ts.actual);
        });

    py::class_<Event> event_cls(m, "Event");
    event_cls
        .def(py::init<std::string, EventTimestamp>(), py::arg("event_id"), py::arg("timestamp"),
             "An event node with its ID and timestamp")
        .def_readwrite("event_id", &Event::event_id, "Node identifier")
        .def_readwrite("timestamp", &Event::ts, "Event timing info")
        .def("__repr__", [](const Event& ev) {
            return py::str("Event(event_id={}, timestamp={})").format(py::repr(py::cast(ev.event_id)), py::repr(py::cast(ev.ts)));
        });

    py::class_<Activity> activity_cls(m, "Activity");
    activity_cls
        .def(py::init<ActivityIndex, double, ActivityType>(), py::arg("idx"), py::arg("minimal_duration"),
             py::arg("activity_type"), "An activity (edge) with index, base duration and type")
        .def_readwrite("idx", &Activity::idx, "Index of the activity")
        .def_readwrite("minimal_duration", &Activity::duration, "Base duration")
        .def_readwrite("activity_type", &Activity::activity_type, "Type ID for delay dist.")
        .def("__repr__", [](const Activity& a) {
            return py::str("Activity(idx={}, minimal_duration={}, activity_type={})").format(a.idx, a.duration, a.activity_type);
        });

    py::class_<DagContext> ctx_cls(m, "DagContext");
    ctx_cls
        .def(py::init<std::vector<Event>, std::unordered_map<std::pair<EventIndex, EventIndex>, Activity, PairHash>,
                      std::vector<std::pair<EventIndex, Preds>>, double>(),
             py::arg("events"), py::arg("activities"), py::arg("precedence_list"), py::arg("max_delay"),
             "Wraps a DAG: events, activity_map, precedence_list, max_delay")
        .def_readwrite("events", &DagContext::events)
        .def_readwrite("activities", &DagContext::activity_map)
        .def_readwrite("precedence_list", &DagContext::precedence_list)
        .def_readwrite("max_delay", &DagContext::max_delay)
        .def("__repr__", [](const DagContext& ctx) {
            return py::str("DagContext(events={}, activities={}, precedence_list={}, max_delay={})")
                .format(py::repr(py::cast(ctx.events)), py::repr(py::cast(ctx.activity_map)),
                        py::repr(py::cast(ctx.precedence_list)), ctx.max_delay);
        });

    // frozen dataclasses, as the reference does at import (_core.cpp:446-488)
    py::object dataclass_fn = py::module_::import("dataclasses").attr("dataclass");
    py::dict dc_opts;
    dc_opts["frozen"] = true;
    dc_opts["slots"] = true;
    dc_opts["init"] = false;
    py::object dataclass = dataclass_fn(**dc_opts);
    py::module types_mod = py::module_::import("mc_dagprop_b200.types");
    py::object Second = types_mod.attr("Second");
    py::dict ts_ann;
    ts_ann["earliest"] = Second;
    ts_ann["latest"] = Second;
    ts_ann["actual"] = Second;
    ts_cls.attr("__annotations__") = ts_ann;
    dataclass(ts_cls);
    py::dict ev_ann;
    ev_ann["event_id"] = types_mod.attr("EventId");
    ev_ann["timestamp"] = ts_cls;
    event_cls.attr("__annotations__") = ev_ann;
    dataclass(event_cls);
    py::dict act_ann;
    act_ann["idx"] = types_mod.attr("ActivityIndex");
    act_ann["minimal_duration"] = Second;
    act_ann["activity_type"] = types_mod.attr("ActivityType");
    activity_cls.attr("__annotations__") = act_ann;
    dataclass(activity_cls);
    py::object typing = py::module_::import("typing");
    py::dict ctx_ann;
    ctx_ann["events"] = typing.attr("Sequence");
    ctx_ann["activities"] = typing.attr("Mapping");
    ctx_ann["precedence_list"] = typing.attr("Sequence");
    ctx_ann["max_delay"] = Second;
    ctx_cls.attr("__annotations__") = ctx_ann;
    dataclass(ctx_cls);

    // SimResult: zero-copy numpy views whose base is the SimResult object (reference _core.cpp:491-516)
    py::class_<SimResult>(m, "SimResult", py::buffer_protocol())
        .def_buffer([](SimResult& r) -> py::buffer_info {
            return py::buffer_info(r.realized(), sizeof(double), py::format_descriptor<double>::format(), 1,
                                   {r.batch->E}, {sizeof(double)});
        })
        .def_property_readonly(
            "realized", [](const SimResult& r) { return py::array(py::ssize_t(r.batch->E), r.realized(), py::cast(r)); },
            "Final event times as a NumPy array")
        .def_property_readonly(
            "durations", [](const SimResult& r) { return py::array(py::ssize_t(r.batch->A), r.durations(), py::cast(r)); },
            "Per-link durations (incl. extra) as a NumPy array")
        .def_property_readonly(
            "cause_event", [](const SimResult& r) { return py::array(py::ssize_t(r.batch->E), r.cause(), py::cast(r)); },
            "Index of predecessor causing each event as a NumPy array");

    py::class_<GenericDelayGenerator>(m, "GenericDelayGenerator")
        .def(py::init<>(), "Create a new delay-generator")
        .def("set_seed", &GenericDelayGenerator::set_seed, py::arg("seed"), "Set RNG seed for reproducibility")
        .def("add_constant", &GenericDelayGenerator::add_constant, py::arg("activity_type"), py::arg("factor"),
             "Constant: delay = factor * duration")
        .def("add_exponential", &GenericDelayGenerator::add_exponential, py::arg("activity_type"), py::arg("lambda_"),
             py::arg("max_scale"), "Exponential(lambda) truncated at max_scale")
        .def("add_gamma", &GenericDelayGenerator::add_gamma, py::arg("activity_type"), py::arg("shape"), py::arg("scale"),
             py::arg("max_scale") = std::numeric_limits<double>::infinity(), "Gamma(shape,scale) truncated at max_scale")
        .def(
            "add_empirical_absolute",
            [](GenericDelayGenerator& g, ActivityType t, std::vector<double> values, std::vector<double> weights) {
                g.add_table(MCDP_DIST_EMP_ABS, t, std::move(values), std::move(weights));
            },
            py::arg("activity_type"), py::arg("values"), py::arg("weights"),
            "Empirical absolute: draw one of your provided values, weighted by weights.")
        .def(
            "add_empirical_relative",
            [](GenericDelayGenerator& g, ActivityType t, std::vector<double> factors, std::vector<double> weights) {
                g.add_table(MCDP_DIST_EMP_REL, t, std::move(factors), std::move(weights));
            },
            py::arg("activity_type"), py::arg("factors"), py::arg("weights"),
            "Empirical relative: draw a factor in [0,inf), then multiply by the activity duration.");

    py::class_<MonteCarloPropagator>(m, "MonteCarloPropagator")
        .def(py::init<const DagContext&, const GenericDelayGenerator&, int, const py::object&>(), py::arg("context"),
             py::arg("generator"), py::arg("device") = 0, py::arg("devices") = py::none(),
             "Construct simulator with context and delay-generator (device: CUDA ordinal; devices: several ordinals -- "
             "every call then shards its seeds over them)")
        .def_static(
            "from_arrays",
            [](py::array_t<double, py::array::c_style | py::array::forcecast> earliest,
               py::array_t<int32_t, py::array::c_style | py::array::forcecast> act_idx,
               py::array_t<double, py::array::c_style | py::array::forcecast> act_base,
               py::array_t<int32_t, py::array::c_style | py::array::forcecast> act_type,
               py::array_t<int32_t, py::array::c_style | py::array::forcecast> prec_target,
               py::array_t<int64_t, py::array::c_style | py::array::forcecast> prec_off,
               py::array_t<int32_t, py::array::c_style | py::array::forcecast> pred_src,
               py::array_t<int32_t, py::array::c_style | py::array::forcecast> pred_act, double max_delay,
               const GenericDelayGenerator& gen, int device, const py::object& devices) {
                return std::make_unique<MonteCarloPropagator>(earliest, act_idx, act_base, act_type, prec_target, prec_off,
                                                              pred_src, pred_act, max_delay, gen, device, devices);
            },
            py::arg("earliest"), py::arg("act_idx"), py::arg("act_base"), py::arg("act_type"), py::arg("prec_target"),
            py::arg("prec_off"), py::arg("pred_src"), py::arg("pred_act"), py::arg("max_delay"), py::arg("generator"),
            py::arg("device") = 0, py::arg("devices") = py::none(), "Construct from flat numpy arrays (no per-object conversion)")
        .def("node_count", &MonteCarloPropagator::node_count, "Number of events")
        .def("activity_count", &MonteCarloPropagator::activity_count, "Number of links")
        .def("level_count", &MonteCarloPropagator::level_count, "Number of topological levels")
        .def("device", &MonteCarloPropagator::device, "CUDA device ordinal (the first one of a multi-device propagator)")
        .def("devices", &MonteCarloPropagator::devices, "CUDA device ordinals the seeds of a call are sharded over")
        .def("set_option", &MonteCarloPropagator::set_option, py::arg("option"), py::arg("value"))
        .def("run", &MonteCarloPropagator::run, py::arg("seed"), "Run single sim")
        .def("run_many", &MonteCarloPropagator::run_many, py::arg("seeds"), "Run batch sims")
        .def("run_many_arrays", &MonteCarloPropagator::run_many_arrays, py::arg("seeds"),
             "Run batch sims, return (realized[n,E], durations[n,A], cause_event[n,E])")
        .def("run_with_durations", &MonteCarloPropagator::run_with_durations, py::arg("durations"),
             "Propagate caller-supplied durations[n,A]; returns (realized[n,E], cause_event[n,E])")
        .def("run_many_reduced", &MonteCarloPropagator::run_many_reduced, py::arg("seeds"),
             py::arg("thresholds") = std::vector<double>{}, py::arg("n_bins") = 0, py::arg("hist_lo") = 0.0,
             py::arg("hist_hi") = 1.0, py::arg("cause_counts") = false,
             "Per-event statistics of realized - earliest without materialising samples; cause_counts adds "
             "cause_activity[A] / cause_none[E]: how often each activity decided its target event / no predecessor did");
}
