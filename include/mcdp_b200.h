/* mcdp_b200.h -- C ABI of libmcdp_b200.so, the B200 (sm_100a) implementation of the
 * Monte-Carlo hot path of WonJayne/mc_dagprop.
 *
 * The reference has no FFI seam: its pybind11 module calls `Simulator` directly
 * (reference src/mc_dagprop/monte_carlo/_core.cpp:545-551).  This header is the seam a
 * maintainer would cut there; each entry point names the reference code it replaces.
 * INTEGRATION.md shows the C++ and ctypes stubs that bind it.
 *
 * Conventions: plain C types, caller-owned buffers, no exceptions across the boundary.
 * Every function that can fail returns an int status (MCDP_OK == 0); mcdp_last_error()
 * returns the calling thread's last message.  The binding layer rethrows it as Python
 * RuntimeError where the reference throws std::runtime_error.  There is NO CPU fallback:
 * without a usable CUDA device every run call fails with MCDP_ERR_CUDA.
 *
 * Device arrays are event-major, sample-minor: row r (an event or an activity index) of
 * sample column s lives at base[r * ld + s]; ld (elements) must be a multiple of 64 and >= n
 * (columns n..ld-1 are padding the kernels may write), bases 16-byte aligned, so that one
 * warp reads/writes 64 adjacent samples of a row as 16-byte vectors.  Host arrays of the *_host calls are sample-major ([n][E], [n][A]) exactly like
 * the reference's per-sample SimResult vectors (_core.cpp:65-69).
 */
#ifndef MCDP_B200_H
#define MCDP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCDP_ABI_VERSION 5 /* 2: + mcdp_run_attribution_device / _host; 3: + mcdp_plan_launch_shape,
                              MCDP_OPT_SAMPLES_PER_LANE (both additive); 4: + MCDP_OPT_CLUSTER_SIZE, mcdp_planset_*,
                              launch_shape slot 6 reports the cluster size (generator contract mcdp-philox-v2);
                              5: + MCDP_OPT_SMALL_CALL_MAX, mcdp_analytic_last_profile (both additive) */

enum {
    MCDP_OK = 0,
    MCDP_ERR_INVALID = 1, /* what the reference reports as std::runtime_error (and the UB it does not check) */
    MCDP_ERR_CUDA = 2,    /* CUDA runtime / no device / out of memory */
    MCDP_ERR_ARG = 3      /* null pointer, bad size, bad layout */
};

/* Distribution kinds, one per reference Dist struct. p* meaning in comments. */
enum {
    MCDP_DIST_CONSTANT = 0,    /* ConstantDist          _core.cpp:72-76    p0 = factor                       */
    MCDP_DIST_EXPONENTIAL = 1, /* ExponentialDist       _core.cpp:78-90    p0 = lambda (mean), p1 = max_scale */
    MCDP_DIST_GAMMA = 2,       /* GammaDist             _core.cpp:92-105   p0 = shape, p1 = scale, p2 = max_scale */
    MCDP_DIST_EMP_ABS = 3,     /* EmpiricalAbsoluteDist _core.cpp:110-126  table: values, weights            */
    MCDP_DIST_EMP_REL = 4      /* EmpiricalRelativeDist _core.cpp:129-141  table: factors, weights           */
};

/* DagContext (_core.cpp:52-62) flattened.  Only Event.ts.earliest is read by the engine
 * (_core.cpp:319,333); activity_map keys are ignored (_core.cpp:214-229).  Precedence entry i
 * targets prec_target[i] with predecessors (pred_src[k], pred_act[k]), k in
 * [prec_off[i], prec_off[i+1]), in the caller's order (it decides ties, _core.cpp:343). */
typedef struct {
    int32_t n_events;
    const double* earliest; /* [n_events] */
    int32_t n_act_entries;
    const int32_t* act_idx;  /* Activity.idx; activity_count() = max + 1, gaps = zero-duration links (_core.cpp:213-219) */
    const double* act_base;  /* Activity.minimal_duration */
    const int32_t* act_type; /* Activity.activity_type */
    int32_t n_prec_entries;
    const int32_t* prec_target; /* [n_prec_entries] */
    const int64_t* prec_off;    /* [n_prec_entries + 1] */
    const int32_t* pred_src;    /* [prec_off[n_prec_entries]] */
    const int32_t* pred_act;
    double max_delay;
} mcdp_graph_desc;

/* GenericDelayGenerator::dist_map_ (_core.cpp:146-159) flattened: entry t applies to
 * activity_type dist_type[t]; a later entry for the same type replaces an earlier one.
 * Table of entry t: tab_values/tab_weights[tab_off[t] .. tab_off[t+1]). */
typedef struct {
    int32_t n_dists;
    const int32_t* dist_type;
    const int32_t* kind;
    const double* p0;
    const double* p1;
    const double* p2;
    const int64_t* tab_off; /* [n_dists + 1] */
    const double* tab_values;
    const double* tab_weights;
} mcdp_dists_desc;

/* Fused per-event statistics of the delay  realized[e] - earliest[e]  (north_star "reduced
 * mode"; no reference counterpart -- it replaces the numpy post-processing of
 * demo/monte_carlo.py:58-63).  All outputs ACCUMULATE (+=) so chunks, launches and GPUs add up. */
#define MCDP_MAX_THRESHOLDS 4
typedef struct {
    int32_t n_thresholds;                  /* 0..4: counts of delay > thresholds[i] */
    double thresholds[MCDP_MAX_THRESHOLDS];
    int32_t n_bins;                        /* 0 = no histogram, else 1..1024 */
    double hist_lo, hist_hi;               /* bin = clamp(floor((delay-lo)/(hi-lo)*n_bins), 0, n_bins-1) */
} mcdp_stats_desc;

typedef struct mcdp_plan mcdp_plan;

/* device ordinal for mcdp_plan_create that only validates and compiles the plan on the host
 * (introspection and error checking without a GPU); every run call on such a plan fails with
 * MCDP_ERR_CUDA -- there is no CPU execution path. */
#define MCDP_DEVICE_NONE (-1)

/* Per-plan options (mcdp_plan_set_option). */
enum {
    MCDP_OPT_STREAM_KEY = 0,    /* uint32 Philox key word 0 (default 0) */
    MCDP_OPT_WARPS_PER_GROUP = 1, /* warps that share one 64-sample group and split each topological level; 0 = auto */
    MCDP_OPT_GROUPS_PER_CTA = 2,  /* 64-sample groups per CTA; 0 = auto */
    MCDP_OPT_HOST_CHUNK = 3,      /* samples per device chunk in the *_host calls; 0 = auto from free HBM */
    MCDP_OPT_RNG_STREAM = 4,      /* 0 = Philox contract (default); 1 = reference-compatible stream: Xoshiro256++
                                     in activity-index order with the libstdc++ transforms (_core.cpp:313-329),
                                     full-output calls only, at most 64 distributions */
    MCDP_OPT_SAMPLES_PER_LANE = 5, /* samples a lane owns in the sweep kernel: 2 = 64-sample groups at 64 registers,
                                     4 = 128-sample groups at 96 registers with 256-bit row accesses (needs 32-byte
                                     aligned buffers, else falls back to 2); 0 = auto.  Results do not depend on it. */
    MCDP_OPT_CLUSTER_SIZE = 6,     /* quad kernel, launches with fewer sample groups than SMs: CTAs per thread-block
                                     cluster that share one group and split its levels (2, 4, 8); 1 = never; 0 = auto.
                                     Results do not depend on it. */
    MCDP_OPT_SMALL_CALL_MAX = 7    /* full-output and injected calls of at most this many samples (run(seed),
                                     _core.cpp:312-353) take the two-kernel path of mcdp_small_sweep.cuh: one thread per
                                     activity draws, one thread per event propagates.  -1 = auto (default), 0 = never.
                                     Results do not depend on it. */
};

const char* mcdp_last_error(void);
int32_t mcdp_abi_version(void);
int32_t mcdp_device_count(void); /* 0 when no CUDA device is usable */

/* Replaces Simulator::Simulator (_core.cpp:193-307): validates (reserved type -1, max_delay >= 0,
 * cycle -> same messages as the reference; additionally index bounds, which the reference leaves
 * undefined), flattens the distributions, levels the DAG, builds the evaluation-ordered
 * event/predecessor stream and uploads it to `device`. */
int32_t mcdp_plan_create(const mcdp_graph_desc* graph, const mcdp_dists_desc* dists, int32_t device, mcdp_plan** out);
void mcdp_plan_destroy(mcdp_plan* plan);
int32_t mcdp_plan_set_option(mcdp_plan* plan, int32_t option, int64_t value);

int32_t mcdp_plan_node_count(const mcdp_plan* plan);     /* Simulator::node_count      _core.cpp:309 */
int32_t mcdp_plan_activity_count(const mcdp_plan* plan); /* Simulator::activity_count  _core.cpp:310 */
int64_t mcdp_plan_pred_count(const mcdp_plan* plan);
int32_t mcdp_plan_level_count(const mcdp_plan* plan);
int32_t mcdp_plan_slot_count(const mcdp_plan* plan);     /* realized scratch rows of the reduced mode */
int32_t mcdp_plan_device(const mcdp_plan* plan);
/* evaluation order (a topological order; event_evaluation_order_ of _core.cpp:177) and level of each position */
int32_t mcdp_plan_get_order(const mcdp_plan* plan, int32_t* order_out, int32_t* level_out);
/* cumulative table the device searches for activity_type (== std::discrete_distribution's _M_cp); returns length or -1 */
int64_t mcdp_plan_get_cumulative(const mcdp_plan* plan, int32_t activity_type, double* cp_out, int64_t cap);

/* The chunk stream the sweep kernel walks (introspection for tests and tooling; layout: csrc/mcdp_records.h,
 * 32-byte units, MCDP_CHUNK_UNITS per chunk).  rows: 0 = event ids (full / injected modes), 1 = recycled scratch
 * slots (reduced modes); dense: 0 = level-aligned chunks (several warps per group), 1 = no level alignment (one warp
 * per group).  Returns the number of chunks; copies min(n_chunks * 512, cap_bytes) bytes to units_out and, when not
 * NULL, level_count + 1 chunk positions (dense: 2) to chunk_level_begin_out.  Works on host-only plans. */
#define MCDP_CHUNK_UNITS 16
int64_t mcdp_plan_get_chunks(const mcdp_plan* plan, int32_t rows, int32_t dense, void* units_out, int64_t cap_bytes,
                             int32_t* chunk_level_begin_out);

/* The launch shape a call over n samples would take (introspection for tests, tooling and the benchmark's
 * kernel label; works on host-only plans, which assume 148 SMs).  out8 = {samples per lane (2 pair kernel, 4 quad
 * kernel), warps per group, groups per CTA, threads per CTA, grid size, dynamic shared memory bytes, CTAs per
 * thread-block cluster (1 = no cluster launch), 1 if the tables are staged in shared memory}. */
int32_t mcdp_plan_launch_shape(const mcdp_plan* plan, int64_t n, int32_t reduced, int32_t n_bins, int64_t* out8);

/* Samples per launch of a reduced / attribution call over n samples: such a call keeps realized[slot][sample] scratch
 * rows (mcdp_plan_slot_count of them) and splits n into launches that fit its scratch budget -- what the plan has
 * already allocated, else 60 % of the free device memory -- sized to whole waves of sample groups. */
int64_t mcdp_plan_reduced_chunk(mcdp_plan* plan, int64_t n, int32_t n_bins, int32_t attribution);

/* ---- device-buffer entry points (asynchronous on `stream`, a cudaStream_t passed as void*) ----
 * Seeds: d_seeds[n] (device) or, when d_seeds is NULL, the arithmetic run seed0, seed0+1, ... */

/* Replaces Simulator::run_many (_core.cpp:355-361) = n x Simulator::run (_core.cpp:312-353):
 * fused Philox sampling + max-plus sweep.  Writes durations[A][ld], realized[E][ld], cause[E][ld]. */
int32_t mcdp_run_full_device(mcdp_plan* plan, const int32_t* d_seeds, int32_t seed0, int64_t n, double* d_realized,
                             double* d_durations, int32_t* d_cause, int64_t ld, void* stream);

/* Propagation phase only (_core.cpp:332-350) with caller-supplied durations[A][ld]:
 * the duration-injection mode (bit-exact fp64 parity tier). */
int32_t mcdp_run_injected_device(mcdp_plan* plan, const double* d_durations, int64_t n, double* d_realized,
                                 int32_t* d_cause, int64_t ld, void* stream);

/* Sampling + sweep + fused statistics; no [.,n] array is written except plan-owned scratch.
 * d_sum[E], d_sumsq[E] (f64), d_late[n_thresholds][E] (u64), d_hist[E][n_bins] (u32); NULL = skip. */
int32_t mcdp_run_reduced_device(mcdp_plan* plan, const int32_t* d_seeds, int32_t seed0, int64_t n,
                                const mcdp_stats_desc* desc, double* d_sum, double* d_sumsq,
                                unsigned long long* d_late, uint32_t* d_hist, void* stream);

/* The same pass plus delay-cause attribution (replaces the np.bincount a caller runs over
 * SimResult.cause_event joined with the precedence list): d_cause_act[A] (u64) counts the samples in which a
 * precedence entry with that activity index decided realized[target] (_core.cpp:343-346: the last entry whose
 * clamped arrival reached the running maximum), d_cause_none[E] (u64) those with cause_event == -1.  Entries
 * whose activity index has no duration row (>= activity_count()) are not counted.  Both accumulate (+=); pass
 * both or neither. */
int32_t mcdp_run_attribution_device(mcdp_plan* plan, const int32_t* d_seeds, int32_t seed0, int64_t n,
                                    const mcdp_stats_desc* desc, double* d_sum, double* d_sumsq,
                                    unsigned long long* d_late, uint32_t* d_hist, unsigned long long* d_cause_act,
                                    unsigned long long* d_cause_none, void* stream);

/* [rows][ld] event-major -> [n][rows] sample-major (and back), for callers that hold device buffers. */
int32_t mcdp_transpose_f64_device(const double* d_in, int64_t rows, int64_t n, int64_t ld, double* d_out, void* stream);
int32_t mcdp_transpose_i32_device(const int32_t* d_in, int64_t rows, int64_t n, int64_t ld, int32_t* d_out, void* stream);

/* ---- host-buffer entry points (synchronous; chunked, copies overlapped with compute) ---- */

/* run_many with host results, sample-major like n SimResult objects: realized[n][E],
 * durations[n][A], cause[n][E]; any output may be NULL.  Host buffers may be pageable or
 * pinned (mcdp_host_alloc). */
int32_t mcdp_run_many_host(mcdp_plan* plan, const int32_t* seeds, int64_t n, double* realized, double* durations,
                           int32_t* cause);
int32_t mcdp_run_injected_host(mcdp_plan* plan, const double* durations, int64_t n, double* realized, int32_t* cause);
int32_t mcdp_run_reduced_host(mcdp_plan* plan, const int32_t* seeds, int64_t n, const mcdp_stats_desc* desc,
                              double* sum, double* sumsq, unsigned long long* late, uint32_t* hist);

int32_t mcdp_run_attribution_host(mcdp_plan* plan, const int32_t* seeds, int64_t n, const mcdp_stats_desc* desc,
                                  double* sum, double* sumsq, unsigned long long* late, uint32_t* hist,
                                  unsigned long long* cause_act, unsigned long long* cause_none);

/* ---- several devices behind one call (additive) ----
 * The reference's only scale-out hook is run_many (_core.cpp:355-361: a loop over independent seeds) plus one
 * Simulator per thread (test/test_simulator.py:201-215).  A plan set compiles the DagContext once and holds one
 * device-resident copy per listed device (a device may be listed more than once); every *_multi call shards its
 * seeds into contiguous blocks (whole 128-sample groups), one per device, and returns exactly what the single-device
 * call returns: identical bits for full outputs and integer statistics, f64 sums to summation order.  Full outputs
 * need no exchange (each device copies its rows into the caller's arrays); the reduced statistics are folded by a
 * reduce-scatter over peer memory (device d sums slice d of every accumulator over all devices through NVLink peer
 * mappings, fixed order) whose slices go straight to the caller -- it needs peer access between the devices. */
typedef struct mcdp_planset mcdp_planset;
int32_t mcdp_planset_create(const mcdp_graph_desc* graph, const mcdp_dists_desc* dists, const int32_t* devices,
                            int32_t n_devices, mcdp_planset** out); /* 1..16 devices */
void mcdp_planset_destroy(mcdp_planset* set);
int32_t mcdp_planset_size(const mcdp_planset* set);
mcdp_plan* mcdp_planset_plan(mcdp_planset* set, int32_t i); /* borrowed: plan of the i-th listed device */
int32_t mcdp_planset_set_option(mcdp_planset* set, int32_t option, int64_t value); /* on every plan of the set */
int32_t mcdp_run_many_host_multi(mcdp_planset* set, const int32_t* seeds, int64_t n, double* realized, double* durations,
                                 int32_t* cause);
int32_t mcdp_run_injected_host_multi(mcdp_planset* set, const double* durations, int64_t n, double* realized, int32_t* cause);
int32_t mcdp_run_reduced_host_multi(mcdp_planset* set, const int32_t* seeds, int64_t n, const mcdp_stats_desc* desc,
                                    double* sum, double* sumsq, unsigned long long* late, uint32_t* hist);
int32_t mcdp_run_attribution_host_multi(mcdp_planset* set, const int32_t* seeds, int64_t n, const mcdp_stats_desc* desc,
                                        double* sum, double* sumsq, unsigned long long* late, uint32_t* hist,
                                        unsigned long long* cause_act, unsigned long long* cause_none);

/* ---- analytic (PMF) propagation (additive) ----
 * Replaces the pure-numpy engine of the reference: DiscretePMF.convolve / maximum (analytic/_pmf.py:107-148),
 * AnalyticPropagator.run (analytic/_propagator.py:89-148) and its bound handling _convert_to_simulated_event
 * (:158-265).  All times are integer seconds on the grid of `step`; PMF i has its first value at pmf_start[i] and the
 * probabilities pmf_probs[pmf_off[i] .. pmf_off[i+1]) at spacing `step`. */
typedef struct {
    int32_t n_events;
    const int64_t* lower;   /* [E] int(round(earliest))                                   _propagator.py:150-156 */
    const int64_t* upper;   /* [E] int(round(min(latest, earliest + max_delay)))                                 */
    const int64_t* origin;  /* [E] value of an event without predecessors: round(earliest / step) * step  :104-106 */
    int64_t step;
    int32_t n_prec_entries; /* precedence list as in mcdp_graph_desc; pred_pmf[k] = PMF index of that edge */
    const int32_t* prec_target;
    const int64_t* prec_off;
    const int32_t* pred_src;
    const int32_t* pred_pmf;
    int32_t n_pmfs;
    const int64_t* pmf_start;
    const int64_t* pmf_off; /* [n_pmfs + 1] */
    const double* pmf_probs;
    int32_t underflow_rule, overflow_rule; /* 1 truncate, 2 remove, 3 redistribute (analytic/_context.py:42-66) */
} mcdp_analytic_desc;

/* Bins the result arrays must hold; out_off (may be NULL) receives the [E + 1] slot offsets of the events. */
int64_t mcdp_analytic_out_capacity(const mcdp_analytic_desc* desc, int64_t* out_off);
/* One launch per topological level, one CTA per event.  Event e's PMF: first value out_start[e], out_len[e] bins at
 * out_probs[out_off[e] ..]; underflow[e] / overflow[e] as SimulatedEvent reports them.  Failures of the reference's
 * checks (cycle, no bound bin to truncate onto, empty PMF) return MCDP_ERR_INVALID with its message. */
int32_t mcdp_analytic_run(const mcdp_analytic_desc* desc, int32_t device, int64_t* out_start, int32_t* out_len, int64_t* out_off,
                          double* out_probs, int64_t out_cap, double* underflow, double* overflow);

/* Phases of this thread's last mcdp_analytic_run, in milliseconds: out5[0] host preparation (precedence by event,
 * levels, output slots), [1] device allocations and uploads, [2] the level kernels (all launches, synchronised),
 * [3] results to the host; out5[4] = number of levels (= kernel launches).  Measurement aid of scripts/bench_analytic.py;
 * the reference has no counterpart. */
int32_t mcdp_analytic_last_profile(double* out5);
/* A single PMF operation on the device: op 0 = convolve(a, b), 1 = maximum(a, b), 2 = clip a to [min_value, max_value]
 * with the flow rules (b unused). */
int32_t mcdp_pmf_op(int32_t op, int32_t device, int64_t step, int64_t a_start, int32_t a_len, const double* a_probs, int64_t b_start,
                    int32_t b_len, const double* b_probs, int64_t min_value, int64_t max_value, int32_t underflow_rule,
                    int32_t overflow_rule, int64_t* out_start, int32_t* out_len, double* out_probs, int64_t out_cap,
                    double* out_underflow, double* out_overflow);

void* mcdp_host_alloc(size_t bytes); /* pinned host memory, NULL on failure */
void mcdp_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MCDP_B200_H */
