/* oracle/mcdp_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the Monte-Carlo hot path of WonJayne/mc_dagprop
 * (reference: src/mc_dagprop/monte_carlo/_core.cpp, _custom_rng.hpp, and the
 * libstdc++ 13 <random> transforms the reference delegates to).  Every function
 * cites the reference file:line it restates.  Parity status: PINNED -- checked
 * bit-for-bit against the reference's own golden vectors (tests/test_oracle_pinned.py)
 * and against the unmodified reference compiled into oracle/_ref/ (same tests,
 * plus tests/golden/ fixtures generated from it by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load this library.  The product (mc_dagprop_b200) never does.
 *
 * The second half ("spec" functions) restates the *device* generator contract
 * mcdp-philox-v2 so that the CUDA kernels can be checked
 * draw-for-draw; it is not reference behaviour and is labelled as such.
 */
#ifndef MCDP_ORACLE_H
#define MCDP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    MCDP_OR_CONSTANT = 0,    /* p0 = factor                         (_core.cpp:72-76)   */
    MCDP_OR_EXPONENTIAL = 1, /* p0 = lambda (mean), p1 = max_scale  (_core.cpp:78-90)   */
    MCDP_OR_GAMMA = 2,       /* p0 = shape, p1 = scale, p2 = max    (_core.cpp:92-105)  */
    MCDP_OR_EMP_ABS = 3,     /* table values + weights              (_core.cpp:110-126) */
    MCDP_OR_EMP_REL = 4      /* table factors + weights             (_core.cpp:129-141) */
};

/* ---- reference RNG (vendored header of the reference) ------------------- */
typedef struct {
    uint64_t s[4];
} mcdp_or_xoshiro;
uint64_t mcdp_or_splitmix64_next(uint64_t* state);                 /* _custom_rng.hpp:533-538 */
void mcdp_or_xoshiro_seed(mcdp_or_xoshiro* g, uint64_t seed);      /* _custom_rng.hpp:570-576 */
uint64_t mcdp_or_xoshiro_next(mcdp_or_xoshiro* g);                 /* _custom_rng.hpp:589-599 */
double mcdp_or_canonical(mcdp_or_xoshiro* g);                      /* random.tcc:3346-3381    */

/* ---- reference simulator ------------------------------------------------ */
typedef struct mcdp_or_sim mcdp_or_sim;

/* Same flat description as oracle/ref_driver.cpp::mcdp_ref_create.  Returns NULL and
 * fills err on the reference's RuntimeError conditions (_core.cpp:196-201,262-264,119,134). */
mcdp_or_sim* mcdp_or_sim_create(int32_t n_events, const double* earliest, int32_t n_act_entries,
                                const int32_t* act_idx, const double* act_base, const int32_t* act_type,
                                int32_t n_prec_entries, const int32_t* prec_target, const int64_t* prec_off,
                                const int32_t* pred_src, const int32_t* pred_act, double max_delay,
                                int32_t n_dists, const int32_t* dist_type, const int32_t* dist_kind,
                                const double* p0, const double* p1, const double* p2, const int64_t* tab_off,
                                const double* tab_values, const double* tab_weights, char* err, size_t errlen);
void mcdp_or_sim_destroy(mcdp_or_sim* sim);
int32_t mcdp_or_sim_node_count(const mcdp_or_sim* sim);      /* _core.cpp:309 */
int32_t mcdp_or_sim_activity_count(const mcdp_or_sim* sim);  /* _core.cpp:310 */
/* topological order + CSR exactly as the reference builds them (_core.cpp:248-285) */
void mcdp_or_sim_get_order(const mcdp_or_sim* sim, int32_t* order_out);
int64_t mcdp_or_sim_pred_count(const mcdp_or_sim* sim);
void mcdp_or_sim_get_csr(const mcdp_or_sim* sim, int64_t* off_out, int32_t* src_out, int32_t* act_out);
/* cumulative table of dist t as std::discrete_distribution builds it (random.tcc:2655-2678) */
int64_t mcdp_or_sim_get_cp(const mcdp_or_sim* sim, int32_t dist_type, double* cp_out, int64_t cap);

/* Simulator::run for every seed (_core.cpp:312-353 looped as in :355-361), Xoshiro256++
 * stream in activity-index order.  Sample-major outputs [n][E], [n][A], [n][E]; NULL = skip. */
int32_t mcdp_or_sim_run_many(mcdp_or_sim* sim, const int32_t* seeds, int64_t n, double* realized,
                             double* durations, int32_t* cause);

/* Propagation phase only (_core.cpp:332-350) with caller-supplied durations [n][A]. */
int32_t mcdp_or_sim_run_injected(mcdp_or_sim* sim, const double* durations, int64_t n, double* realized,
                                 int32_t* cause);

/* ---- device generator contract mcdp-philox-v2 (NOT reference behaviour) - */
void mcdp_or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* durations per the device contract, then the reference propagation. */
int32_t mcdp_or_sim_run_many_spec(mcdp_or_sim* sim, const int32_t* seeds, int64_t n, uint32_t stream_key,
                                  double* realized, double* durations, int32_t* cause);

#ifdef __cplusplus
}
#endif
#endif
