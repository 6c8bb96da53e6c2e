/* oracle/mcdp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See mcdp_oracle.h.
 *
 * Part 1 restates the reference (file:line cited per function; "ref:" = /root/reference/
 * src/mc_dagprop/monte_carlo/, "libstdc++:" = /usr/include/c++/13/bits/).  Parity PINNED by
 * tests/test_oracle_pinned.py.  Part 2 restates the device generator contract.
 * Built with -ffp-contract=off: the reference wheel targets baseline x86-64 (no FMA),
 * so every a*b+c below is a rounded multiply followed by a rounded add.
 */
#include "mcdp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/* Part 1a: RNG                                                              */
/* ======================================================================== */

/* ref: _custom_rng.hpp:533-538 (SplitMix64::operator()) */
uint64_t mcdp_or_splitmix64_next(uint64_t* state) {
    uint64_t z = (*state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* ref: _custom_rng.hpp:570-576 (Xoshiro256PP::seed(result_type)) */
void mcdp_or_xoshiro_seed(mcdp_or_xoshiro* g, uint64_t seed) {
    uint64_t sm = seed;
    for (int i = 0; i < 4; ++i) g->s[i] = mcdp_or_splitmix64_next(&sm);
}

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

/* ref: _custom_rng.hpp:589-599 (Xoshiro256PP::operator()) */
uint64_t mcdp_or_xoshiro_next(mcdp_or_xoshiro* g) {
    uint64_t* s = g->s;
    const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

/* libstdc++: random.tcc:3346-3381 generate_canonical<double,53> with a 64-bit URBG:
 * one call, sum = double(u64) (round-to-nearest), divided by 2^64, clamped below 1. */
double mcdp_or_canonical(mcdp_or_xoshiro* g) {
    double sum = (double)mcdp_or_xoshiro_next(g);
    double ret = sum / 18446744073709551616.0;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}

/* ======================================================================== */
/* Part 1b: distributions                                                    */
/* ======================================================================== */

typedef struct {
    int kind;
    double p0, p1, p2;
    int64_t n;        /* table length */
    double* vals;     /* values / factors */
    double* cp;       /* cumulative probabilities, cp_n == 0 when n < 2 */
    int64_t cp_n;
    /* std::normal_distribution cached second variate (random.tcc:1820-1841); it
     * lives in the per-type distribution object and survives rng_.seed(). */
    int saved_available;
    double saved;
    double malpha, a2; /* gamma param_type::_M_initialize, random.tcc:2336-2345 */
    /* device-contract constants (Part 2) */
    double exp_F;
} or_dist;

/* libstdc++: random.tcc:2655-2678 discrete_distribution::param_type::_M_initialize */
static void or_discrete_init(or_dist* d, const double* weights) {
    d->cp = NULL;
    d->cp_n = 0;
    if (d->n < 2) return; /* _M_prob.clear(): sampling returns 0 without a draw */
    double sum = 0.0;
    for (int64_t i = 0; i < d->n; ++i) sum += weights[i];
    d->cp = (double*)malloc(sizeof(double) * (size_t)d->n);
    double acc = 0.0;
    for (int64_t i = 0; i < d->n; ++i) {
        double p = weights[i] / sum; /* __normalize */
        acc = (i == 0) ? p : acc + p; /* std::partial_sum */
        d->cp[i] = acc;
    }
    d->cp[d->n - 1] = 1.0;
    d->cp_n = d->n;
}

/* std::lower_bound(cp, p): first index with cp[i] >= p (random.tcc:2709-2713) */
static int64_t or_lower_bound(const double* cp, int64_t n, double p) {
    int64_t lo = 0, len = n;
    while (len > 0) {
        int64_t half = len >> 1;
        if (cp[lo + half] < p) {
            lo += half + 1;
            len -= half + 1;
        } else {
            len = half;
        }
    }
    return lo;
}

/* libstdc++: random.tcc:2696-2714 discrete_distribution::operator() */
static int64_t or_discrete_draw(or_dist* d, mcdp_or_xoshiro* g) {
    if (d->cp_n == 0) return 0;
    double p = mcdp_or_canonical(g);
    return or_lower_bound(d->cp, d->cp_n, p);
}

/* libstdc++: random.tcc:1809-1844 normal_distribution::operator() (Marsaglia polar, cached) */
static double or_normal(or_dist* d, mcdp_or_xoshiro* g) {
    double ret;
    if (d->saved_available) {
        d->saved_available = 0;
        ret = d->saved;
    } else {
        double x, y, r2;
        do {
            x = 2.0 * mcdp_or_canonical(g) - 1.0;
            y = 2.0 * mcdp_or_canonical(g) - 1.0;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        const double mult = sqrt(-2 * log(r2) / r2);
        d->saved = x * mult;
        d->saved_available = 1;
        ret = y * mult;
    }
    ret = ret * 1.0 + 0.0; /* stddev 1, mean 0 */
    return ret;
}

/* libstdc++: random.tcc:2352-2393 gamma_distribution::operator() (Marsaglia-Tsang) */
static double or_gamma_draw(or_dist* d, mcdp_or_xoshiro* g) {
    const double alpha = d->p0, beta = d->p1;
    double u, v, n;
    const double a1 = d->malpha - 1.0 / 3.0;
    do {
        do {
            n = or_normal(d, g);
            v = 1.0 + d->a2 * n;
        } while (v <= 0.0);
        v = v * v * v;
        u = mcdp_or_canonical(g);
    } while (u > 1.0 - 0.0331 * n * n * n * n && (log(u) > (0.5 * n * n + a1 * (1.0 - v + log(v)))));
    if (alpha == d->malpha) return a1 * v * beta;
    do u = mcdp_or_canonical(g);
    while (u == 0.0);
    return pow(u, 1.0 / alpha) * a1 * v * beta;
}

/* ref: _core.cpp:72-141 the five Dist::sample(rng, base) bodies; returns the EXTRA delay */
static double or_sample_extra(or_dist* d, mcdp_or_xoshiro* g, double base) {
    switch (d->kind) {
        case MCDP_OR_CONSTANT: /* _core.cpp:75 */
            return base * d->p0;
        case MCDP_OR_EXPONENTIAL: { /* _core.cpp:83-89; random.h:4897-4905 with lambda()=1/lam */
            const double rate = 1.0 / d->p0;
            double x;
            do {
                x = -log(1.0 - mcdp_or_canonical(g)) / rate;
            } while (x > d->p1);
            return x * base;
        }
        case MCDP_OR_GAMMA: { /* _core.cpp:98-104 */
            double x;
            do {
                x = or_gamma_draw(d, g);
            } while (x > d->p2);
            return x * base;
        }
        case MCDP_OR_EMP_ABS: /* _core.cpp:125 */
            return d->vals[or_discrete_draw(d, g)];
        case MCDP_OR_EMP_REL: /* _core.cpp:140 */
            return d->vals[or_discrete_draw(d, g)] * base;
    }
    return 0.0;
}

/* ======================================================================== */
/* Part 1c: Simulator                                                        */
/* ======================================================================== */

struct mcdp_or_sim {
    int32_t E, A;
    double max_delay;
    double* earliest;
    double* base;      /* activities_[link].duration, 0.0 for idx gaps (_core.cpp:218) */
    int32_t* act_dist; /* activity_to_dist_index_, -1 = no delay (_core.cpp:219-228) */
    int32_t n_dists;
    or_dist* dists;
    int32_t* dist_types;
    int32_t* order; /* event_evaluation_order_ */
    int64_t* off;   /* predecessor_offsets_ */
    int32_t* src;   /* flat_predecessor_sources_ */
    int32_t* act;   /* flat_predecessor_edges_ */
    mcdp_or_xoshiro rng;
    double* realized_scratch;
    double* dur_scratch;
    int32_t* cause_scratch;
};

static void set_err(char* err, size_t errlen, const char* msg) {
    if (err && errlen) {
        strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

void mcdp_or_sim_destroy(mcdp_or_sim* s) {
    if (!s) return;
    free(s->earliest);
    free(s->base);
    free(s->act_dist);
    if (s->dists) {
        for (int i = 0; i < s->n_dists; ++i) {
            free(s->dists[i].vals);
            free(s->dists[i].cp);
        }
    }
    free(s->dists);
    free(s->dist_types);
    free(s->order);
    free(s->off);
    free(s->src);
    free(s->act);
    free(s->realized_scratch);
    free(s->dur_scratch);
    free(s->cause_scratch);
    free(s);
}

/* ref: _core.cpp:193-307 Simulator::Simulator */
mcdp_or_sim* mcdp_or_sim_create(int32_t n_events, const double* earliest, int32_t n_act_entries,
                                const int32_t* act_idx, const double* act_base, const int32_t* act_type,
                                int32_t n_prec_entries, const int32_t* prec_target, const int64_t* prec_off,
                                const int32_t* pred_src, const int32_t* pred_act, double max_delay,
                                int32_t n_dists, const int32_t* dist_type, const int32_t* dist_kind,
                                const double* p0, const double* p1, const double* p2, const int64_t* tab_off,
                                const double* tab_values, const double* tab_weights, char* err, size_t errlen) {
    /* _core.cpp:196-201 */
    for (int t = 0; t < n_dists; ++t) {
        if (dist_type[t] == -1) {
            set_err(err, errlen, "Activity type -1 is reserved for no delay");
            return NULL;
        }
    }
    if (max_delay < 0.0) {
        set_err(err, errlen, "max_delay must be non-negative");
        return NULL;
    }
    mcdp_or_sim* s = (mcdp_or_sim*)calloc(1, sizeof(*s));
    s->E = n_events;
    s->max_delay = max_delay;
    s->earliest = (double*)malloc(sizeof(double) * (size_t)(n_events > 0 ? n_events : 1));
    memcpy(s->earliest, earliest, sizeof(double) * (size_t)n_events);

    /* _core.cpp:204-210: flatten distributions (later add_* for a type overwrite earlier ones) */
    s->n_dists = n_dists;
    s->dists = (or_dist*)calloc((size_t)(n_dists > 0 ? n_dists : 1), sizeof(or_dist));
    s->dist_types = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_dists > 0 ? n_dists : 1));
    for (int t = 0; t < n_dists; ++t) {
        or_dist* d = &s->dists[t];
        s->dist_types[t] = dist_type[t];
        d->kind = dist_kind[t];
        d->p0 = p0[t];
        d->p1 = p1[t];
        d->p2 = p2[t];
        if (d->kind == MCDP_OR_EMP_ABS || d->kind == MCDP_OR_EMP_REL) {
            d->n = tab_off[t + 1] - tab_off[t];
            d->vals = (double*)malloc(sizeof(double) * (size_t)(d->n > 0 ? d->n : 1));
            memcpy(d->vals, tab_values + tab_off[t], sizeof(double) * (size_t)d->n);
            or_discrete_init(d, tab_weights + tab_off[t]);
        }
        if (d->kind == MCDP_OR_GAMMA) {
            d->malpha = d->p0 < 1.0 ? d->p0 + 1.0 : d->p0; /* random.tcc:2339 */
            const double a1 = d->malpha - 1.0 / 3.0;
            d->a2 = 1.0 / sqrt(9.0 * a1);
        }
        if (d->kind == MCDP_OR_EXPONENTIAL) {
            d->exp_F = isinf(d->p1) ? 1.0 : -expm1(-d->p1 / d->p0);
        }
    }

    /* _core.cpp:213-229 */
    int32_t max_idx = -1;
    for (int i = 0; i < n_act_entries; ++i)
        if (act_idx[i] > max_idx) max_idx = act_idx[i];
    s->A = max_idx + 1;
    const size_t An = (size_t)(s->A > 0 ? s->A : 1);
    s->base = (double*)calloc(An, sizeof(double));
    s->act_dist = (int32_t*)malloc(sizeof(int32_t) * An);
    for (int i = 0; i < s->A; ++i) s->act_dist[i] = -1;
    for (int i = 0; i < n_act_entries; ++i) {
        const int32_t link = act_idx[i];
        s->base[link] = act_base[i];
        s->act_dist[link] = -1;
        /* last matching dist entry wins == unordered_map overwrite semantics of add_* */
        for (int t = 0; t < n_dists; ++t)
            if (dist_type[t] == act_type[i]) s->act_dist[link] = t;
    }

    /* _core.cpp:232-245: preds_by_target (last entry for a target wins), indegree, adjacency */
    const int32_t E = n_events;
    const size_t En = (size_t)(E > 0 ? E : 1);
    int32_t* entry_of = (int32_t*)malloc(sizeof(int32_t) * En);
    int32_t* indeg = (int32_t*)calloc(En, sizeof(int32_t));
    int64_t* adj_cnt = (int64_t*)calloc(En + 1, sizeof(int64_t));
    for (int i = 0; i < E; ++i) entry_of[i] = -1;
    for (int i = 0; i < n_prec_entries; ++i) {
        entry_of[prec_target[i]] = i;
        indeg[prec_target[i]] = (int32_t)(prec_off[i + 1] - prec_off[i]);
        for (int64_t k = prec_off[i]; k < prec_off[i + 1]; ++k) adj_cnt[pred_src[k] + 1]++;
    }
    for (int i = 0; i < E; ++i) adj_cnt[i + 1] += adj_cnt[i];
    const int64_t n_adj = adj_cnt[E];
    int32_t* adj = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_adj > 0 ? n_adj : 1));
    int64_t* adj_pos = (int64_t*)malloc(sizeof(int64_t) * En);
    for (int i = 0; i < E; ++i) adj_pos[i] = adj_cnt[i];
    for (int i = 0; i < n_prec_entries; ++i)
        for (int64_t k = prec_off[i]; k < prec_off[i + 1]; ++k) adj[adj_pos[pred_src[k]]++] = prec_target[i];

    /* _core.cpp:248-264: Kahn, FIFO queue, roots in ascending index */
    s->order = (int32_t*)malloc(sizeof(int32_t) * En);
    int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * En);
    int64_t qh = 0, qt = 0, n_order = 0;
    for (int i = 0; i < E; ++i)
        if (indeg[i] == 0) queue[qt++] = i;
    while (qh < qt) {
        const int32_t n = queue[qh++];
        s->order[n_order++] = n;
        for (int64_t k = adj_cnt[n]; k < adj_cnt[n + 1]; ++k) {
            const int32_t dst = adj[k];
            /* a doubly listed target can reach 0 only once per decrement chain; the queue
             * cannot overflow E because the reference pushes a node when --indegree == 0 */
            if (--indeg[dst] == 0 && qt < E) queue[qt++] = dst;
        }
    }
    int cyc = (n_order != E);
    if (!cyc) {
        /* _core.cpp:267-285: CSR by event id, preds in the caller's order */
        s->off = (int64_t*)calloc(En + 1, sizeof(int64_t));
        for (int i = 0; i < E; ++i) {
            const int32_t en = entry_of[i];
            s->off[i + 1] = s->off[i] + (en < 0 ? 0 : prec_off[en + 1] - prec_off[en]);
        }
        const int64_t P = s->off[E];
        s->src = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P > 0 ? P : 1));
        s->act = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P > 0 ? P : 1));
        for (int i = 0; i < E; ++i) {
            const int32_t en = entry_of[i];
            if (en < 0) continue;
            int64_t w = s->off[i];
            for (int64_t k = prec_off[en]; k < prec_off[en + 1]; ++k, ++w) {
                s->src[w] = pred_src[k];
                s->act[w] = pred_act[k];
            }
        }
    }
    free(entry_of);
    free(indeg);
    free(adj_cnt);
    free(adj);
    free(adj_pos);
    free(queue);
    if (cyc) {
        set_err(err, errlen, "Invalid DAG: cycle detected in precedence list");
        mcdp_or_sim_destroy(s);
        return NULL;
    }
    /* _core.cpp:298-306 scratch; no-dist links keep their base duration */
    s->realized_scratch = (double*)malloc(sizeof(double) * En);
    s->cause_scratch = (int32_t*)malloc(sizeof(int32_t) * En);
    s->dur_scratch = (double*)malloc(sizeof(double) * An);
    for (int i = 0; i < s->A; ++i) s->dur_scratch[i] = s->base[i];
    return s;
}

int32_t mcdp_or_sim_node_count(const mcdp_or_sim* s) { return s->E; }
int32_t mcdp_or_sim_activity_count(const mcdp_or_sim* s) { return s->A; }
void mcdp_or_sim_get_order(const mcdp_or_sim* s, int32_t* out) { memcpy(out, s->order, sizeof(int32_t) * (size_t)s->E); }
int64_t mcdp_or_sim_pred_count(const mcdp_or_sim* s) { return s->off[s->E]; }
void mcdp_or_sim_get_csr(const mcdp_or_sim* s, int64_t* off, int32_t* src, int32_t* act) {
    memcpy(off, s->off, sizeof(int64_t) * ((size_t)s->E + 1));
    memcpy(src, s->src, sizeof(int32_t) * (size_t)s->off[s->E]);
    memcpy(act, s->act, sizeof(int32_t) * (size_t)s->off[s->E]);
}
int64_t mcdp_or_sim_get_cp(const mcdp_or_sim* s, int32_t dist_type, double* cp_out, int64_t cap) {
    for (int t = s->n_dists - 1; t >= 0; --t) {
        if (s->dist_types[t] != dist_type) continue;
        const int64_t n = s->dists[t].cp_n;
        for (int64_t i = 0; i < n && i < cap; ++i) cp_out[i] = s->dists[t].cp[i];
        return n;
    }
    return -1;
}

/* ref: _core.cpp:332-350 the propagation sweep.  std::min(a,b) == (b < a) ? b : a. */
static void or_propagate(const mcdp_or_sim* s, const double* dur, double* realized, int32_t* cause) {
    for (int i = 0; i < s->E; ++i) realized[i] = s->earliest[i]; /* _core.cpp:318-320 */
    for (int oi = 0; oi < s->E; ++oi) {
        const int32_t e = s->order[oi];
        const double earliest = s->earliest[e];
        const double ub = earliest + s->max_delay;
        double latest = realized[e];
        int32_t c = -1;
        for (int64_t k = s->off[e]; k < s->off[e + 1]; ++k) {
            const int32_t src = s->src[k];
            /* An activity index >= activity_count() reads past actual_durations_ in the reference
             * (undefined behaviour; its own LargeScaleTest, test_simulator.py:176-199, depends on
             * that read yielding 0.0).  Defined here as a zero-duration link. */
            const int32_t a = s->act[k];
            double t = realized[src] + (a < s->A ? dur[a] : 0.0);
            t = (ub < t) ? ub : t;
            if (t >= latest) {
                latest = t;
                c = src;
            }
        }
        realized[e] = (ub < latest) ? ub : latest;
        cause[e] = c;
    }
}

static void or_emit(const mcdp_or_sim* s, int64_t i, double* realized, double* durations, int32_t* cause) {
    if (realized) memcpy(realized + i * s->E, s->realized_scratch, sizeof(double) * (size_t)s->E);
    if (durations) memcpy(durations + i * s->A, s->dur_scratch, sizeof(double) * (size_t)s->A);
    if (cause) memcpy(cause + i * s->E, s->cause_scratch, sizeof(int32_t) * (size_t)s->E);
}

/* ref: _core.cpp:312-353 Simulator::run, looped as run_many does (_core.cpp:355-361) */
int32_t mcdp_or_sim_run_many(mcdp_or_sim* s, const int32_t* seeds, int64_t n, double* realized, double* durations,
                             int32_t* cause) {
    for (int64_t i = 0; i < n; ++i) {
        mcdp_or_xoshiro_seed(&s->rng, (uint64_t)(int64_t)seeds[i]); /* int -> uint64 sign-extends */
        for (int link = 0; link < s->A; ++link) {                   /* _core.cpp:323-329 */
            const int32_t di = s->act_dist[link];
            if (di < 0) continue;
            const double b = s->base[link];
            const double extra = or_sample_extra(&s->dists[di], &s->rng, b);
            s->dur_scratch[link] = b + extra;
        }
        or_propagate(s, s->dur_scratch, s->realized_scratch, s->cause_scratch);
        or_emit(s, i, realized, durations, cause);
    }
    return 0;
}

int32_t mcdp_or_sim_run_injected(mcdp_or_sim* s, const double* durations, int64_t n, double* realized, int32_t* cause) {
    for (int64_t i = 0; i < n; ++i) {
        or_propagate(s, durations + i * s->A, s->realized_scratch, s->cause_scratch);
        or_emit(s, i, realized, NULL, cause);
    }
    return 0;
}

/* ======================================================================== */
/* Part 2: device generator contract "mcdp-philox-v2"   */
/* NOT reference behaviour: the reference stream is Xoshiro256++ (Part 1).   */
/* ======================================================================== */

/* Philox4x32-10, Salmon et al. SC'11 (same constants as curand_philox4x32_x.h) */
void mcdp_or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define SPEC_TAG_PAIR 0x50414952u /* 'PAIR' */
#define SPEC_TAG_SOLO 0x534F4C4Fu /* 'SOLO' */
#define SPEC_KEY1 0x4D434450u     /* 'MCDP' */
#define SPEC_TAG_ERLANG 0x45524C47u /* 'ERLG' */
#define SPEC_TAG_QUAD 0x51554144u /* 'QUAD' */
#define SPEC_TAG_GAM0 0x47414D30u /* 'GAM0' */
#define SPEC_TAG_GBST 0x47425354u /* 'GBST' */
#define SPEC_GAMMA_MAX_ATTEMPTS 65536u
/* width rules of v2: a 32-bit uniform from a QUAD block for tables of at most 4096 entries (64 bits from a PAIR
 * block beyond); exponentials draw one 32-bit word w, refined by a second word when w lies in the top 2^-20 */
#define SPEC_QUAD_TABLE_MAX_LEN 4096
#define SPEC_EXP_TAIL_WORD 0xFFFFF000u

static double spec_u52(uint64_t x) { return ((double)(x >> 12) + 0.5) * 0x1p-52; }
static double spec_u32(uint32_t w) { return ((double)w + 0.5) * 0x1p-32; }
static double spec_u23(uint32_t w) { return ((double)(w >> 9) + 0.5) * 0x1p-23; }

/* 64 random bits of (activity, seed, draw j): one Philox block serves the seed pair {2k, 2k+1} */
static uint64_t spec_pair_bits(uint32_t act, uint32_t seed, uint32_t j, uint32_t stream_key) {
    const uint32_t ctr[4] = {seed >> 1, act, j, SPEC_TAG_PAIR};
    const uint32_t key[2] = {stream_key, SPEC_KEY1};
    uint32_t r[4];
    mcdp_or_philox4x32_10(ctr, key, r);
    return (seed & 1u) ? (((uint64_t)r[3] << 32) | r[2]) : (((uint64_t)r[1] << 32) | r[0]);
}

/* 32 random bits of (activity, seed, draw j): word (seed & 3) of the block of the seed quad {4k .. 4k+3} */
static uint32_t spec_quad_word(uint32_t act, uint32_t seed, uint32_t j, uint32_t tag, uint32_t stream_key) {
    const uint32_t ctr[4] = {seed >> 2, act, j, tag};
    const uint32_t key[2] = {stream_key, SPEC_KEY1};
    uint32_t r[4];
    mcdp_or_philox4x32_10(ctr, key, r);
    return r[seed & 3u];
}

static double spec_sample_extra(const or_dist* d, double base, uint32_t act, uint32_t seed, uint32_t stream_key) {
    switch (d->kind) {
        case MCDP_OR_CONSTANT:
            return base * d->p0;
        case MCDP_OR_EXPONENTIAL: {
            /* inverse CDF of the law truncated to [0, max_scale] == law of the reference's rejection loop */
            const uint32_t w = spec_quad_word(act, seed, 0u, SPEC_TAG_QUAD, stream_key);
            double x;
            if (w >= SPEC_EXP_TAIL_WORD && d->exp_F >= 0x1p-10) {
                /* far tail: 1 - u = ((2^32 - w) - (w' + 1/2) 2^-32) 2^-32 with the second word w' (draw j = 1) */
                const uint32_t w2 = spec_quad_word(act, seed, 1u, SPEC_TAG_QUAD, stream_key);
                const double one_minus_u = ((double)(uint32_t)(0u - w) - spec_u32(w2)) * 0x1p-32;
                const double one_minus_F = isinf(d->p1) ? 0.0 : exp(-d->p1 / d->p0);
                x = -d->p0 * log(fma(d->exp_F, one_minus_u, one_minus_F));
            } else {
                x = -d->p0 * log1p(-spec_u32(w) * d->exp_F);
            }
            if (x > d->p1) x = d->p1;
            return x * base;
        }
        case MCDP_OR_EMP_ABS:
        case MCDP_OR_EMP_REL: {
            int64_t idx = 0;
            if (d->cp_n) {
                const double u = d->cp_n <= SPEC_QUAD_TABLE_MAX_LEN
                                     ? spec_u32(spec_quad_word(act, seed, 0u, SPEC_TAG_QUAD, stream_key))
                                     : spec_u52(spec_pair_bits(act, seed, 0u, stream_key));
                idx = or_lower_bound(d->cp, d->cp_n, u);
            }
            return d->kind == MCDP_OR_EMP_ABS ? d->vals[idx] : d->vals[idx] * base;
        }
        case MCDP_OR_GAMMA: {
            const double a1 = d->malpha - 1.0 / 3.0;
            const uint32_t key[2] = {stream_key, SPEC_KEY1};
            double x = 0.0;
            const double twice = 2.0 * d->p0;
            if (twice == floor(twice) && twice >= 1.0 && twice <= 8.0 && twice != 7.0) {
                /* exact transformation for 2*shape in {1..6, 8}: k = floor(shape) unit exponentials (+ half a squared
                 * Box-Muller normal); one uniform (shape 1): the seed's word of block (seed>>2, act, j, 'ERLG'); two:
                 * the seed's half of block (seed>>1, act, j, 'ERLG'); else the four words of block (seed, act, j,
                 * 'ERLG'); draw j+1 when x > max_scale */
                const int k = (int)floor(d->p0), half = ((int)twice) & 1;
                for (uint32_t j = 0; j < SPEC_GAMMA_MAX_ATTEMPTS; ++j) {
                    uint32_t w[4], r[4];
                    if (k + 2 * half == 1) {
                        w[0] = spec_quad_word(act, seed, j, SPEC_TAG_ERLANG, stream_key);
                        w[1] = w[2] = w[3] = 0u;
                    } else if (k + 2 * half <= 2) {
                        const uint32_t ctr[4] = {seed >> 1, act, j, SPEC_TAG_ERLANG};
                        mcdp_or_philox4x32_10(ctr, key, r);
                        w[0] = (seed & 1u) ? r[2] : r[0];
                        w[1] = (seed & 1u) ? r[3] : r[1];
                        w[2] = w[3] = 0u;
                    } else {
                        const uint32_t ctr[4] = {seed, act, j, SPEC_TAG_ERLANG};
                        mcdp_or_philox4x32_10(ctr, key, w);
                    }
                    double e = 0.0;
                    if (k > 0) {
                        double prod = spec_u32(w[0]);
                        for (int i = 1; i < k; ++i) prod *= spec_u32(w[i]);
                        e = -log(prod);
                    }
                    if (half) {
                        const double c = cos(6.283185307179586476925286766559 * (double)(int32_t)w[k + 1] * 0x1p-32);
                        e += -log(spec_u32(w[k])) * c * c;
                    }
                    x = d->p1 * e;
                    if (x <= d->p2) return x * base;
                }
                if (x > d->p2) x = d->p2;
                return x * base;
            }
            for (uint32_t t = 0; t < SPEC_GAMMA_MAX_ATTEMPTS; ++t) {
                /* Box-Muller normal: radius from a 23-bit uniform, angle 2 pi * int32(w) / 2^32.  The device
                 * evaluates this deviate and the two acceptance comparisons with fp32 hardware
                 * approximations; this restatement is the exact-arithmetic definition.
                 * Attempt 0: block (seed>>1, act, 0, 'GAM0') holds ONE Box-Muller pair for the seed pair -- radius
                 * word 0, angle word 1; the even seed takes the cos branch and accept word 2, the odd seed the sin
                 * branch and accept word 3 -- and the shape < 1 boost uniform is the seed's word of block
                 * (seed>>2, act, 0, 'GBST').  Attempt t >= 1: block (seed, act, t, 'SOLO'): radius, angle (cos
                 * branch), accept, boost. */
                uint32_t w[4];
                double n, u, boost_u;
                if (t == 0u) {
                    const uint32_t ctr[4] = {seed >> 1, act, 0u, SPEC_TAG_GAM0};
                    mcdp_or_philox4x32_10(ctr, key, w);
                    const double ang = 6.283185307179586476925286766559 * (double)(int32_t)w[1] * 0x1p-32;
                    n = sqrt(-2.0 * log(spec_u23(w[0]))) * ((seed & 1u) ? sin(ang) : cos(ang));
                    u = spec_u23((seed & 1u) ? w[3] : w[2]);
                    boost_u = d->p0 != d->malpha ? spec_u23(spec_quad_word(act, seed, 0u, SPEC_TAG_GBST, stream_key)) : 1.0;
                } else {
                    const uint32_t ctr[4] = {seed, act, t, SPEC_TAG_SOLO};
                    mcdp_or_philox4x32_10(ctr, key, w);
                    n = sqrt(-2.0 * log(spec_u23(w[0]))) * cos(6.283185307179586476925286766559 * (double)(int32_t)w[1] * 0x1p-32);
                    u = spec_u23(w[2]);
                    boost_u = spec_u23(w[3]);
                }
                double v = 1.0 + d->a2 * n;
                if (v <= 0.0) continue;
                v = v * v * v;
                const double n2 = n * n;
                if (u > 1.0 - 0.0331 * n2 * n2 && log(u) > 0.5 * n2 + a1 * (1.0 - v + log(v))) continue;
                x = a1 * v * d->p1;
                if (d->p0 != d->malpha) x *= pow(boost_u, 1.0 / d->p0);
                if (x > d->p2) continue;
                return x * base;
            }
            /* attempt cap reached (the reference would spin forever): clamp */
            if (x > d->p2) x = d->p2;
            return x * base;
        }
    }
    return 0.0;
}

int32_t mcdp_or_sim_run_many_spec(mcdp_or_sim* s, const int32_t* seeds, int64_t n, uint32_t stream_key, double* realized,
                                  double* durations, int32_t* cause) {
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t seed = (uint32_t)seeds[i];
        for (int link = 0; link < s->A; ++link) {
            const int32_t di = s->act_dist[link];
            if (di < 0) continue;
            const double b = s->base[link];
            s->dur_scratch[link] = b + spec_sample_extra(&s->dists[di], b, (uint32_t)link, seed, stream_key);
        }
        or_propagate(s, s->dur_scratch, s->realized_scratch, s->cause_scratch);
        or_emit(s, i, realized, durations, cause);
    }
    return 0;
}
