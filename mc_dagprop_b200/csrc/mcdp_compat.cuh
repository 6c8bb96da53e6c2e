// mcdp_compat.cuh -- reference-stream compatibility sampler.
//
// Draws the durations exactly the way Simulator::run does (reference _core.cpp:313-329): one
// Xoshiro256++ stream per sample, seeded through SplitMix64 from the sign-extended int seed
// (_custom_rng.hpp:507-600), consumed strictly in activity-index order through restatements of the
// libstdc++ 13 transforms (generate_canonical random.tcc:3346-3381, exponential random.h:4897-4905,
// Marsaglia-polar normal with its cached second variate random.tcc:1809-1844, Marsaglia-Tsang gamma
// random.tcc:2352-2393, discrete_distribution random.tcc:2696-2714).  One thread per sample writes
// durations[A][ld]; the max-plus sweep then runs in duration-injection mode.  Constant and
// empirical draws are bit-identical to the reference (integer arithmetic, compares and exactly
// rounded multiplies/adds only); exponential and gamma agree to the last ulps of CUDA's vs
// glibc's log/sqrt/pow.  gamma's cached normal starts empty at every run (the reference lets it
// leak across runs, which no caller can rely on).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mcdp_records.h"
#include "mcdp_sampling.cuh"  // kGammaMaxAttempts: one attempt cap for both generator streams

namespace mcdp {

// one activity in index order
struct alignas(16) ActRec {
    double base;
    uint32_t dist;  // index into DistRec[], kNoDist = none
    uint32_t pad;
};
static_assert(sizeof(ActRec) == 16, "ActRec must be 16 bytes");

struct CompatParams {
    const ActRec* acts;
    const DistRec* dists;
    const double* tab_pool;
    const int32_t* seeds;  // nullptr => seed0 + sample index
    double* durations;     // [A][ld]
    double* norm_cache;    // [n_dists][ld] scratch: gamma's cached normal per (distribution, sample)
    int64_t n, ld;
    int32_t A, n_dists;
    int32_t seed0;
};

struct Xoshiro256pp {
    uint64_t s0, s1, s2, s3;
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t& st) {
    uint64_t z = (st += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
__device__ __forceinline__ uint64_t xoshiro_next(Xoshiro256pp& g) {
    const uint64_t result = rotl64(g.s0 + g.s3, 23) + g.s0;
    const uint64_t t = g.s1 << 17;
    g.s2 ^= g.s0;
    g.s3 ^= g.s1;
    g.s1 ^= g.s2;
    g.s0 ^= g.s3;
    g.s2 ^= t;
    g.s3 = rotl64(g.s3, 45);
    return result;
}
// generate_canonical<double, 53> over a 64-bit URBG: double(u64) / 2^64, clamped below 1
__device__ __forceinline__ double canonical(Xoshiro256pp& g) {
    const double r = __ull2double_rn(xoshiro_next(g)) * 0x1p-64;
    return r >= 1.0 ? 0x1.fffffffffffffp-1 : r;
}

__device__ __forceinline__ double compat_normal(Xoshiro256pp& g, double* cache, bool& cached) {
    if (cached) {
        cached = false;
        return __dadd_rn(__dmul_rn(*cache, 1.0), 0.0);
    }
    double x, y, r2;
    do {
        x = __dadd_rn(__dmul_rn(2.0, canonical(g)), -1.0);
        y = __dadd_rn(__dmul_rn(2.0, canonical(g)), -1.0);
        r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    } while (r2 > 1.0 || r2 == 0.0);
    const double mult = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
    *cache = __dmul_rn(x, mult);
    cached = true;
    return __dadd_rn(__dmul_rn(__dmul_rn(y, mult), 1.0), 0.0);
}

__global__ void __launch_bounds__(128) compat_sample_kernel(const __grid_constant__ CompatParams p) {
    const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= p.ld) return;
    const int32_t seed = p.seeds ? (s < p.n ? __ldg(p.seeds + s) : 0) : int32_t(uint32_t(p.seed0) + uint32_t(s));
    Xoshiro256pp g;
    {
        uint64_t st = uint64_t(int64_t(seed));  // int -> uint64 sign-extends (rng_.seed(seed), _core.cpp:313)
        g.s0 = splitmix64(st);
        g.s1 = splitmix64(st);
        g.s2 = splitmix64(st);
        g.s3 = splitmix64(st);
    }
    uint64_t cached_mask = 0;  // bit d: distribution d holds a cached normal (n_dists <= 64 on this path)
    for (int a = 0; a < p.A; ++a) {
        const int4 araw = __ldg(reinterpret_cast<const int4*>(p.acts + a));
        ActRec ar;
        ar.base = __hiloint2double(araw.y, araw.x);
        ar.dist = uint32_t(araw.z);
        double dur = ar.base;
        if (ar.dist != kNoDist) {
            const DistRec& d = p.dists[ar.dist];
            double extra;
            switch (d.kind) {
                case MCDP_DIST_CONSTANT:
                    extra = __dmul_rn(ar.base, d.p[0]);
                    break;
                case MCDP_DIST_EXPONENTIAL: {
                    const double rate = d.p[7];  // 1.0 / lambda, divided on the host like ExponentialDist's ctor
                    double x;
                    uint32_t tries = 0u;  // the reference spins until x <= max_scale; capped like the Philox path
                    do {
                        x = __ddiv_rn(-log(__dadd_rn(1.0, -canonical(g))), rate);
                    } while (x > d.p[1] && ++tries < kGammaMaxAttempts);
                    if (x > d.p[1]) x = d.p[1];
                    extra = __dmul_rn(x, ar.base);
                    break;
                }
                case MCDP_DIST_GAMMA: {
                    const double alpha = d.p[0], beta = d.p[1], a1 = d.p[3], a2 = d.p[4];
                    double* cache = p.norm_cache + size_t(ar.dist) * p.ld + s;
                    bool cached = (cached_mask >> ar.dist) & 1ull;
                    double x;
                    uint32_t tries = 0u;  // same cap for the truncation loop (a tiny max_scale would hang the device)
                    do {
                        double u, v, n;
                        do {
                            do {
                                n = compat_normal(g, cache, cached);
                                v = __dadd_rn(1.0, __dmul_rn(a2, n));
                            } while (v <= 0.0);
                            v = __dmul_rn(__dmul_rn(v, v), v);
                            u = canonical(g);
                            // u > 1 - 0.0331 n^4  &&  log(u) > 0.5 n^2 + a1 (1 - v + log v)
                        } while (u > __dadd_rn(1.0, -__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(0.0331, n), n), n), n)) &&
                                 log(u) > __dadd_rn(__dmul_rn(__dmul_rn(0.5, n), n),
                                                    __dmul_rn(a1, __dadd_rn(__dadd_rn(1.0, -v), log(v)))));
                        if (!(d.flags & 1)) {
                            x = __dmul_rn(__dmul_rn(a1, v), beta);
                        } else {
                            do u = canonical(g);
                            while (u == 0.0);
                            x = __dmul_rn(__dmul_rn(__dmul_rn(pow(u, __ddiv_rn(1.0, alpha)), a1), v), beta);
                        }
                    } while (x > d.p[2] && ++tries < kGammaMaxAttempts);
                    if (x > d.p[2]) x = d.p[2];
                    cached_mask = cached ? (cached_mask | (1ull << ar.dist)) : (cached_mask & ~(1ull << ar.dist));
                    extra = __dmul_rn(x, ar.base);
                    break;
                }
                default: {  // empirical tables: std::lower_bound over the cumulative array, no draw for < 2 entries
                    const int len = d.tab_len;
                    const double* cp = p.tab_pool + d.tab_off + guide_doubles(uint32_t(d.guide_log2));
                    int lo = 0;
                    if (len >= 2) {
                        const double u = canonical(g);
                        int cnt = len;
                        while (cnt > 0) {
                            const int half = cnt >> 1;
                            if (__ldg(cp + lo + half) < u) {
                                lo += half + 1;
                                cnt -= half + 1;
                            } else {
                                cnt = half;
                            }
                        }
                    }
                    const double v = __ldg(cp + len + lo);
                    extra = d.kind == MCDP_DIST_EMP_ABS ? v : __dmul_rn(v, ar.base);
                    break;
                }
            }
            dur = __dadd_rn(ar.base, extra);
        }
        p.durations[size_t(a) * p.ld + s] = dur;
    }
}

}  // namespace mcdp
