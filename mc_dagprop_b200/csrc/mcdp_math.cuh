// mcdp_math.cuh -- lean device math for the sampler.
//
// The CUDA math library's fp64 log / log1p / pow / cospi are IEEE-careful (denormals, NaN,
// infinities, ~1 ulp) and materialise every polynomial coefficient with two UMOVs; in the first
// profile (profiles/r01_ncu_c3_v0_summary.json) they made up most of the 620 warp instructions
// per 64 edge-samples.  The sampler only ever evaluates them on well-conditioned arguments
// (uniforms strictly inside (0,1)), so the versions here drop the special cases; the log reads a
// small table from shared memory instead of dividing.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>

namespace mcdp {

// ---- natural log through a 512-entry table in shared memory --------------------------------------
// x = 2^e m with m in [0.70711, 1.41421) (the split point is the fp64 high word 0x3FE6A09F, so
// e == 0 around 1 and log keeps its RELATIVE accuracy there).  The 2^20 high-word values of one
// such span are cut into 512 intervals; entry j holds {rc, lc} = {1 / c_j, -log(rc)} for the
// interval's midpoint c_j -- except the interval that contains 1.0, which holds {1, 0} so that
// r = m - 1 is exact.  Then |r| = |m rc - 1| <= 2^-10 and
//     log x = e ln2 + lc + log1p(r),   log1p(r) = r - r^2/2 + r^3/3 - r^4/4 + r^5/5   (next term < 2e-19)
// 9 fp64 + 6 integer instructions and one 16-byte shared-memory load; < 3 ulp for normal positive
// finite x (no zero / denormal / inf / NaN handling: the sampler only passes uniforms in (0,1)
// and products of up to four of them).  The host builds the table (make_log_table) in long double.
constexpr int kLogTabEntries = 512;
constexpr int kLogTabBytes = kLogTabEntries * 16;
constexpr int kLogHiBase = 0x3FE6A09F;

#ifdef __CUDACC__
__device__ __forceinline__ double log_pos(double x, uint32_t log_tab) {
    const int hi = __double2hiint(x);
    const int t = hi - kLogHiBase;
    const int e = t >> 20;  // floor
    const double m = __hiloint2double(hi - (e << 20), __double2loint(x));
    double rc, lc;
    asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(rc), "=d"(lc) : "r"(log_tab + ((uint32_t(t) >> 7) & 0x1FF0u)));
    const double r = fma(m, rc, -1.0);
    const double z = r * r;
    double q = 0.2;
    q = fma(q, r, -0.25);
    q = fma(q, r, 0x1.5555555555555p-2);
    q = fma(q, r, -0.5);
    return fma(double(e), 0x1.62e42fefa39efp-1, lc) + fma(z, q, r);
}
#endif

// rc / lc pairs of the table above; `out` holds 2 * kLogTabEntries doubles
inline void make_log_table(double* out) {
    for (int j = 0; j < kLogTabEntries; ++j) {
        const uint64_t ha = uint64_t(uint32_t(kLogHiBase + j * 2048)) << 32;
        const uint64_t hb = uint64_t(uint32_t(kLogHiBase + (j + 1) * 2048)) << 32;
        double a, b;
        memcpy(&a, &ha, 8);
        memcpy(&b, &hb, 8);
        double rc = 1.0, lc = 0.0;
        if (!(a <= 1.0 && 1.0 < b)) {
            rc = 1.0 / (0.5 * (a + b));
            lc = double(-logl(static_cast<long double>(rc)));
        }
        out[2 * j] = rc;
        out[2 * j + 1] = lc;
    }
}

#ifdef __CUDACC__
// -log1p(-w) for w in [0, 1): series when w is tiny (keeps relative accuracy), log otherwise
__device__ __forceinline__ double neg_log1m(double w, bool tiny, uint32_t log_tab) {
    if (tiny) {  // w < 2^-10: w + w^2/2 + ... + w^6/6, next term < 2^-60 relative
        double p = 1.0 / 6.0;
        p = fma(p, w, 0.2);
        p = fma(p, w, 0.25);
        p = fma(p, w, 1.0 / 3.0);
        p = fma(p, w, 0.5);
        p = fma(p, w, 1.0);
        return p * w;
    }
    return -log_pos(1.0 - w, log_tab);
}

// ---- fp32 hardware approximations (MUFU) used for the gamma sampler's normal deviate and its
// accept/reject decisions; absolute errors ~2^-21..2^-22 (PTX ISA, *.approx.ftz.f32) ----
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float cos_approx(float x) {
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#endif  // __CUDACC__

}  // namespace mcdp
