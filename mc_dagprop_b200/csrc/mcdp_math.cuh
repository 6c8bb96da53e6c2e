// mcdp_math.cuh -- lean device math for the sampler.
//
// The CUDA math library's fp64 log / log1p / pow / cospi are IEEE-careful (denormals, NaN,
// infinities, ~1 ulp) and materialise every polynomial coefficient with two UMOVs; in the first
// profile (profiles/r01_ncu_c3_v0_summary.json) they made up most of the 620 warp instructions
// per 64 edge-samples.  The sampler only ever evaluates them on well-conditioned arguments
// (uniforms strictly inside (0,1)), so the versions here drop the special cases and read their
// coefficients as constant-bank operands.
#pragma once
#include <cuda_runtime.h>

namespace mcdp {

// log(m) = 2 s + 2 s^3 g(s^2), s = (m-1)/(m+1), m in [sqrt(.5), sqrt(2)); g fitted (Chebyshev
// interpolation in 80-bit arithmetic, scripts/fit_log_poly.py) to 1.6e-16 absolute on s^2 <= 0.0295,
// i.e. < 1e-17 relative in log(m).
__constant__ double kLogG[7] = {0x1.5555555555558p-2, 0x1.99999999952aap-3, 0x1.2492492df775fp-3, 0x1.c71c62dd9fff0p-4,
                                0x1.7462b6e894664p-4, 0x1.39fe16006493ap-4, 0x1.2b5be18007317p-4};

__device__ __forceinline__ double rcp_approx(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));  // MUFU.RCP64H, ~2^-23 relative
    return r;
}

// 1/d for normal positive d, two Newton steps from the hardware seed (< 1 ulp-ish, no special cases)
__device__ __forceinline__ double rcp_pos(double d) {
    double r = rcp_approx(d);
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

// natural log of a normal, positive, finite x (no zero / denormal / inf / NaN handling)
__device__ __forceinline__ double log_pos(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000FFFFF) | 0x3FF00000;
    if (hi >= 0x3FF6A09F) {  // m > sqrt(2): halve so that m is in [0.7071, 1.4142)
        hi -= 0x00100000;
        e += 1;
    }
    const double m = __hiloint2double(hi, lo);
    const double f = m - 1.0;
    const double s = f * rcp_pos(m + 1.0);
    const double z = s * s;
    double g = kLogG[6];
    g = fma(g, z, kLogG[5]);
    g = fma(g, z, kLogG[4]);
    g = fma(g, z, kLogG[3]);
    g = fma(g, z, kLogG[2]);
    g = fma(g, z, kLogG[1]);
    g = fma(g, z, kLogG[0]);
    const double lm = fma(s * z, g + g, s + s);
    const double de = double(e);
    return fma(de, 0x1.62e42fefa39efp-1, fma(de, 0x1.abc9e3b39803fp-56, lm));  // e*ln2 (hi + lo) + log(m)
}

// -log1p(-w) for w in [0, 1): series when w is tiny (keeps relative accuracy), log otherwise
__device__ __forceinline__ double neg_log1m(double w, bool tiny) {
    if (tiny) {  // w < 2^-10: w + w^2/2 + ... + w^6/6, next term < 2^-60 relative
        double p = 1.0 / 6.0;
        p = fma(p, w, 0.2);
        p = fma(p, w, 0.25);
        p = fma(p, w, 1.0 / 3.0);
        p = fma(p, w, 0.5);
        p = fma(p, w, 1.0);
        return p * w;
    }
    return -log_pos(1.0 - w);
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

}  // namespace mcdp
