// mcdp_sampling.cuh -- device generator "mcdp-philox-v2" (DESIGN.md section 4).
//
// Replaces the reference's sequential Xoshiro256++ stream (_custom_rng.hpp:551-600) and the
// libstdc++ <random> transforms behind Dist::sample (_core.cpp:72-141) by counter-based draws
// keyed on (seed, activity index): every (seed, activity) pair owns its random bits, so the
// result of a sample does not depend on which thread, chunk or GPU computes it and the
// activities can be drawn in evaluation order, fused with the max-plus sweep.
//
//   key     = (stream_key, 'MCDP')                     -- round keys precomputed, read as constant operands
//   QUAD    ctr = (seed >> 2, act, j, 'QUAD')          -- one block serves seeds {4k .. 4k+3}: 32 bits each (word seed & 3):
//                                                         empirical tables of <= 4096 entries, exponentials (j = 1
//                                                         refines a draw that falls into the top 2^-20), gamma shape 1
//   PAIR    ctr = (seed >> 1, act, j, 'PAIR')          -- one block serves seeds {2k, 2k+1}: 64 bits each (larger tables)
//   SOLO    ctr = (seed,      act, t, 'SOLO')          -- gamma attempt t >= 1: 4 x 32 bits
//   GAM0    ctr = (seed >> 1, act, 0, 'GAM0')          -- gamma attempt 0 of the seed pair: one Box-Muller pair
//                                                         (cos branch: even seed, sin branch: odd seed) + two accept words
//   GBST    ctr = (seed >> 2, act, 0, 'GBST')          -- shape < 1 boost uniform of attempt 0, word seed & 3
// v1 (round 1) spent a 64-bit PAIR draw on every table lookup and exponential: two Philox blocks per lane-quad where
// one suffices, 26 % of all instructions of the headline workload were Philox rounds.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "mcdp_math.cuh"
#include "mcdp_records.h"

namespace mcdp {

constexpr uint32_t kKey1 = 0x4D434450u;     // 'MCDP'
constexpr uint32_t kTagPair = 0x50414952u;  // 'PAIR'
constexpr uint32_t kTagSolo = 0x534F4C4Fu;  // 'SOLO'
constexpr uint32_t kTagQuad = 0x51554144u;  // 'QUAD'
constexpr uint32_t kTagGam0 = 0x47414D30u;  // 'GAM0'
constexpr uint32_t kTagGbst = 0x47425354u;  // 'GBST'
constexpr uint32_t kGammaMaxAttempts = 65536u;

struct Philox4 {
    uint32_t x, y, z, w;
};

// 32 x 32 -> (hi, lo).  Written as mul.wide + register-pair unpack: the plain C form
// `uint64_t p = (uint64_t)a * b; hi = p >> 32` makes ptxas 12.9 add a zero-valued uniform register
// to every high word (one extra VIADD per multiply, +50% on the whole generator).
__device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0,%1}, p;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

// The ten Philox round keys of word 0, key0 + r * 0x9E3779B9, precomputed on the host: they arrive
// through the kernel parameter block, so each round's XOR reads its key as a constant-bank operand.
struct PhiloxKeys {
    uint32_t k[10];
};

// Philox4x32-10 (Salmon et al., SC'11): per block 20 IMAD.WIDE + 20 LOP3.
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 const PhiloxKeys& key0) {
    uint32_t k1 = kKey1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        mulhilo32(0xD2511F53u, c0, h0, l0);
        mulhilo32(0xCD9E8D57u, c2, h1, l1);
        const uint32_t n0 = h1 ^ c1 ^ key0.k[r];
        const uint32_t n2 = h0 ^ c3 ^ k1;
        c1 = l1;
        c3 = l0;
        c0 = n0;
        c2 = n2;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// The same block out of line, for the cold paths of the samplers (seeds that are not an aligned run of four, redraws
// after a truncation): ~100 instructions that would otherwise be inlined at every such site and weigh on the
// register allocation and the instruction footprint of the sweep loop.
__device__ __noinline__ Philox4 philox4x32_10_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys* key0) {
    return philox4x32_10(c0, c1, c2, c3, *key0);
}

// ((X >> 12) + 0.5) * 2^-52 for X = hi:lo, in (0,1): one exact subtraction, no int->fp conversion.
__device__ __forceinline__ double uniform52(uint32_t lo, uint32_t hi) {
    const uint32_t khi = hi >> 12;
    const uint32_t klo = (lo >> 12) | (hi << 20);
    return __hiloint2double(static_cast<int>(0x3FF00000u | khi), static_cast<int>(klo)) - (1.0 - 0x1p-53);
}

// (w + 0.5) * 2^-32 in (0,1), exact.
__device__ __forceinline__ double uniform32(uint32_t w) {
    return __hiloint2double(0x41300000, static_cast<int>(w)) - (1048576.0 - 0x1p-33);
}

__device__ __forceinline__ uint32_t philox_word(const Philox4& r, uint32_t k) {
    return k == 0u ? r.x : (k == 1u ? r.y : (k == 2u ? r.z : r.w));
}

// one 32-bit draw for each of the two samples of a pair-kernel lane (QUAD-style blocks with tag `tag`)
__device__ __forceinline__ void draw32x2(uint32_t seed_a, uint32_t seed_b, uint32_t ja, uint32_t jb, uint32_t act, uint32_t tag,
                                         const PhiloxKeys& key0, uint32_t& wa, uint32_t& wb) {
    if ((seed_a >> 2) == (seed_b >> 2) && ja == jb) {
        const Philox4 r = philox4x32_10(seed_a >> 2, act, ja, tag, key0);
        wa = philox_word(r, seed_a & 3u);
        wb = philox_word(r, seed_b & 3u);
    } else {
        wa = philox_word(philox4x32_10(seed_a >> 2, act, ja, tag, key0), seed_a & 3u);
        wb = philox_word(philox4x32_10(seed_b >> 2, act, jb, tag, key0), seed_b & 3u);
    }
}

// The refined tail of an exponential draw (contract v2): w >= kExpTailWord is the sample's first word; the second word
// v is the seed's word of QUAD block j = 1; 1 - u = ((2^32 - w) - (v + 1/2) 2^-32) 2^-32 exactly and
// x = -lambda ln((1 - F) + F (1 - u)).  One sample in 2^20 comes here: kept out of line so that its Philox block and
// log do not weigh on the register allocation of the sweep loop.
__device__ __forceinline__ double exp_tail(double lam, double F, double one_minus_F, uint32_t w, uint32_t seed, uint32_t act,
                                        const PhiloxKeys* key0, uint32_t log_tab) {
    const uint32_t v = philox_word(philox4x32_10(seed >> 2, act, 1u, kTagQuad, *key0), seed & 3u);
    const double one_minus_u = (double(0u - w) - uniform32(v)) * 0x1p-32;
    return -lam * log_pos(fma(F, one_minus_u, one_minus_F), log_tab);
}

// Table / distribution-record access.  SMEM: the staged copy in shared memory, addressed by 32-bit
// shared-window addresses and explicit ld.shared (no generic pointers: the kernel keeps ONE base
// register instead of re-deriving the window base at every use); else global memory through L1/L2.
template <bool SMEM>
struct Mem;
template <>
struct Mem<true> {
    using ptr = uint32_t;
    static __device__ __forceinline__ double f64(ptr a) {
        double v;
        asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ uint32_t u32(ptr a) {
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    }
};
template <>
struct Mem<false> {
    using ptr = const char*;
    static __device__ __forceinline__ double f64(ptr a) { return __ldg(reinterpret_cast<const double*>(a)); }
    static __device__ __forceinline__ uint32_t u32(ptr a) { return __ldg(reinterpret_cast<const uint32_t*>(a)); }
};
// one DistRec, read field by field on demand
template <bool SMEM>
struct DistView {
    typename Mem<SMEM>::ptr base;
    __device__ __forceinline__ double p(int k) const { return Mem<SMEM>::f64(base + (32 + 8 * k)); }
    __device__ __forceinline__ int flags() const { return int(Mem<SMEM>::u32(base + 16)); }
    __device__ __forceinline__ int pad0() const { return int(Mem<SMEM>::u32(base + 20)); }
    __device__ __forceinline__ int pad1() const { return int(Mem<SMEM>::u32(base + 24)); }
};
static_assert(offsetof(DistRec, flags) == 16 && offsetof(DistRec, pad0) == 20 && offsetof(DistRec, pad1) == 24 &&
                  offsetof(DistRec, p) == 32,
              "DistView offsets follow DistRec");

// exp_tail with the parameters read from the distribution record inside the out-of-line function (fewer live
// registers at the call site)
template <bool SMEM>
__device__ __noinline__ double exp_tail_rec(typename Mem<SMEM>::ptr rec, uint32_t w, uint32_t seed, uint32_t act,
                                            const PhiloxKeys* key0, uint32_t log_tab) {
    const DistView<SMEM> d{rec};
    return exp_tail(d.p(0), d.p(2), d.p(3), w, seed, act, key0, log_tab);
}

// Inverse-CDF lookup with std::lower_bound semantics (first cp[i] >= u; libstdc++
// random.tcc:2709-2713): the guide table (four buckets per entry) gives the first candidate, one
// unconditional compare-and-step follows, and a scan loop that almost never iterates finishes.
template <bool SMEM>
__device__ __forceinline__ void emp_value2(typename Mem<SMEM>::ptr guide_b, typename Mem<SMEM>::ptr cp_b, uint32_t g,
                                           uint32_t len8, bool scan, uint32_t hi_a, double ua, uint32_t hi_b, double ub,
                                           double& va, double& vb) {
    // floor(u * 2^g) == (X >> 12) >> (52 - g) == hi >> (32 - g),  1 <= g <= 20.  The two samples of
    // the lane are looked up in lockstep (one basic block) so their shared-memory latencies overlap.
    const uint32_t ja = hi_a >> (32u - g), jb = hi_b >> (32u - g);
    uint32_t off_a = Mem<SMEM>::u32(guide_b + ja * 4u) * 8u;
    uint32_t off_b = Mem<SMEM>::u32(guide_b + jb * 4u) * 8u;
    const double ca = Mem<SMEM>::f64(cp_b + off_a);
    const double cb = Mem<SMEM>::f64(cp_b + off_b);
    off_a += ca < ua ? 8u : 0u;  // cp[len-1] == 1.0 > u: stays in range
    off_b += cb < ub ? 8u : 0u;
    if (scan) {  // only tables whose guide buckets may hold several boundaries
        while (Mem<SMEM>::f64(cp_b + off_a) < ua) off_a += 8u;
        while (Mem<SMEM>::f64(cp_b + off_b) < ub) off_b += 8u;
    }
    va = Mem<SMEM>::f64(cp_b + off_a + len8);
    vb = Mem<SMEM>::f64(cp_b + off_b + len8);
}

// (k + 1/2) * 2^-23 for the top 23 bits k of w: an fp32 uniform strictly inside (0,1), exact.
__device__ __forceinline__ float uniform23(uint32_t w) {
    return __uint_as_float(0x3F800000u | (w >> 9)) - (1.0f - 0x1p-24f);
}

// One Marsaglia-Tsang attempt (libstdc++ random.tcc:2352-2393 restated for SIMT).  The normal deviate (Box-Muller)
// and the two accept/reject comparisons are evaluated with fp32 hardware approximations -- they only steer the
// draw -- while the variate itself, x = d * v^3 * scale, is formed in fp64.  An attempt is accepted when the
// Marsaglia-Tsang test passes and x <= max_scale (the reference's outer `while (x > max_scale)` loop,
// _core.cpp:98-104, simply continues the attempt sequence).
//   attempt 0: the GAM0 block of the seed pair carries ONE Box-Muller pair (radius word, angle word) and one accept
//              word per seed: the even seed takes r cos(a), the odd seed r sin(a) -- two independent normals for
//              the price of one block; the shape < 1 boost uniform of attempt 0 is the seed's GBST word;
//   attempt t >= 1: the SOLO block of (seed, act, t): radius, angle (cos branch), accept word, boost word.
struct GammaHalf {  // the fp32-steered part of one attempt, split so that several attempts can run in lockstep
    float nf, n2, u;
    double v;
    bool pos, ok;
};
__device__ __forceinline__ float sin_approx(float x) {
    float r;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Box-Muller radius and angle of a (radius word, angle word) pair
__device__ __forceinline__ void box_muller_polar(uint32_t w_radius, uint32_t w_angle, float& r, float& ang) {
    const float r2 = -1.3862943611198906f * lg2_approx(uniform23(w_radius));  // -2 ln u1
    ang = float(int(w_angle)) * 1.4629180792671596e-9f;                        // 2 pi * int32 / 2^32, [-pi, pi)
    r = sqrt_approx(fmaxf(r2, 0.0f));
}
template <bool SMEM>
__device__ __forceinline__ void gamma_front(const DistView<SMEM>& d, float nf, uint32_t w_accept, GammaHalf& h) {
    h.nf = nf;
    double v = fma(d.p(4), double(nf), 1.0);
    h.pos = v > 0.0;
    h.v = v * v * v;
    h.u = uniform23(w_accept);
    h.n2 = nf * nf;
    h.ok = h.u <= fmaf(-0.0331f * h.n2, h.n2, 1.0f);
}
template <bool SMEM>
__device__ __forceinline__ void gamma_exact(const DistView<SMEM>& d, GammaHalf& h) {
    // log(u) <= n^2/2 + d (1 - v + log v)
    const float vf = float(h.v);
    h.ok = 0.6931471805599453f * lg2_approx(h.u) <=
           fmaf(0.5f, h.n2, float(d.p(3)) * (1.0f - vf + 0.6931471805599453f * lg2_approx(vf)));
}
template <bool SMEM>
__device__ __forceinline__ bool gamma_back(const DistView<SMEM>& d, uint32_t w_boost, const GammaHalf& h, double& x) {
    x = d.p(6) * h.v;  // d * scale * v^3
    if (d.flags() & 1) x *= double(ex2_approx(lg2_approx(uniform23(w_boost)) * float(d.p(5))));  // u^(1/shape), shape < 1
    return h.pos && h.ok && x <= d.p(2);
}
// attempt t >= 1 of one sample
template <bool SMEM>
__device__ __forceinline__ bool gamma_eval(const DistView<SMEM>& d, const Philox4& w, double& x) {
    GammaHalf h;
    float r, ang;
    box_muller_polar(w.x, w.y, r, ang);
    gamma_front<SMEM>(d, r * cos_approx(ang), w.z, h);
    if (!h.ok) gamma_exact<SMEM>(d, h);
    return gamma_back<SMEM>(d, w.w, h, x);
}
// The first attempts (t = 0) of N samples in lockstep: nf[i] / w_accept[i] / w_boost[i] prepared by the caller from the
// GAM0 / GBST blocks.  One shared branch for the exact test, everything else straight-line.
template <bool SMEM, int N>
__device__ __forceinline__ void gamma_first_attempts(const DistView<SMEM>& d, const float (&nf)[N], const uint32_t (&w_accept)[N],
                                                     const uint32_t (&w_boost)[N], double (&x)[N], bool (&ok)[N]) {
    GammaHalf h[N];
    bool all = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        gamma_front<SMEM>(d, nf[i], w_accept[i], h[i]);
        all = all && h[i].ok;
    }
    if (!all) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool squeeze = h[i].ok;  // a passed squeeze test stays accepted
            gamma_exact<SMEM>(d, h[i]);
            h[i].ok = h[i].ok || squeeze;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) ok[i] = gamma_back<SMEM>(d, w_boost[i], h[i], x[i]);
}
// the normal of `seed`'s attempt 0 out of its pair's GAM0 block, and its accept word
__device__ __forceinline__ void gam0_take(const Philox4& blk, uint32_t seed, float& nf, uint32_t& w_accept) {
    float r, ang;
    box_muller_polar(blk.x, blk.y, r, ang);
    const bool odd = seed & 1u;
    nf = r * (odd ? sin_approx(ang) : cos_approx(ang));
    w_accept = odd ? blk.w : blk.z;
}

// Gamma variates for the two samples of a thread.  First attempts run straight-line for both
// samples; afterwards the whole warp iterates a uniform retry loop in which every lane retries one
// pending sample, so a rejection costs the warp one extra attempt instead of one per sample.
template <bool SMEM>
__device__ __forceinline__ void gamma_variate2(const DistView<SMEM>& d, uint32_t seed_a, uint32_t seed_b, bool paired,
                                               uint32_t act, const PhiloxKeys& key0, double& xa, double& xb) {
    float nf[2];
    uint32_t wacc[2], wboost[2] = {0u, 0u};
    if (paired) {  // seeds {2k, 2k+1}: one GAM0 block, cos branch / sin branch
        const Philox4 blk = philox4x32_10(seed_a >> 1, act, 0u, kTagGam0, key0);
        float r, ang;
        box_muller_polar(blk.x, blk.y, r, ang);
        nf[0] = r * cos_approx(ang);
        nf[1] = r * sin_approx(ang);
        wacc[0] = blk.z;
        wacc[1] = blk.w;
    } else {
        gam0_take(philox4x32_10(seed_a >> 1, act, 0u, kTagGam0, key0), seed_a, nf[0], wacc[0]);
        gam0_take(philox4x32_10(seed_b >> 1, act, 0u, kTagGam0, key0), seed_b, nf[1], wacc[1]);
    }
    if (d.flags() & 1) draw32x2(seed_a, seed_b, 0u, 0u, act, kTagGbst, key0, wboost[0], wboost[1]);
    double x[2];
    bool ok[2];
    gamma_first_attempts<SMEM, 2>(d, nf, wacc, wboost, x, ok);
    xa = x[0];
    xb = x[1];
    bool need_a = !ok[0], need_b = !ok[1];
    uint32_t ta = 1u, tb = 1u;
    while (__any_sync(0xFFFFFFFFu, need_a || need_b)) {
        const bool do_a = need_a;
        const uint32_t seed = do_a ? seed_a : seed_b;
        const uint32_t t = do_a ? ta : tb;
        double xx;
        const bool acc = gamma_eval<SMEM>(d, philox4x32_10(seed, act, t, kTagSolo, key0), xx);
        const bool give_up = t + 1u >= kGammaMaxAttempts;  // the reference would spin forever: clamp
        if (give_up) xx = fmin(xx, d.p(2));
        if (do_a) {
            ++ta;
            if (acc || give_up) {
                xa = xx;
                need_a = false;
            }
        } else if (need_b) {
            ++tb;
            if (acc || give_up) {
                xb = xx;
                need_b = false;
            }
        }
    }
}

// Gamma with 2 * shape in {1, ..., 6, 8}: exact transformation without rejection.
//   Gamma(k + h/2, scale) = scale * ( -ln(u_1 ... u_k)  +  h * (-ln u') cos^2(2 pi u'') ),   k = floor(shape), h in {0,1}
// (a sum of k unit exponentials plus, for half-integer shapes, half the square of a Box-Muller
// normal).  The logs are fp64 (log_pos), cos is the fp32 hardware approximation.  Up to two 32-bit
// uniforms per sample come from a PAIR-style block shared by the seed pair (one: from a QUAD-style block shared by
// four seeds), three or four from a block per seed; draw j + 1 is used when x > max_scale (the reference's outer
// redraw loop, _core.cpp:98-104).
constexpr uint32_t kTagErlang = 0x45524C47u;  // 'ERLG'

__device__ __forceinline__ double half_term(uint32_t wu, uint32_t wa, uint32_t log_tab) {
    const float c = cos_approx(float(int(wa)) * 1.4629180792671596e-9f);  // cos(2 pi * int32 / 2^32)
    return -log_pos(uniform32(wu), log_tab) * double(c * c);
}
// words are consumed in order: k product uniforms, then (u', u'') of the half term.  `variant` = 2 * k + h
// (warp-uniform): one straight-line case per supported shape.
template <bool SMEM>
__device__ __forceinline__ double erlang_value(const DistView<SMEM>& d, int variant, uint32_t w0, uint32_t w1, uint32_t w2,
                                               uint32_t w3, uint32_t log_tab) {
    double e;
    switch (variant) {
        case 1: e = half_term(w0, w1, log_tab); break;                                                   // shape 1/2
        case 2: e = -log_pos(uniform32(w0), log_tab); break;                                             // 1
        case 3: e = half_term(w1, w2, log_tab) - log_pos(uniform32(w0), log_tab); break;                 // 3/2
        case 4: e = -log_pos(uniform32(w0) * uniform32(w1), log_tab); break;                             // 2
        case 5: e = half_term(w2, w3, log_tab) - log_pos(uniform32(w0) * uniform32(w1), log_tab); break;  // 5/2
        case 6: e = -log_pos(uniform32(w0) * uniform32(w1) * uniform32(w2), log_tab); break;             // 3
        default: e = -log_pos((uniform32(w0) * uniform32(w1)) * (uniform32(w2) * uniform32(w3)), log_tab); break;  // 4
    }
    return d.p(1) * e;
}

// draw j of both samples; one Philox block serves the seed pair when a sample needs <= 2 words
template <bool SMEM>
__device__ __forceinline__ void erlang_draw2(const DistView<SMEM>& d, int variant, uint32_t seed_a, uint32_t seed_b, bool paired,
                                             uint32_t act, uint32_t ja, uint32_t jb, const PhiloxKeys& key0,
                                             uint32_t log_tab, double& ya, double& yb) {
    if (variant == 2) {  // one 32-bit uniform per sample: QUAD-style block
        uint32_t wa, wb;
        draw32x2(seed_a, seed_b, ja, jb, act, kTagErlang, key0, wa, wb);
        ya = erlang_value<SMEM>(d, variant, wa, 0u, 0u, 0u, log_tab);
        yb = erlang_value<SMEM>(d, variant, wb, 0u, 0u, 0u, log_tab);
    } else if (variant == 1 || variant == 4) {  // 64 bits per sample: PAIR-style block
        if (paired && ja == jb) {
            const Philox4 r = philox4x32_10(seed_a >> 1, act, ja, kTagErlang, key0);
            ya = erlang_value<SMEM>(d, variant, r.x, r.y, 0u, 0u, log_tab);
            yb = erlang_value<SMEM>(d, variant, r.z, r.w, 0u, 0u, log_tab);
        } else {
            const Philox4 ra = philox4x32_10(seed_a >> 1, act, ja, kTagErlang, key0);
            const Philox4 rb = philox4x32_10(seed_b >> 1, act, jb, kTagErlang, key0);
            const bool oa = seed_a & 1u, ob = seed_b & 1u;
            ya = erlang_value<SMEM>(d, variant, oa ? ra.z : ra.x, oa ? ra.w : ra.y, 0u, 0u, log_tab);
            yb = erlang_value<SMEM>(d, variant, ob ? rb.z : rb.x, ob ? rb.w : rb.y, 0u, 0u, log_tab);
        }
    } else {
        const Philox4 ra = philox4x32_10(seed_a, act, ja, kTagErlang, key0);
        const Philox4 rb = philox4x32_10(seed_b, act, jb, kTagErlang, key0);
        ya = erlang_value<SMEM>(d, variant, ra.x, ra.y, ra.z, ra.w, log_tab);
        yb = erlang_value<SMEM>(d, variant, rb.x, rb.y, rb.z, rb.w, log_tab);
    }
}

template <bool SMEM>
__device__ __forceinline__ void erlang_variate2(const DistView<SMEM>& d, uint32_t seed_a, uint32_t seed_b, bool paired,
                                                uint32_t act, const PhiloxKeys& key0, uint32_t log_tab, double& xa,
                                                double& xb) {
    const int variant = 2 * d.pad0() + d.pad1();
    const double mx = d.p(2);
    erlang_draw2<SMEM>(d, variant, seed_a, seed_b, paired, act, 0u, 0u, key0, log_tab, xa, xb);
    // truncation (_core.cpp:98-104): draws 1, 2, ... until x <= max_scale; rare, so off the straight path
    bool need_a = xa > mx, need_b = xb > mx;
    if (__any_sync(0xFFFFFFFFu, need_a || need_b)) {
        uint32_t ja = 0u, jb = 0u;
        do {
            ja += need_a ? 1u : 0u;
            jb += need_b ? 1u : 0u;
            double ya, yb;
            erlang_draw2<SMEM>(d, variant, seed_a, seed_b, paired, act, ja, jb, key0, log_tab, ya, yb);
            if (need_a) {
                xa = ya;
                need_a = ya > mx && ja + 1u < kGammaMaxAttempts;
            }
            if (need_b) {
                xb = yb;
                need_b = yb > mx && jb + 1u < kGammaMaxAttempts;
            }
        } while (__any_sync(0xFFFFFFFFu, need_a || need_b));
        xa = xa > mx ? mx : xa;  // only after the attempt cap
        xb = xb > mx ? mx : xb;
    }
}

// Extra delays of one activity for the two samples a thread owns.  `meta`/`tab_off` come from the
// precedence record (kind, guide bits, table length / pool block).  `paired`: the seeds are
// {2k, 2k+1}, so one PAIR block serves both.  Returns extra (the value Dist::sample returns); the
// caller forms base + extra with separately rounded operations like the reference build.
template <bool SMEM>
__device__ __forceinline__ void sample_extra2(uint32_t meta, uint32_t tab_off, typename Mem<SMEM>::ptr dists,
                                              uint32_t dist, typename Mem<SMEM>::ptr tab, double base, uint32_t act,
                                              uint32_t seed_a, uint32_t seed_b, bool paired, const PhiloxKeys& key0,
                                              uint32_t log_tab, double& ea, double& eb) {
    const uint32_t kind = meta >> 29;
    if (kind == MCDP_DIST_CONSTANT) {
        const DistView<SMEM> d{dists + dist * uint32_t(sizeof(DistRec))};
        ea = eb = __dmul_rn(base, d.p(0));  // _core.cpp:75
        return;
    }
    if (kind == MCDP_DIST_GAMMA) {
        const DistView<SMEM> d{dists + dist * uint32_t(sizeof(DistRec))};
        double xa, xb;
        if (d.flags() & 8)
            erlang_variate2<SMEM>(d, seed_a, seed_b, paired, act, key0, log_tab, xa, xb);
        else
            gamma_variate2<SMEM>(d, seed_a, seed_b, paired, act, key0, xa, xb);
        ea = __dmul_rn(xa, base);
        eb = __dmul_rn(xb, base);
        return;
    }
    if (kind == MCDP_DIST_EXPONENTIAL) {
        // inverse CDF of the exponential truncated to [0, max_scale]: the law of the reference's rejection loop
        // (_core.cpp:83-89), without the loop.  One 32-bit word per sample, refined by a second one in the far tail.
        const DistView<SMEM> d{dists + dist * uint32_t(sizeof(DistRec))};
        uint32_t wa, wb;
        draw32x2(seed_a, seed_b, 0u, 0u, act, kTagQuad, key0, wa, wb);
        const double lam = d.p(0), mx = d.p(1), F = d.p(2);
        const bool tiny = d.flags() & 2;  // F < 2^-10: series keeps the relative accuracy
        double xa, xb;
        if (tiny) {
            xa = lam * neg_log1m(uniform32(wa) * F, true, log_tab);
            xb = lam * neg_log1m(uniform32(wb) * F, true, log_tab);
        } else {
            xa = lam * neg_log1m(uniform32(wa) * F, false, log_tab);
            xb = lam * neg_log1m(uniform32(wb) * F, false, log_tab);
            if (wa >= kExpTailWord) xa = exp_tail_rec<SMEM>(d.base, wa, seed_a, act, &key0, log_tab);
            if (wb >= kExpTailWord) xb = exp_tail_rec<SMEM>(d.base, wb, seed_b, act, &key0, log_tab);
        }
        xa = xa > mx ? mx : xa;
        xb = xb > mx ? mx : xb;
        ea = __dmul_rn(xa, base);
        eb = __dmul_rn(xb, base);
        return;
    }
    // One uniform per sample for the table lookup: 32 bits from a QUAD block for tables of <= 4096 entries, else 64
    // bits from a PAIR block.
    uint32_t hi_a, hi_b;
    double ua, ub;
    if ((meta & 0x7FFFFFu) <= kQuadTableMaxLen) {
        draw32x2(seed_a, seed_b, 0u, 0u, act, kTagQuad, key0, hi_a, hi_b);
        ua = uniform32(hi_a);
        ub = uniform32(hi_b);
    } else {
        uint32_t lo_a, lo_b;
        if (paired) {  // seeds {2k, 2k+1}: one block, words 0-1 / 2-3
            const Philox4 r = philox4x32_10(seed_a >> 1, act, 0u, kTagPair, key0);
            lo_a = r.x;
            hi_a = r.y;
            lo_b = r.z;
            hi_b = r.w;
        } else {
            const Philox4 ra = philox4x32_10(seed_a >> 1, act, 0u, kTagPair, key0);
            const Philox4 rb = philox4x32_10(seed_b >> 1, act, 0u, kTagPair, key0);
            const bool odd_a = seed_a & 1u, odd_b = seed_b & 1u;
            lo_a = odd_a ? ra.z : ra.x;
            hi_a = odd_a ? ra.w : ra.y;
            lo_b = odd_b ? rb.z : rb.x;
            hi_b = odd_b ? rb.w : rb.y;
        }
        ua = uniform52(lo_a, hi_a);
        ub = uniform52(lo_b, hi_b);
    }
    // empirical tables: pool block = [guide u32 x 2^g][cp f64 x len][values f64 x len]; the record
    // carries the byte offsets of guide (tab_off) and cp (dist)
    const uint32_t g = (meta >> 24) & 31u, len8 = (meta & 0x7FFFFFu) * 8u;
    const bool scan = meta & 0x800000u;
    const typename Mem<SMEM>::ptr guide_b = tab + tab_off;
    const typename Mem<SMEM>::ptr cp_b = tab + dist;
    double va, vb;
    emp_value2<SMEM>(guide_b, cp_b, g, len8, scan, hi_a, ua, hi_b, ub, va, vb);
    if (kind == MCDP_DIST_EMP_ABS) {  // _core.cpp:125
        ea = va;
        eb = vb;
    } else {  // _core.cpp:140
        ea = __dmul_rn(va, base);
        eb = __dmul_rn(vb, base);
    }
}

}  // namespace mcdp
