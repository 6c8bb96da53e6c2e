// mcdp_sampling4.cuh -- the generator contract of mcdp_sampling.cuh evaluated for the FOUR adjacent
// samples a lane owns in the quad sweep (mcdp_quad_sweep.cuh).
//
// Nothing about the contract changes: every (seed, activity) pair reads the same Philox blocks and
// goes through the same transforms as in sample_extra2, so the two kernels return identical bits
// (tests/test_gpu_parity.py::test_quad_matches_pair).  What changes is the cost structure: the kind
// dispatch, the distribution-record loads and the Erlang variant switch are taken once per four
// samples, and four independent dependency chains are in flight per lane.
//
// Replaces Dist::sample (reference _core.cpp:72-141) like mcdp_sampling.cuh does.
#pragma once
#include "mcdp_sampling.cuh"

namespace mcdp {

// `quad`: the lane's seeds are s, s+1, s+2, s+3 with s even, i.e. two aligned PAIR blocks (s >> 1 and
// (s + 2) >> 1) serve all four samples; `quad4`: additionally s is a multiple of four, so ONE QUAD block
// (s >> 2) serves all four.
struct Seeds4 {
    uint32_t s[4];
    bool quad, quad4;
};

// Cold paths -- seeds that are not an aligned run of four, redraws after a truncation -- exist in two builds: inlined,
// or behind out-of-line calls (OOL).  Measured on one B200 (ms, inlined / out of line): C3 full 25.45 / 25.72, C2 full
// 31.6 / 32.5, C3-MT 29.3 / 29.8, C3 reduced 28.7 / 28.1, C5 reduced 170.0 / 170.5: the calls cost the full-output
// kernels their register allocation, the smaller code helps the reduced kernels, whose close (statistics) is large.
// The sweep kernels pick per mode.
template <bool OOL>
__device__ __forceinline__ Philox4 philox_cold(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& key0) {
    if constexpr (OOL) return philox4x32_10_call(c0, c1, c2, c3, &key0);
    else return philox4x32_10(c0, c1, c2, c3, key0);
}

// one 32-bit draw per sample from QUAD-style blocks with tag `tag`, draw index j[i]
template <bool OOL>
__device__ __forceinline__ void draw32x4(const Seeds4& sd, bool same_j, const uint32_t (&j)[4], uint32_t act, uint32_t tag,
                                         const PhiloxKeys& key0, uint32_t (&w)[4]) {
    if (sd.quad4 && same_j) {
        const Philox4 r = philox4x32_10(sd.s[0] >> 2, act, j[0], tag, key0);
        w[0] = r.x;
        w[1] = r.y;
        w[2] = r.z;
        w[3] = r.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = philox_word(philox_cold<OOL>(sd.s[i] >> 2, act, j[i], tag, key0), sd.s[i] & 3u);
    }
}

// one 64-bit draw per sample from PAIR-style blocks with tag `tag`, draw index j[i]
template <bool OOL>
__device__ __forceinline__ void draw64x4(const Seeds4& sd, bool same_j, const uint32_t (&j)[4], uint32_t act, uint32_t tag,
                                         const PhiloxKeys& key0, uint32_t (&lo)[4], uint32_t (&hi)[4]) {
    if (sd.quad && same_j) {
        // (s[2] >> 1, not (s[0] >> 1) + 1: the seeds wrap from 0xFFFFFFFF to 0 inside a quad that starts at -2)
        const Philox4 r0 = philox4x32_10(sd.s[0] >> 1, act, j[0], tag, key0);
        const Philox4 r1 = philox4x32_10(sd.s[2] >> 1, act, j[2], tag, key0);
        lo[0] = r0.x;
        hi[0] = r0.y;
        lo[1] = r0.z;
        hi[1] = r0.w;
        lo[2] = r1.x;
        hi[2] = r1.y;
        lo[3] = r1.z;
        hi[3] = r1.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const Philox4 r = philox_cold<OOL>(sd.s[i] >> 1, act, j[i], tag, key0);
            const bool odd = sd.s[i] & 1u;
            lo[i] = odd ? r.z : r.x;
            hi[i] = odd ? r.w : r.y;
        }
    }
}

// The same lookup for draws from one 32-bit word: the boundaries are compared as integers (thr[i] = floor(cp[i] 2^32
// - 1/2): w > thr[i] <=> cp[i] < (w + 1/2) 2^-32, exactly), the uniform is never formed.
template <bool SMEM>
__device__ __forceinline__ void emp_value4_narrow(typename Mem<SMEM>::ptr guide_b, typename Mem<SMEM>::ptr cp_b, uint32_t g,
                                                  uint32_t len8, bool scan, const uint32_t (&w)[4], double (&v)[4]) {
    const typename Mem<SMEM>::ptr thr_b = cp_b + 2u * len8;
    uint32_t idx[4], t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) idx[i] = Mem<SMEM>::u32(guide_b + (w[i] >> (32u - g)) * 4u);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = Mem<SMEM>::u32(thr_b + idx[i] * 4u);
#pragma unroll
    for (int i = 0; i < 4; ++i) idx[i] += w[i] > t[i] ? 1u : 0u;  // thr[len-1] == 2^32 - 1: stays in range
    if (scan) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            while (w[i] > Mem<SMEM>::u32(thr_b + idx[i] * 4u)) ++idx[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = Mem<SMEM>::f64(cp_b + idx[i] * 8u + len8);
}

// inverse-CDF lookup of four samples in lockstep (emp_value2 widened)
template <bool SMEM>
__device__ __forceinline__ void emp_value4(typename Mem<SMEM>::ptr guide_b, typename Mem<SMEM>::ptr cp_b, uint32_t g,
                                           uint32_t len8, bool scan, const uint32_t (&hi)[4], const double (&u)[4],
                                           double (&v)[4]) {
    uint32_t off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) off[i] = Mem<SMEM>::u32(guide_b + (hi[i] >> (32u - g)) * 4u) * 8u;
    double c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = Mem<SMEM>::f64(cp_b + off[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) off[i] += c[i] < u[i] ? 8u : 0u;  // cp[len-1] == 1.0 > u: stays in range
    if (scan) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            while (Mem<SMEM>::f64(cp_b + off[i]) < u[i]) off[i] += 8u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = Mem<SMEM>::f64(cp_b + off[i] + len8);
}

// x[0] > m || ... || x[3] > m as four compares and a predicate OR.  Written in PTX because the C form is folded into
// max(x0..x3) > m, and fp64 max is a three-instruction NaN-propagating sequence on sm_100 (25 instructions in all).
__device__ __forceinline__ bool any_gt4(const double (&x)[4], double m) {
    uint32_t r;
    asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
        "setp.gt.f64 p0, %1, %5;\n\tsetp.gt.f64 p1, %2, %5;\n\tsetp.gt.f64 p2, %3, %5;\n\tsetp.gt.f64 p3, %4, %5;\n\t"
        "or.pred p0, p0, p1;\n\tor.pred p2, p2, p3;\n\tor.pred p0, p0, p2;\n\t"
        "selp.u32 %0, 1, 0, p0;\n\t}"
        : "=r"(r)
        : "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]), "d"(m));
    return r != 0u;
}

// erlang_value for four samples: ONE switch on the warp-uniform variant
template <bool SMEM>
__device__ __forceinline__ void erlang_value4(const DistView<SMEM>& d, int variant, const uint32_t (&w0)[4],
                                              const uint32_t (&w1)[4], const uint32_t (&w2)[4], const uint32_t (&w3)[4],
                                              uint32_t log_tab, double (&y)[4]) {
    double e[4];
    switch (variant) {
        case 1:
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = half_term(w0[i], w1[i], log_tab);
            break;
        case 2:
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = -log_pos(uniform32(w0[i]), log_tab);
            break;
        case 3:
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = half_term(w1[i], w2[i], log_tab) - log_pos(uniform32(w0[i]), log_tab);
            break;
        case 4:
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = -log_pos(uniform32(w0[i]) * uniform32(w1[i]), log_tab);
            break;
        case 5:
#pragma unroll
            for (int i = 0; i < 4; ++i)
                e[i] = half_term(w2[i], w3[i], log_tab) - log_pos(uniform32(w0[i]) * uniform32(w1[i]), log_tab);
            break;
        case 6:
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = -log_pos(uniform32(w0[i]) * uniform32(w1[i]) * uniform32(w2[i]), log_tab);
            break;
        default:
#pragma unroll
            for (int i = 0; i < 4; ++i)
                e[i] = -log_pos((uniform32(w0[i]) * uniform32(w1[i])) * (uniform32(w2[i]) * uniform32(w3[i])), log_tab);
            break;
    }
    const double scale = d.p(1);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = scale * e[i];
}

// draw j[i] of the four samples (erlang_draw2 widened; same blocks and word order)
template <bool SMEM, bool OOL>
__device__ __forceinline__ void erlang_draw4(const DistView<SMEM>& d, int variant, const Seeds4& sd, bool same_j,
                                             const uint32_t (&j)[4], uint32_t act, const PhiloxKeys& key0, uint32_t log_tab,
                                             double (&y)[4]) {
    uint32_t w0[4], w1[4], w2[4] = {0u, 0u, 0u, 0u}, w3[4] = {0u, 0u, 0u, 0u};
    if (variant == 2) {  // one 32-bit uniform per sample: a QUAD-style block
        draw32x4<OOL>(sd, same_j, j, act, kTagErlang, key0, w0);
#pragma unroll
        for (int i = 0; i < 4; ++i) w1[i] = 0u;
    } else if (variant == 1 || variant == 4) {  // 64 bits per sample: PAIR-style blocks
        draw64x4<OOL>(sd, same_j, j, act, kTagErlang, key0, w0, w1);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const Philox4 r = philox4x32_10(sd.s[i], act, j[i], kTagErlang, key0);
            w0[i] = r.x;
            w1[i] = r.y;
            w2[i] = r.z;
            w3[i] = r.w;
        }
    }
    erlang_value4<SMEM>(d, variant, w0, w1, w2, w3, log_tab, y);
}

// Truncation redraws of ONE sample (_core.cpp:98-104: draws 1, 2, ... until x <= max_scale), out of line: rare, and
// inlined (a second copy of every variant body and of the block generators) it was the largest cold region of the loop.
// Same blocks and word order as erlang_draw4 / erlang_draw2 use for draw j of this seed.
template <bool SMEM>
__device__ __noinline__ double erlang_redraw1(typename Mem<SMEM>::ptr rec, uint32_t seed, uint32_t act, const PhiloxKeys* key0,
                                              uint32_t log_tab) {
    const DistView<SMEM> d{rec};
    const int variant = 2 * d.pad0() + d.pad1();
    const double mx = d.p(2);
    double x = 0.0;
    for (uint32_t j = 1u; j < kGammaMaxAttempts; ++j) {
        uint32_t w0, w1 = 0u, w2 = 0u, w3 = 0u;
        if (variant == 2) {
            w0 = philox_word(philox4x32_10(seed >> 2, act, j, kTagErlang, *key0), seed & 3u);
        } else if (variant == 1 || variant == 4) {
            const Philox4 r = philox4x32_10(seed >> 1, act, j, kTagErlang, *key0);
            const bool odd = seed & 1u;
            w0 = odd ? r.z : r.x;
            w1 = odd ? r.w : r.y;
        } else {
            const Philox4 r = philox4x32_10(seed, act, j, kTagErlang, *key0);
            w0 = r.x;
            w1 = r.y;
            w2 = r.z;
            w3 = r.w;
        }
        x = erlang_value<SMEM>(d, variant, w0, w1, w2, w3, log_tab);
        if (!(x > mx)) return x;
    }
    return x > mx ? mx : x;  // only after the attempt cap
}

template <bool SMEM, bool OOL>
__device__ __forceinline__ void erlang_variate4(const DistView<SMEM>& d, const Seeds4& sd, uint32_t act,
                                                const PhiloxKeys& key0, uint32_t log_tab, double (&x)[4]) {
    const int variant = 2 * d.pad0() + d.pad1();
    const double mx = d.p(2);
    uint32_t j[4] = {0u, 0u, 0u, 0u};
    erlang_draw4<SMEM, OOL>(d, variant, sd, true, j, act, key0, log_tab, x);
    // truncation (_core.cpp:98-104): draws 1, 2, ... until x <= max_scale; rare, so off the straight path
    if constexpr (OOL) {
        if (__builtin_expect(any_gt4(x, mx), 0)) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x[i] > mx) x[i] = erlang_redraw1<SMEM>(d.base, sd.s[i], act, &key0, log_tab);
        }
    } else if (__any_sync(0xFFFFFFFFu, any_gt4(x, mx))) {
        bool need[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) need[i] = x[i] > mx;
        do {
#pragma unroll
            for (int i = 0; i < 4; ++i) j[i] += need[i] ? 1u : 0u;
            double y[4];
            erlang_draw4<SMEM, OOL>(d, variant, sd, j[0] == j[1] && j[1] == j[2] && j[2] == j[3], j, act, key0, log_tab, y);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (need[i]) {
                    x[i] = y[i];
                    need[i] = y[i] > mx && j[i] + 1u < kGammaMaxAttempts;
                }
            }
        } while (__any_sync(0xFFFFFFFFu, need[0] || need[1] || need[2] || need[3]));
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = x[i] > mx ? mx : x[i];  // only after the attempt cap
    }
}

// Marsaglia-Tsang for four samples: the first attempts come out of the two GAM0 blocks of the lane's seed pairs (one
// Box-Muller pair each: cos branch even seed, sin branch odd seed) and run in lockstep; then ONE warp-wide retry loop
// in which every lane retries its first pending sample from that sample's SOLO blocks (gamma_variate2 widened).
template <bool SMEM, bool OOL>
__device__ __forceinline__ void gamma_variate4(const DistView<SMEM>& d, const Seeds4& sd, uint32_t act,
                                               const PhiloxKeys& key0, double (&x)[4]) {
    float nf[4];
    uint32_t wacc[4], wboost[4] = {0u, 0u, 0u, 0u};
    if (sd.quad) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const Philox4 blk = philox4x32_10(sd.s[2 * h] >> 1, act, 0u, kTagGam0, key0);
            float r, ang;
            box_muller_polar(blk.x, blk.y, r, ang);
            nf[2 * h] = r * cos_approx(ang);
            nf[2 * h + 1] = r * sin_approx(ang);
            wacc[2 * h] = blk.z;
            wacc[2 * h + 1] = blk.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) gam0_take(philox_cold<OOL>(sd.s[i] >> 1, act, 0u, kTagGam0, key0), sd.s[i], nf[i], wacc[i]);
    }
    if (d.flags() & 1) {
        const uint32_t j0[4] = {0u, 0u, 0u, 0u};
        draw32x4<OOL>(sd, true, j0, act, kTagGbst, key0, wboost);
    }
    bool need[4];
    gamma_first_attempts<SMEM, 4>(d, nf, wacc, wboost, x, need);
#pragma unroll
    for (int i = 0; i < 4; ++i) need[i] = !need[i];
    uint32_t t[4] = {1u, 1u, 1u, 1u};
    while (__any_sync(0xFFFFFFFFu, need[0] || need[1] || need[2] || need[3])) {
        const int idx = need[0] ? 0 : (need[1] ? 1 : (need[2] ? 2 : 3));
        const uint32_t seed = idx == 0 ? sd.s[0] : (idx == 1 ? sd.s[1] : (idx == 2 ? sd.s[2] : sd.s[3]));
        const uint32_t tt = idx == 0 ? t[0] : (idx == 1 ? t[1] : (idx == 2 ? t[2] : t[3]));
        double xx;
        const bool ok = gamma_eval<SMEM>(d, philox4x32_10(seed, act, tt, kTagSolo, key0), xx);
        const bool give_up = tt + 1u >= kGammaMaxAttempts;  // the reference would spin forever: clamp
        if (give_up) xx = fmin(xx, d.p(2));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (idx == i && need[i]) {
                ++t[i];
                if (ok || give_up) {
                    x[i] = xx;
                    need[i] = false;
                }
            }
        }
    }
}

// Extra delays of one activity for the four samples of a lane (sample_extra2 widened).
template <bool SMEM, bool OOL>
__device__ __forceinline__ void sample_extra4(uint32_t meta, uint32_t tab_off, typename Mem<SMEM>::ptr dists, uint32_t dist,
                                              typename Mem<SMEM>::ptr tab, double base, uint32_t act, const Seeds4& sd,
                                              const PhiloxKeys& key0, uint32_t log_tab, double (&e)[4]) {
    const uint32_t kind = meta >> 29;
    if (kind == MCDP_DIST_CONSTANT) {
        const DistView<SMEM> d{dists + dist * uint32_t(sizeof(DistRec))};
        const double c = __dmul_rn(base, d.p(0));  // _core.cpp:75
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = c;
        return;
    }
    if (kind == MCDP_DIST_GAMMA) {
        const DistView<SMEM> d{dists + dist * uint32_t(sizeof(DistRec))};
        double x[4];
        if (d.flags() & 8)
            erlang_variate4<SMEM, OOL>(d, sd, act, key0, log_tab, x);
        else
            gamma_variate4<SMEM, OOL>(d, sd, act, key0, x);
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = __dmul_rn(x[i], base);
        return;
    }
    uint32_t hi[4];
    double u[4];
    const uint32_t j0[4] = {0u, 0u, 0u, 0u};
    if (kind == MCDP_DIST_EXPONENTIAL) {
        // inverse CDF of the exponential truncated to [0, max_scale] (see sample_extra2; _core.cpp:83-89): ONE QUAD
        // block for the lane's four samples; a sample in the top 2^-20 is refined out of line (exp_tail)
        const DistView<SMEM> d{dists + dist * uint32_t(sizeof(DistRec))};
        draw32x4<OOL>(sd, true, j0, act, kTagQuad, key0, hi);
        const double lam = d.p(0), mx = d.p(1), F = d.p(2);
        double x[4];
        if (d.flags() & 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = lam * neg_log1m(uniform32(hi[i]) * F, true, log_tab);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = lam * neg_log1m(uniform32(hi[i]) * F, false, log_tab);
// (the parameters are re-read from the record INSIDE the out-of-line function: passing lambda / F / 1 - F as
            // arguments kept six more registers live across the call site and cost C3 2.6 %; measured on one B200, C3 /
            // C2 full, ms: no refinement 25.5 / 33.8, arguments 26.2 / 32.6, record pointer 25.3 / 31.6)
            if (__builtin_expect(max(max(hi[0], hi[1]), max(hi[2], hi[3])) >= kExpTailWord, 0)) {  // one lane-quad in 2^18
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (hi[i] >= kExpTailWord) x[i] = exp_tail_rec<SMEM>(d.base, hi[i], sd.s[i], act, &key0, log_tab);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[i] = x[i] > mx ? mx : x[i];
            e[i] = __dmul_rn(x[i], base);
        }
        return;
    }
    // table lookups: 32 bits from ONE QUAD block for the whole lane-quad (tables of <= 4096 entries), else 64 bits
    // from two PAIR blocks
    // empirical tables (pool block layout: see sample_extra2)
    const uint32_t g = (meta >> 24) & 31u, len8 = (meta & 0x7FFFFFu) * 8u;
    const bool scan = meta & 0x800000u;
    const typename Mem<SMEM>::ptr guide_b = tab + tab_off;
    const typename Mem<SMEM>::ptr cp_b = tab + dist;
    double v[4];
    if ((meta & 0x7FFFFFu) <= kQuadTableMaxLen) {
        draw32x4<OOL>(sd, true, j0, act, kTagQuad, key0, hi);
        emp_value4_narrow<SMEM>(guide_b, cp_b, g, len8, scan, hi, v);
    } else {
        uint32_t lo[4];
        draw64x4<OOL>(sd, true, j0, act, kTagPair, key0, lo, hi);
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = uniform52(lo[i], hi[i]);
        emp_value4<SMEM>(guide_b, cp_b, g, len8, scan, hi, u, v);
    }
    if (kind == MCDP_DIST_EMP_ABS) {  // _core.cpp:125
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = v[i];
    } else {  // _core.cpp:140
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = __dmul_rn(v[i], base);
    }
}

}  // namespace mcdp
