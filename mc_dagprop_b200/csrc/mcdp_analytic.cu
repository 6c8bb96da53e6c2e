// mcdp_analytic.cu -- the analytic (PMF) propagator on sm_100a: SURVEY.md section 8(f) rank 4.
//
// Replaces the pure-numpy engine of the reference:
//   DiscretePMF.convolve / maximum        analytic/_pmf.py:107-148
//   AnalyticPropagator.run                analytic/_propagator.py:89-148
//   _convert_to_simulated_event (bounds)  analytic/_propagator.py:158-265
// Every event's arrival-time distribution is a probability mass function on the integer grid of `step` seconds.
// An event with predecessors takes  max_i (PMF(src_i) (*) PMF(activity_i)),  clipped to its [earliest, latest]
// window with the context's underflow / overflow rules.
//
// Mapping: events of one topological level are independent -> one launch per level, one CTA per event -- or, when the
// level is narrower than the machine, a thread-block cluster of 2, 4 or 8 CTAs per event that shares the event's
// convolutions.  Inside a CTA the three operations are block-cooperative:
//   * convolution: thread k owns output bins k, k + 256, ...; each bin is a compensated dot product (Dot2: TwoProd via
//     FMA + TwoSum, the rounding errors summed separately, added once at the end) -- the reference accumulates in
//     80-bit np.longdouble and casts to float64; both are within an ulp of the exactly rounded sum;
//   * maximum of two independent variables: P(max = x) = p_a(x) F_b(x) + p_b(x) F_a(x - 1) on the union grid, the two
//     CDFs by one block-wide double-double prefix-sum pass (per-thread segments + a shuffle scan of the segment totals);
//   * the mass corrections of the reference (`_rescale`, the clip-and-normalise of `_convert_to_simulated_event`)
//     are block reductions in double-double followed by the same compare-and-scale steps; the mass of a PMF is
//     computed where it is produced and travels with it.
// Inputs and results live in global memory (L2-resident at these sizes); the intermediates of an event -- seven
// arrays as long as the widest window, sized by the host from the event bounds -- in the CTA's shared memory when
// they fit, else in a global slab.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mcdp_b200.h"

int32_t mcdp_set_error(int32_t code, const std::string& msg);  // mcdp_capi.cu: the calling thread's last error

namespace {

constexpr int kThreads = 256;

// thread-block cluster: rank of this CTA, CTAs per cluster, barrier with release / acquire ordering of global writes
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- double-double helpers (error-free transformations; explicit intrinsics so that nothing is contracted) ----
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd two_sum(double a, double b) {
    const double s = __dadd_rn(a, b);
    const double bb = __dadd_rn(s, -a);
    const double e = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
    return dd{s, e};
}
__device__ __forceinline__ dd quick_two_sum(double a, double b) {  // |a| >= |b|
    const double s = __dadd_rn(a, b);
    return dd{s, __dadd_rn(b, -__dadd_rn(s, -a))};
}
__device__ __forceinline__ dd dd_add(dd x, dd y) {
    dd s = two_sum(x.hi, y.hi);
    s.lo = __dadd_rn(s.lo, __dadd_rn(x.lo, y.lo));
    return quick_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_add_d(dd x, double y) {
    dd s = two_sum(x.hi, y);
    s.lo = __dadd_rn(s.lo, x.lo);
    return quick_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_mul_d(dd x, double y) {  // (x.hi + x.lo) * y
    const double p = __dmul_rn(x.hi, y);
    const double e = __fma_rn(x.hi, y, -p);
    return quick_two_sum(p, __fma_rn(x.lo, y, e));
}
__device__ __forceinline__ double dd_round(dd x) { return __dadd_rn(x.hi, x.lo); }

__device__ __forceinline__ dd warp_sum_dd(dd v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dd w;
        w.hi = __shfl_xor_sync(0xFFFFFFFFu, v.hi, o);
        w.lo = __shfl_xor_sync(0xFFFFFFFFu, v.lo, o);
        v = dd_add(v, w);
    }
    return v;
}
// sum of p[0..n) over the block in double-double; every thread gets the rounded total
__device__ double block_sum(const double* p, int n, dd* s_part) {
    dd acc{0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += kThreads) acc = dd_add_d(acc, p[i]);
    acc = warp_sum_dd(acc);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    dd tot{0.0, 0.0};
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) tot = dd_add(tot, s_part[w]);
    return dd_round(tot);
}
// two sums in one pass (the same additions in the same order as two block_sum calls, one set of barriers)
__device__ void block_sum2(const double* p0, int n0, const double* p1, int n1, dd* s_part, double* r0, double* r1) {
    dd a0{0.0, 0.0}, a1{0.0, 0.0};
    for (int i = threadIdx.x; i < n0; i += kThreads) a0 = dd_add_d(a0, p0[i]);
    for (int i = threadIdx.x; i < n1; i += kThreads) a1 = dd_add_d(a1, p1[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dd w0, w1;
        w0.hi = __shfl_xor_sync(0xFFFFFFFFu, a0.hi, o);
        w0.lo = __shfl_xor_sync(0xFFFFFFFFu, a0.lo, o);
        w1.hi = __shfl_xor_sync(0xFFFFFFFFu, a1.hi, o);
        w1.lo = __shfl_xor_sync(0xFFFFFFFFu, a1.lo, o);
        a0 = dd_add(a0, w0);
        a1 = dd_add(a1, w1);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        s_part[threadIdx.x >> 5] = a0;
        s_part[kThreads / 32 + (threadIdx.x >> 5)] = a1;
    }
    __syncthreads();
    dd t0{0.0, 0.0}, t1{0.0, 0.0};
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        t0 = dd_add(t0, s_part[w]);
        t1 = dd_add(t1, s_part[kThreads / 32 + w]);
    }
    *r0 = dd_round(t0);
    *r1 = dd_round(t1);
}
// numpy.isclose(a, b, rtol, atol) for finite values
__device__ __forceinline__ bool is_close(double a, double b, double rtol, double atol) { return fabs(a - b) <= atol + rtol * fabs(b); }

struct Pmf {
    long long start;  // value of bin 0, seconds
    int len;
    const double* p;
    double mass;      // block_sum(p, len), or negative: not computed yet
};
// masses are computed once where a PMF is produced and travel with it (the reference recomputes total_mass in every
// operation, _pmf.py:96-105: the same number)
__device__ __forceinline__ double pmf_mass(const Pmf& x, dd* s_part) { return x.mass >= 0.0 ? x.mass : block_sum(x.p, x.len, s_part); }

// expected mass of a binary operation (reference _pmf.py:96-105) and the rescale that follows it (:80-94)
__device__ double expected_mass(double m1, double m2) {
    return (is_close(m1, 1.0, 1e-12, 1e-15) && is_close(m2, 1.0, 1e-12, 1e-15)) ? 1.0 : m1 * m2;
}
__device__ double rescale(double* p, int n, double expected, dd* s_part) {  // returns the mass of p afterwards
    double mass = block_sum(p, n, s_part);
    if (mass > 0.0 && !is_close(mass, expected, 1e-12, 1e-15)) {
        const double f = expected / mass;
        for (int i = threadIdx.x; i < n; i += kThreads) p[i] *= f;
        __syncthreads();
        mass = block_sum(p, n, s_part);
    }
    __syncthreads();
    return mass;
}

// One output bin of a (*) b: a compensated dot product (Ogita, Rump, Oishi: Dot2).  Every product and every
// running sum is split into its rounded value and its exact rounding error (TwoProd by FMA, TwoSum); the errors are
// collected in a second plain sum and added at the end, which gives the result as if computed in twice the working
// precision and rounded once -- the reference accumulates in 80-bit np.longdouble (_pmf.py:113-118).  Ten fp64
// operations per term, and the loop-carried chains are one addition each, so four interleaved partial sums keep the
// fp64 pipe busy (the convolutions are what an event's time consists of, and they run at the pipe's throughput).
__device__ __forceinline__ void dot2_step(double& sum, double& err, double x, double y) {
    const double p = __dmul_rn(x, y);
    const double ep = __fma_rn(x, y, -p);
    const dd s = two_sum(sum, p);
    sum = s.hi;
    err = __dadd_rn(err, __dadd_rn(ep, s.lo));
}
__device__ __forceinline__ double convolve_bin(const Pmf& a, const Pmf& b, int k) {
    const int j0 = max(0, k - (b.len - 1)), j1 = min(k, a.len - 1);
    double sum[4] = {0.0, 0.0, 0.0, 0.0}, err[4] = {0.0, 0.0, 0.0, 0.0};
    int j = j0;
    for (; j + 3 <= j1; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) dot2_step(sum[u], err[u], a.p[j + u], b.p[k - j - u]);
    }
    for (int u = 0; j <= j1; ++j, ++u) dot2_step(sum[u], err[u], a.p[j], b.p[k - j]);
    dd t = two_sum(sum[0], sum[1]);
    double e = __dadd_rn(__dadd_rn(err[0], err[1]), t.lo);
    const dd t2 = two_sum(sum[2], sum[3]);
    e = __dadd_rn(e, __dadd_rn(__dadd_rn(err[2], err[3]), t2.lo));
    const dd t3 = two_sum(t.hi, t2.hi);
    return __dadd_rn(t3.hi, __dadd_rn(e, t3.lo));
}

// out = a (*) b  (reference DiscretePMF.convolve, _pmf.py:107-125); `a` is `self`.
// `shared_bins` != nullptr: the CTAs of a thread-block cluster work on the same event.  The dot products -- the fp64
// bulk of an event -- are dealt out bin by bin over the cluster's ranks into `shared_bins` (global memory);
// after a cluster barrier every CTA copies all of them into its own `out` and carries on alone, so everything after
// the convolution is computed redundantly, bit-identically, by every CTA of the cluster.
__device__ double block_convolve(const Pmf& a, const Pmf& b, long long step, double* out, long long* o_start, int* o_len, dd* s_part,
                                 double* shared_bins = nullptr) {  // returns the mass of the result
    const double ma = pmf_mass(a, s_part);
    const double mb = pmf_mass(b, s_part);
    int n;
    if (a.len == 1) {
        n = b.len;
        const double w = a.p[0];
        for (int k = threadIdx.x; k < n; k += kThreads) out[k] = b.p[k] * w;
    } else if (b.len == 1) {
        n = a.len;
        const double w = b.p[0];
        for (int k = threadIdx.x; k < n; k += kThreads) out[k] = a.p[k] * w;
    } else if (shared_bins) {
        n = a.len + b.len - 1;
        // bins dealt out one by one: the long dot products sit in the middle of the result, every rank gets its share
        const int rank = int(cluster_rank()), csize = int(cluster_size());
        for (int k = rank + csize * int(threadIdx.x); k < n; k += csize * kThreads) shared_bins[k] = convolve_bin(a, b, k);
        cluster_sync();
        for (int k = threadIdx.x; k < n; k += kThreads) out[k] = __ldcg(shared_bins + k);
    } else {
        n = a.len + b.len - 1;
        for (int k = threadIdx.x; k < n; k += kThreads) out[k] = convolve_bin(a, b, k);
    }
    *o_start = a.start + b.start;
    *o_len = n;
    __syncthreads();
    (void)step;
    return rescale(out, n, expected_mass(ma, mb), s_part);
}

// inclusive prefix sums, as double-double, of TWO functions on a grid of n bins in one pass (one set of barriers, two
// independent dependency chains): x(i) = src[i - first] for first <= i < first + len, else 0
struct Window {
    const double* src;
    int first, len;
    double *c_hi, *c_lo;
    __device__ __forceinline__ double at(int i) const { return (i >= first && i < first + len) ? src[i - first] : 0.0; }
};
__device__ void block_cumsum2(int n, const Window& A, const Window& B, dd* s_seg) {
    const int per = (n + kThreads - 1) / kThreads;
    const int i0 = min(n, int(threadIdx.x) * per), i1 = min(n, i0 + per);
    dd acc_a{0.0, 0.0}, acc_b{0.0, 0.0};
    for (int i = i0; i < i1; ++i) {
        acc_a = dd_add_d(acc_a, A.at(i));
        acc_b = dd_add_d(acc_b, B.at(i));
        A.c_hi[i] = acc_a.hi;
        A.c_lo[i] = acc_a.lo;
        B.c_hi[i] = acc_b.hi;
        B.c_lo[i] = acc_b.lo;
    }
    // exclusive scan of the 256 segment totals: shuffles inside a warp, then the eight warp totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    dd inc_a = acc_a, inc_b = acc_b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        dd wa, wb;
        wa.hi = __shfl_up_sync(0xFFFFFFFFu, inc_a.hi, o);
        wa.lo = __shfl_up_sync(0xFFFFFFFFu, inc_a.lo, o);
        wb.hi = __shfl_up_sync(0xFFFFFFFFu, inc_b.hi, o);
        wb.lo = __shfl_up_sync(0xFFFFFFFFu, inc_b.lo, o);
        if (lane >= o) {
            inc_a = dd_add(wa, inc_a);
            inc_b = dd_add(wb, inc_b);
        }
    }
    dd exc_a, exc_b;
    exc_a.hi = __shfl_up_sync(0xFFFFFFFFu, inc_a.hi, 1);
    exc_a.lo = __shfl_up_sync(0xFFFFFFFFu, inc_a.lo, 1);
    exc_b.hi = __shfl_up_sync(0xFFFFFFFFu, inc_b.hi, 1);
    exc_b.lo = __shfl_up_sync(0xFFFFFFFFu, inc_b.lo, 1);
    if (lane == 0) exc_a = exc_b = dd{0.0, 0.0};
    __syncthreads();  // s_seg may still be read by the previous call
    if (lane == 31) {
        s_seg[warp] = inc_a;
        s_seg[kThreads / 32 + warp] = inc_b;
    }
    __syncthreads();
    dd before_a{0.0, 0.0}, before_b{0.0, 0.0};
    for (int w = 0; w < warp; ++w) {
        before_a = dd_add(before_a, s_seg[w]);
        before_b = dd_add(before_b, s_seg[kThreads / 32 + w]);
    }
    const dd off_a = dd_add(before_a, exc_a), off_b = dd_add(before_b, exc_b);
    if (threadIdx.x > 0) {
        for (int i = i0; i < i1; ++i) {
            const dd va = dd_add(off_a, dd{A.c_hi[i], A.c_lo[i]});
            const dd vb = dd_add(off_b, dd{B.c_hi[i], B.c_lo[i]});
            A.c_hi[i] = va.hi;
            A.c_lo[i] = va.lo;
            B.c_hi[i] = vb.hi;
            B.c_lo[i] = vb.lo;
        }
    }
    __syncthreads();
}

// out = max(a, b) of independent variables (reference DiscretePMF.maximum, _pmf.py:127-148); `a` is `self`.
// c[0..3]: four scratch arrays of the union length
__device__ double block_maximum(const Pmf& a, const Pmf& b, long long step, double* out, long long* o_start, int* o_len,
                                double* const* c, dd* s_part, dd* s_seg) {  // returns the mass of the result
    const double ma = pmf_mass(a, s_part);
    const double mb = pmf_mass(b, s_part);
    const long long lo = min(a.start, b.start);
    const long long hi = max(a.start + (long long)(a.len - 1) * step, b.start + (long long)(b.len - 1) * step);
    const int n = int((hi - lo) / step) + 1;
    const int oa = int((a.start - lo) / step), ob = int((b.start - lo) / step);
    auto pa = [&](int i) { return (i >= oa && i < oa + a.len) ? a.p[i - oa] : 0.0; };
    auto pb = [&](int i) { return (i >= ob && i < ob + b.len) ? b.p[i - ob] : 0.0; };
    block_cumsum2(n, Window{a.p, oa, a.len, c[0], c[1]}, Window{b.p, ob, b.len, c[2], c[3]}, s_seg);
    for (int i = threadIdx.x; i < n; i += kThreads) {
        // p_a(i) F_b(i) + p_b(i) F_a(i - 1), in double-double, rounded once
        dd t = dd_mul_d(dd{c[2][i], c[3][i]}, pa(i));
        if (i > 0) t = dd_add(t, dd_mul_d(dd{c[0][i - 1], c[1][i - 1]}, pb(i)));
        out[i] = dd_round(t);
    }
    *o_start = lo;
    *o_len = n;
    __syncthreads();
    return rescale(out, n, expected_mass(ma, mb), s_part);
}

enum { kRuleTruncate = 1, kRuleRemove = 2, kRuleRedistribute = 3 };
enum { kClipOk = 0, kClipNoLowerBin = 1, kClipNoUpperBin = 2, kClipEmpty = 3, kClipBadBounds = 4 };

// Clip r to [min_value, max_value] and apply the flow rules (reference _convert_to_simulated_event,
// _propagator.py:158-265).  `out` has room for the kept bins (at least one).  Returns a kClip* status (block-uniform).
__device__ int block_clip(const Pmf& r, long long step, long long min_value, long long max_value, int under_rule, int over_rule,
                          double* out, long long* o_start, int* o_len, double* o_under, double* o_over, double* o_mass, dd* s_part) {
    if (min_value > max_value) return kClipBadBounds;
    // kept bins: values in [min_value, max_value]; bins below / above are contiguous runs (values ascend)
    long long k0 = min_value - r.start, k1 = max_value - r.start;
    k0 = k0 <= 0 ? 0 : (k0 + step - 1) / step;             // first bin with value >= min_value
    k1 = k1 < 0 ? -1 : k1 / step;                          // last bin with value <= max_value
    const long long len_ll = r.len;
    const int first = int(k0 < len_ll ? k0 : len_ll), last = int(k1 < len_ll - 1 ? k1 : len_ll - 1);
    int n = max(0, last - first + 1);
    double under, over;
    block_sum2(r.p, first, r.p + last + 1, r.len - 1 - last, s_part, &under, &over);
    for (int i = threadIdx.x; i < n; i += kThreads) out[i] = r.p[first + i];
    long long start = r.start + (long long)first * step;
    __syncthreads();
    double redistribute = 0.0;
    if (under_rule == kRuleTruncate && under > 0.0) {
        if (n > 0 && start == min_value) {
            if (threadIdx.x == 0) out[0] += under;
        } else if (n == 0) {
            if (threadIdx.x == 0) out[0] = under;
            n = 1;
            start = min_value;
        } else {
            return kClipNoLowerBin;
        }
        under = 0.0;
    } else if (under_rule == kRuleRedistribute && under > 0.0) {
        redistribute += under;
        under = 0.0;
    }
    __syncthreads();
    if (over_rule == kRuleTruncate && over > 0.0) {
        if (n > 0 && start + (long long)(n - 1) * step == max_value) {
            if (threadIdx.x == 0) out[n - 1] += over;
        } else if (n == 0) {
            if (threadIdx.x == 0) out[0] = over;
            n = 1;
            start = max_value;
        } else {
            return kClipNoUpperBin;
        }
        over = 0.0;
    } else if (over_rule == kRuleRedistribute && over > 0.0) {
        redistribute += over;
        over = 0.0;
    }
    __syncthreads();
    if (redistribute > 0.0) {
        const double inside = block_sum(out, n, s_part);
        if (inside == 0.0) {
            if (threadIdx.x == 0) out[0] = redistribute;
            n = 1;
            start = min_value;
        } else {
            for (int i = threadIdx.x; i < n; i += kThreads) out[i] = out[i] + redistribute * (out[i] / inside);
        }
        __syncthreads();
    }
    if (n == 0) return kClipEmpty;
    const double lost = under + over;
    const double target = fmax(0.0, 1.0 - lost);
    const double inside = block_sum(out, n, s_part);
    if (inside > 0.0) {
        for (int i = threadIdx.x; i < n; i += kThreads) out[i] = out[i] / inside * target;
    } else if (target > 0.0) {
        if (threadIdx.x == 0) out[0] = target;
        n = 1;
        start = min_value;
    }
    __syncthreads();
    double mass = block_sum(out, n, s_part);
    const double total = mass + under + over;
    if (!is_close(total, 1.0, 1e-12, 1e-15) && total > 0.0) {
        const double corr = 1.0 / total;
        for (int i = threadIdx.x; i < n; i += kThreads) out[i] *= corr;
        __syncthreads();
        mass = block_sum(out, n, s_part);
    }
    *o_mass = mass;
    *o_start = start;
    *o_len = n;
    *o_under = under;
    *o_over = over;
    return kClipOk;
}

struct AnalyticParams {
    const int32_t* order;      // events in topological order
    const int64_t* pred_off;   // [E + 1] by event id
    const int32_t* pred_src;   // [P]
    const int32_t* pred_pmf;   // [P] activity PMF index
    const long long* lower;    // [E]
    const long long* upper;    // [E]
    const long long* origin;   // [E]
    const long long* pmf_start;
    const int64_t* pmf_off;    // [n_pmfs + 1]
    const double* pmf_probs;
    const double* pmf_mass;    // [n_pmfs] block_sum of each activity PMF (pmf_mass_kernel)
    double* out_mass;          // [E] mass of each event's result
    long long* out_start;      // [E]
    int32_t* out_len;          // [E]
    const int64_t* out_off;    // [E + 1] slot of each event in out_probs
    double* out_probs;
    double* underflow;
    double* overflow;
    int32_t* status;           // [E]
    double* scratch;           // per CTA: 7 arrays of scratch_len (null: they fit the CTA's dynamic shared memory)
    long long scratch_len;
    double* cluster_bins;      // per cluster: 2 arrays of scratch_len, the bins of a convolution dealt out over its CTAs
    long long step;
    int under_rule, over_rule;
};

// mass of every activity PMF, once per run
__global__ void __launch_bounds__(kThreads) pmf_mass_kernel(const int64_t* pmf_off, const double* pmf_probs, double* pmf_mass) {
    __shared__ dd s_part[2 * (kThreads / 32)];
    const int a = blockIdx.x;
    const double m = block_sum(pmf_probs + pmf_off[a], int(pmf_off[a + 1] - pmf_off[a]), s_part);
    if (threadIdx.x == 0) pmf_mass[a] = m;
}

__global__ void __launch_bounds__(kThreads) analytic_level_kernel(const AnalyticParams p, int pos0) {
    __shared__ dd s_part[2 * (kThreads / 32)];
    __shared__ dd s_seg[kThreads];
    __shared__ long long s_start[2];
    __shared__ int s_len[2];
    // one event per thread-block cluster (a narrow level is launched with 2, 4 or 8 CTAs per event, a wide one with 1)
    const unsigned csize = cluster_size(), slot = blockIdx.x / csize;
    const bool lead = cluster_rank() == 0;
    const int ev = p.order[pos0 + slot];
    const int64_t b = p.pred_off[ev], e = p.pred_off[ev + 1];
    double* const out = p.out_probs + p.out_off[ev];
    if (b == e) {  // origin: a unit mass at the rounded earliest time (_propagator.py:103-110)
        if (threadIdx.x == 0 && lead) {
            out[0] = 1.0;
            p.out_mass[ev] = 1.0;
            p.out_start[ev] = p.origin[ev];
            p.out_len[ev] = 1;
            p.underflow[ev] = 0.0;
            p.overflow[ev] = 0.0;
            p.status[ev] = kClipOk;
        }
        return;
    }
    // intermediates of the event: in shared memory when they fit (every step of the chain below is a round trip
    // through them between two barriers), else in a global scratch slab
    extern __shared__ double s_scratch[];
    double* const base = p.scratch ? p.scratch + size_t(blockIdx.x) * 7 * size_t(p.scratch_len) : s_scratch;
    // two buffers in turn: a CTA that is ahead writes the bins of predecessor k + 1 while another still copies those of k;
    // it cannot reach k + 2 before the barrier of k + 1, which every CTA passes only after its copy of k
    double* const cbins = csize > 1 ? p.cluster_bins + size_t(slot) * 2 * size_t(p.scratch_len) : nullptr;
    double* conv = base;
    double* run = base + p.scratch_len;
    double* tmp = base + 2 * p.scratch_len;
    double* const c[4] = {base + 3 * p.scratch_len, base + 4 * p.scratch_len, base + 5 * p.scratch_len, base + 6 * p.scratch_len};
    Pmf r{0, 0, nullptr, -1.0};
    for (int64_t k = b; k < e; ++k) {
        const int src = p.pred_src[k], a = p.pred_pmf[k];
        const Pmf pred{p.out_start[src], p.out_len[src], p.out_probs + p.out_off[src], p.out_mass[src]};
        const Pmf act{p.pmf_start[a], int(p.pmf_off[a + 1] - p.pmf_off[a]), p.pmf_probs + p.pmf_off[a], p.pmf_mass[a]};
        double* const dst = (k == b) ? run : conv;
        const double cv_mass = block_convolve(pred, act, p.step, dst, &s_start[0], &s_len[0], s_part,  // every thread writes the same values
                                              cbins ? cbins + size_t((k - b) & 1) * size_t(p.scratch_len) : nullptr);
        __syncthreads();
        const Pmf cv{s_start[0], s_len[0], dst, cv_mass};
        if (k == b) {
            r = cv;
        } else {
            const double mx_mass = block_maximum(r, cv, p.step, tmp, &s_start[1], &s_len[1], c, s_part, s_seg);
            __syncthreads();
            r = Pmf{s_start[1], s_len[1], tmp, mx_mass};
            double* t = run;  // the result becomes the running PMF
            run = tmp;
            tmp = t;
        }
        __syncthreads();
    }
    if (!lead) return;  // the event's result is written once (no cluster barrier follows)
    long long o_start = 0;
    int o_len = 0;
    double under = 0.0, over = 0.0, mass = 0.0;
    const int st = block_clip(r, p.step, p.lower[ev], p.upper[ev], p.under_rule, p.over_rule, out, &o_start, &o_len, &under, &over, &mass, s_part);
    if (threadIdx.x == 0) {
        p.status[ev] = st;
        p.out_mass[ev] = mass;
        p.out_start[ev] = o_start;
        p.out_len[ev] = o_len;
        p.underflow[ev] = under;
        p.overflow[ev] = over;
    }
}

// single operations (DiscretePMF.convolve / maximum, AnalyticPropagator._convert_to_simulated_event)
struct OpParams {
    Pmf a, b;
    long long step, min_value, max_value;
    int under_rule, over_rule;
    double* out;        // result bins
    double* scratch;    // 4 arrays of the union length (maximum)
    long long scratch_len;
    long long* o_start;
    int32_t* o_len;
    double* o_flow;     // [2] underflow, overflow (clip)
    int32_t* o_status;
};
__global__ void __launch_bounds__(kThreads) pmf_op_kernel(const OpParams p, int op) {
    __shared__ dd s_part[2 * (kThreads / 32)];
    __shared__ dd s_seg[kThreads];
    long long o_start = 0;
    int o_len = 0, st = kClipOk;
    double under = 0.0, over = 0.0;
    if (op == 0) {
        block_convolve(p.a, p.b, p.step, p.out, &o_start, &o_len, s_part);
    } else if (op == 1) {
        double* const c[4] = {p.scratch, p.scratch + p.scratch_len, p.scratch + 2 * p.scratch_len, p.scratch + 3 * p.scratch_len};
        block_maximum(p.a, p.b, p.step, p.out, &o_start, &o_len, c, s_part, s_seg);
    } else {
        double mass = 0.0;
        st = block_clip(p.a, p.step, p.min_value, p.max_value, p.under_rule, p.over_rule, p.out, &o_start, &o_len, &under, &over, &mass, s_part);
    }
    if (threadIdx.x == 0) {
        *p.o_start = o_start;
        *p.o_len = o_len;
        p.o_flow[0] = under;
        p.o_flow[1] = over;
        *p.o_status = st;
    }
}

#define ACUDA(expr)                                                                                        \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            rc = mcdp_set_error(MCDP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
            goto done;                                                                                     \
        }                                                                                                  \
    } while (0)

const char* clip_message(int st) {
    switch (st) {
        case kClipNoLowerBin: return "Underflow mass cannot be truncated: no lower-bound bin present.";
        case kClipNoUpperBin: return "Overflow mass cannot be truncated: no upper-bound bin present.";
        case kClipEmpty: return "PMF must not be empty after clipping";
        case kClipBadBounds: return "min_value must not exceed max_value";
        default: return "";
    }
}

struct DevMem {
    std::vector<void*> ptrs;
    ~DevMem() {
        for (void* q : ptrs) cudaFree(q);
    }
    template <typename T>
    cudaError_t alloc(T** out, size_t n) {
        void* q = nullptr;
        const cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(q);
        *out = static_cast<T*>(q);
        return e;
    }
    template <typename T>
    cudaError_t upload(T** out, const T* src, size_t n) {
        cudaError_t e = alloc(out, n);
        if (e == cudaSuccess && n) e = cudaMemcpy(*out, src, n * sizeof(T), cudaMemcpyHostToDevice);
        return e;
    }
};

int32_t use_device(int32_t device, int* prev) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return mcdp_set_error(MCDP_ERR_CUDA, "no usable CUDA device: the analytic propagator has no CPU execution path");
    }
    if (device < 0 || device >= n) return mcdp_set_error(MCDP_ERR_ARG, "device ordinal out of range");
    cudaGetDevice(prev);
    if (cudaSetDevice(device) != cudaSuccess) return mcdp_set_error(MCDP_ERR_CUDA, "cudaSetDevice failed");
    return MCDP_OK;
}

}  // namespace

extern "C" {

int64_t mcdp_analytic_out_capacity(const mcdp_analytic_desc* d, int64_t* out_off) {
    if (!d || d->n_events < 0 || d->step <= 0) return -1;
    std::vector<char> has_pred(size_t(d->n_events), 0);
    for (int32_t i = 0; i < d->n_prec_entries; ++i)
        if (d->prec_target[i] >= 0 && d->prec_target[i] < d->n_events && d->prec_off[i + 1] > d->prec_off[i]) has_pred[d->prec_target[i]] = 1;
    int64_t off = 0;
    for (int32_t e = 0; e < d->n_events; ++e) {
        if (out_off) out_off[e] = off;
        const int64_t w = has_pred[e] ? std::max<int64_t>(0, d->upper[e] - d->lower[e]) / d->step + 2 : 1;
        off += w;
    }
    if (out_off) out_off[d->n_events] = off;
    return off;
}

namespace {
thread_local double g_profile[5] = {0, 0, 0, 0, 0};  // last mcdp_analytic_run of this thread, see mcdp_analytic_last_profile
double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
}  // namespace

int32_t mcdp_analytic_last_profile(double* out5) {
    if (!out5) return mcdp_set_error(MCDP_ERR_ARG, "null argument");
    std::copy(g_profile, g_profile + 5, out5);
    return MCDP_OK;
}

int32_t mcdp_analytic_run(const mcdp_analytic_desc* d, int32_t device, int64_t* out_start, int32_t* out_len, int64_t* out_off,
                          double* out_probs, int64_t out_cap, double* underflow, double* overflow) {
    if (!d || !out_start || !out_len || !out_off || !out_probs || !underflow || !overflow)
        return mcdp_set_error(MCDP_ERR_ARG, "null argument");
    if (d->step <= 0) return mcdp_set_error(MCDP_ERR_INVALID, "step_size must be positive");
    const int32_t E = d->n_events;
    if (E < 0 || d->n_prec_entries < 0 || d->n_pmfs < 0) return mcdp_set_error(MCDP_ERR_ARG, "negative element count");
    nvtxRangePushA("mcdp:analytic run");
    struct Pop {
        ~Pop() { nvtxRangePop(); }
    } pop;
    auto t_phase = std::chrono::steady_clock::now();
    std::fill(g_profile, g_profile + 5, 0.0);
    // precedence by event id (the last entry for a target wins, like the Monte-Carlo compile), bounds-checked
    std::vector<int32_t> entry_of(size_t(E), -1);
    for (int32_t i = 0; i < d->n_prec_entries; ++i) {
        const int32_t t = d->prec_target[i];
        if (t < 0 || t >= E) return mcdp_set_error(MCDP_ERR_INVALID, "target index " + std::to_string(t) + " out of range");
        for (int64_t k = d->prec_off[i]; k < d->prec_off[i + 1]; ++k) {
            if (d->pred_src[k] < 0 || d->pred_src[k] >= E)
                return mcdp_set_error(MCDP_ERR_INVALID, "predecessor index " + std::to_string(d->pred_src[k]) + " out of range");
            if (d->pred_pmf[k] < 0 || d->pred_pmf[k] >= d->n_pmfs) return mcdp_set_error(MCDP_ERR_INVALID, "activity PMF index out of range");
        }
        entry_of[t] = i;
    }
    std::vector<int64_t> pred_off(size_t(E) + 1, 0);
    for (int32_t e = 0; e < E; ++e) pred_off[e + 1] = pred_off[e] + (entry_of[e] >= 0 ? d->prec_off[entry_of[e] + 1] - d->prec_off[entry_of[e]] : 0);
    const size_t n_pred = size_t(pred_off[size_t(E)]);
    std::vector<int32_t> pred_src(n_pred, 0);
    std::vector<int32_t> pred_pmf(n_pred, 0);
    std::vector<int32_t> indeg(size_t(E), 0);
    std::vector<std::vector<int32_t>> succ(static_cast<size_t>(E));
    for (int32_t e = 0; e < E; ++e) {
        if (entry_of[e] < 0) continue;
        int64_t o = pred_off[e];
        for (int64_t k = d->prec_off[entry_of[e]]; k < d->prec_off[entry_of[e] + 1]; ++k, ++o) {
            pred_src[o] = d->pred_src[k];
            pred_pmf[o] = d->pred_pmf[k];
            succ[d->pred_src[k]].push_back(e);
            ++indeg[e];
        }
    }
    // levels (Kahn by waves): events of a level depend only on earlier levels
    std::vector<int32_t> order, level_begin{0};
    order.reserve(size_t(E));
    {
        std::vector<int32_t> cur;
        for (int32_t e = 0; e < E; ++e)
            if (indeg[e] == 0) cur.push_back(e);
        while (!cur.empty()) {
            std::vector<int32_t> next;
            for (int32_t e : cur) {
                order.push_back(e);
                for (int32_t s : succ[e])
                    if (--indeg[s] == 0) next.push_back(s);
            }
            level_begin.push_back(int32_t(order.size()));
            cur.swap(next);
        }
        if (int32_t(order.size()) != E) return mcdp_set_error(MCDP_ERR_INVALID, "Invalid DAG: cycle detected");
    }
    // output slots and the widest intermediate window (support of every event lies inside its bounds)
    std::vector<int64_t> off(size_t(E) + 1);
    const int64_t cap = mcdp_analytic_out_capacity(d, off.data());
    if (cap > out_cap) return mcdp_set_error(MCDP_ERR_ARG, "out_probs is smaller than mcdp_analytic_out_capacity");
    int64_t scratch_len = 2;
    int64_t max_width = 1;
    for (size_t l = 0; l + 1 < level_begin.size(); ++l) max_width = std::max<int64_t>(max_width, level_begin[l + 1] - level_begin[l]);
    for (int32_t e = 0; e < E; ++e) {
        if (pred_off[e + 1] == pred_off[e]) continue;
        int64_t lo = INT64_MAX, hi = INT64_MIN;
        for (int64_t k = pred_off[e]; k < pred_off[e + 1]; ++k) {
            const int32_t s = pred_src[k], a = pred_pmf[k];
            const bool s_origin = pred_off[s + 1] == pred_off[s];
            const int64_t s_lo = s_origin ? d->origin[s] : std::min(d->lower[s], d->upper[s]);
            const int64_t s_hi = s_origin ? d->origin[s] : std::max(d->lower[s], d->upper[s]);
            const int64_t a_len = d->pmf_off[a + 1] - d->pmf_off[a];
            if (a_len <= 0) return mcdp_set_error(MCDP_ERR_INVALID, "PMF values cannot be empty");
            lo = std::min(lo, s_lo + d->pmf_start[a]);
            hi = std::max(hi, s_hi + d->pmf_start[a] + (a_len - 1) * d->step);
        }
        scratch_len = std::max(scratch_len, (hi - lo) / d->step + 3);
    }
    int prev = -1;
    int32_t rc = use_device(device, &prev);
    if (rc) return rc;
    g_profile[0] = ms_since(t_phase);  // host: precedence by event, levels, output slots
    t_phase = std::chrono::steady_clock::now();
    {
        DevMem mem;
        AnalyticParams p{};
        int32_t* d_order = nullptr;
        int64_t *d_pred_off = nullptr, *d_pmf_off = nullptr, *d_out_off = nullptr;
        int32_t *d_pred_src = nullptr, *d_pred_pmf = nullptr, *d_out_len = nullptr, *d_status = nullptr;
        long long *d_lower = nullptr, *d_upper = nullptr, *d_origin = nullptr, *d_pmf_start = nullptr, *d_out_start = nullptr;
        double *d_pmf_probs = nullptr, *d_out_probs = nullptr, *d_under = nullptr, *d_over = nullptr, *d_scratch = nullptr;
        double *d_cluster_bins = nullptr, *d_out_mass = nullptr, *d_pmf_mass = nullptr;
        std::vector<int32_t> status(static_cast<size_t>(E));
        static_assert(sizeof(long long) == sizeof(int64_t), "64-bit values");
        ACUDA(mem.upload(&d_order, order.data(), order.size()));
        ACUDA(mem.upload(&d_pred_off, pred_off.data(), pred_off.size()));
        ACUDA(mem.upload(&d_pred_src, pred_src.data(), pred_src.size()));
        ACUDA(mem.upload(&d_pred_pmf, pred_pmf.data(), pred_pmf.size()));
        ACUDA(mem.upload(&d_lower, reinterpret_cast<const long long*>(d->lower), size_t(E)));
        ACUDA(mem.upload(&d_upper, reinterpret_cast<const long long*>(d->upper), size_t(E)));
        ACUDA(mem.upload(&d_origin, reinterpret_cast<const long long*>(d->origin), size_t(E)));
        ACUDA(mem.upload(&d_pmf_start, reinterpret_cast<const long long*>(d->pmf_start), size_t(d->n_pmfs)));
        ACUDA(mem.upload(&d_pmf_off, d->pmf_off, size_t(d->n_pmfs) + 1));
        ACUDA(mem.upload(&d_pmf_probs, d->pmf_probs, size_t(d->pmf_off[d->n_pmfs])));
        ACUDA(mem.upload(&d_out_off, off.data(), off.size()));
        ACUDA(mem.alloc(&d_out_start, size_t(E)));
        ACUDA(mem.alloc(&d_out_len, size_t(E)));
        ACUDA(mem.alloc(&d_status, size_t(E)));
        ACUDA(mem.alloc(&d_out_probs, size_t(cap)));
        ACUDA(mem.alloc(&d_out_mass, size_t(E)));
        ACUDA(mem.alloc(&d_pmf_mass, size_t(std::max(d->n_pmfs, 1))));
        ACUDA(mem.alloc(&d_under, size_t(E)));
        ACUDA(mem.alloc(&d_over, size_t(E)));
        const size_t smem_need = size_t(7) * size_t(scratch_len) * sizeof(double);
        int smem_max = 0;
        ACUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        const bool in_smem = smem_need + 8192 <= size_t(smem_max);  // static shared memory of the kernel is below 8 KB
        int sm_count = 1;
        ACUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
        if (in_smem) {
            ACUDA(cudaFuncSetAttribute(analytic_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_need)));
        } else {  // one slab per CTA of the largest launch (a clustered launch has at most sm_count CTAs)
            ACUDA(mem.alloc(&d_scratch, size_t(std::max<int64_t>(max_width, sm_count)) * 7 * size_t(scratch_len)));
        }
        // levels narrower than the machine: 2, 4 or 8 CTAs per event share its convolutions (at most sm_count / 2 clusters)
        ACUDA(mem.alloc(&d_cluster_bins, size_t(std::max(sm_count / 2, 1)) * 2 * size_t(scratch_len)));
        ACUDA(cudaMemset(d_out_probs, 0, size_t(std::max<int64_t>(cap, 1)) * 8));
        p.order = d_order;
        p.pred_off = d_pred_off;
        p.pred_src = d_pred_src;
        p.pred_pmf = d_pred_pmf;
        p.lower = d_lower;
        p.upper = d_upper;
        p.origin = d_origin;
        p.pmf_start = d_pmf_start;
        p.pmf_off = d_pmf_off;
        p.pmf_probs = d_pmf_probs;
        p.pmf_mass = d_pmf_mass;
        p.out_mass = d_out_mass;
        p.out_start = d_out_start;
        p.out_len = d_out_len;
        p.out_off = d_out_off;
        p.out_probs = d_out_probs;
        p.underflow = d_under;
        p.overflow = d_over;
        p.status = d_status;
        p.scratch = d_scratch;
        p.scratch_len = scratch_len;
        p.cluster_bins = d_cluster_bins;
        p.step = d->step;
        p.under_rule = d->underflow_rule;
        p.over_rule = d->overflow_rule;
        ACUDA(cudaDeviceSynchronize());
        g_profile[1] = ms_since(t_phase);  // device allocations and uploads
        t_phase = std::chrono::steady_clock::now();
        if (d->n_pmfs > 0) {
            pmf_mass_kernel<<<unsigned(d->n_pmfs), kThreads>>>(d_pmf_off, d_pmf_probs, d_pmf_mass);
            ACUDA(cudaGetLastError());
        }
        for (size_t l = 0; l + 1 < level_begin.size(); ++l) {
            const int n = level_begin[l + 1] - level_begin[l];
            if (n <= 0) continue;
            const int csize = n * 8 <= sm_count ? 8 : (n * 4 <= sm_count ? 4 : (n * 2 <= sm_count ? 2 : 1));
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(unsigned(n * csize));
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = in_smem ? smem_need : 0;
            cudaLaunchAttribute attr{};
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = unsigned(csize);
            attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            ACUDA(cudaLaunchKernelEx(&cfg, analytic_level_kernel, p, int(level_begin[l])));
        }
        ACUDA(cudaDeviceSynchronize());
        g_profile[2] = ms_since(t_phase);  // one launch per level, all levels
        g_profile[4] = double(level_begin.size() - 1);
        t_phase = std::chrono::steady_clock::now();
        ACUDA(cudaMemcpy(status.data(), d_status, size_t(E) * 4, cudaMemcpyDeviceToHost));
        for (int32_t e : order) {  // the first failing event in evaluation order, like the reference's loop
            if (status[e] != kClipOk) {
                rc = mcdp_set_error(MCDP_ERR_INVALID, std::string(clip_message(status[e])) + " (event " + std::to_string(e) + ")");
                goto done;
            }
        }
        ACUDA(cudaMemcpy(out_start, d_out_start, size_t(E) * 8, cudaMemcpyDeviceToHost));
        ACUDA(cudaMemcpy(out_len, d_out_len, size_t(E) * 4, cudaMemcpyDeviceToHost));
        ACUDA(cudaMemcpy(out_probs, d_out_probs, size_t(cap) * 8, cudaMemcpyDeviceToHost));
        ACUDA(cudaMemcpy(underflow, d_under, size_t(E) * 8, cudaMemcpyDeviceToHost));
        ACUDA(cudaMemcpy(overflow, d_over, size_t(E) * 8, cudaMemcpyDeviceToHost));
        std::copy(off.begin(), off.end(), out_off);
        g_profile[3] = ms_since(t_phase);  // results to the host
    }
done:
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}

// op: 0 convolve (a = self, b = other), 1 maximum, 2 clip a to [min_value, max_value] with the rules
int32_t mcdp_pmf_op(int32_t op, int32_t device, int64_t step, int64_t a_start, int32_t a_len, const double* a_probs, int64_t b_start,
                    int32_t b_len, const double* b_probs, int64_t min_value, int64_t max_value, int32_t underflow_rule,
                    int32_t overflow_rule, int64_t* out_start, int32_t* out_len, double* out_probs, int64_t out_cap,
                    double* out_underflow, double* out_overflow) {
    if (op < 0 || op > 2) return mcdp_set_error(MCDP_ERR_ARG, "unknown PMF operation");
    if (!a_probs || a_len <= 0 || (op < 2 && (!b_probs || b_len <= 0))) return mcdp_set_error(MCDP_ERR_INVALID, "PMF values cannot be empty");
    if (step <= 0 && !(op == 0 && (a_len == 1 || b_len == 1))) return mcdp_set_error(MCDP_ERR_INVALID, "step must be positive");
    if (!out_start || !out_len || !out_probs) return mcdp_set_error(MCDP_ERR_ARG, "null output");
    const int64_t st = std::max<int64_t>(step, 1);
    int64_t need, union_len = 1;
    if (op == 0) {
        need = (a_len == 1) ? b_len : (b_len == 1 ? a_len : int64_t(a_len) + b_len - 1);
    } else if (op == 1) {
        const int64_t lo = std::min(a_start, b_start), hi = std::max(a_start + int64_t(a_len - 1) * st, b_start + int64_t(b_len - 1) * st);
        need = union_len = (hi - lo) / st + 1;
    } else {
        need = std::max<int64_t>(1, a_len);
    }
    if (need > out_cap) return mcdp_set_error(MCDP_ERR_ARG, "out_probs too small: need " + std::to_string(need) + " bins");
    int prev = -1;
    int32_t rc = use_device(device, &prev);
    if (rc) return rc;
    {
        DevMem mem;
        OpParams p{};
        double *d_a = nullptr, *d_b = nullptr, *d_out = nullptr, *d_scratch = nullptr, *d_flow = nullptr;
        long long* d_start = nullptr;
        int32_t *d_len = nullptr, *d_status = nullptr;
        int32_t status = 0;
        double flow[2] = {0.0, 0.0};
        ACUDA(mem.upload(&d_a, a_probs, size_t(a_len)));
        if (op < 2) ACUDA(mem.upload(&d_b, b_probs, size_t(b_len)));
        ACUDA(mem.alloc(&d_out, size_t(need)));
        ACUDA(mem.alloc(&d_scratch, size_t(4 * union_len)));
        ACUDA(mem.alloc(&d_flow, 2));
        ACUDA(mem.alloc(&d_start, 1));
        ACUDA(mem.alloc(&d_len, 1));
        ACUDA(mem.alloc(&d_status, 1));
        p.a = Pmf{a_start, a_len, d_a, -1.0};
        p.b = Pmf{b_start, b_len, d_b, -1.0};
        p.step = st;
        p.min_value = min_value;
        p.max_value = max_value;
        p.under_rule = underflow_rule;
        p.over_rule = overflow_rule;
        p.out = d_out;
        p.scratch = d_scratch;
        p.scratch_len = union_len;
        p.o_start = d_start;
        p.o_len = d_len;
        p.o_flow = d_flow;
        p.o_status = d_status;
        pmf_op_kernel<<<1, kThreads>>>(p, op);
        ACUDA(cudaGetLastError());
        ACUDA(cudaDeviceSynchronize());
        ACUDA(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
        if (status != kClipOk) {
            rc = mcdp_set_error(MCDP_ERR_INVALID, clip_message(status));
            goto done;
        }
        ACUDA(cudaMemcpy(out_start, d_start, 8, cudaMemcpyDeviceToHost));
        ACUDA(cudaMemcpy(out_len, d_len, 4, cudaMemcpyDeviceToHost));
        ACUDA(cudaMemcpy(out_probs, d_out, size_t(need) * 8, cudaMemcpyDeviceToHost));
        ACUDA(cudaMemcpy(flow, d_flow, 16, cudaMemcpyDeviceToHost));
        if (out_underflow) *out_underflow = flow[0];
        if (out_overflow) *out_overflow = flow[1];
    }
done:
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
