// mcdp_sweep.cuh -- what the sweep kernels (mcdp_chunk_sweep.cuh: two samples per lane, mcdp_quad_sweep.cuh: four)
// share: the parameter block, the modes, warp-level helpers and the transpose of the host-facing calls.
//
// The kernels replace Simulator::run (reference _core.cpp:312-353) for a block of seeds:
//   * one warp owns 64 / 128 adjacent samples of the event-major / sample-minor arrays, so every predecessor read and
//     every realized / cause / duration write is a vector access to one contiguous row segment per warp;
//   * `warps_per_group` warps share the same samples and split every topological level between them, meeting at a
//     barrier per level: this is how large DAGs, whose per-sample output footprint limits the resident sample
//     count, still fill the SMs;
//   * the delay of each precedence entry is drawn (Philox, mcdp_sampling.cuh) at the point of use and streamed out,
//     never re-read.
#pragma once
#include "mcdp_sampling.cuh"

namespace mcdp {

// CTA shape: at most 16 warps (groups_per_cta x warps_per_group), two such CTAs per SM => 64
// registers per thread and 32 resident warps per SM.
#ifndef MCDP_MAX_THREADS
#define MCDP_MAX_THREADS 512
#endif
#ifndef MCDP_MIN_BLOCKS
#define MCDP_MIN_BLOCKS 2
#endif

// kModeAttr: reduced statistics + delay-cause attribution (per-activity counts of being the binding predecessor)
enum SweepMode { kModeFull = 0, kModeInjected = 1, kModeReduced = 2, kModeAttr = 3 };

struct SweepParams {
    const ChunkUnit* chunks;             // chunk stream the sweep kernels walk (mcdp_records.h)
    const int32_t* chunk_level_begin;    // [n_levels + 1] positions into chunks
    const PredRec* orphans;
    const DistRec* dists;
    const double* tab_pool;
    const double* log_tab;  // {rc, lc} pairs of mcdp_math.cuh (kLogTabEntries)
    const int32_t* seeds;  // nullptr => seed0 + sample index
    double* realized;      // [rows][ld]  (output in full/injected mode, scratch in reduced mode)
    double* durations;     // [A][ld]     full mode: written
    const double* inj;     // [A][ld]     injected mode: read
    int32_t* cause;        // [E][ld]
    double* sum;           // reduced mode accumulators
    double* sumsq;
    unsigned long long* late;
    uint32_t* hist;
    unsigned long long* cause_act;   // kModeAttr: [A] samples in which an entry with this activity decided its target
    unsigned long long* cause_none;  // kModeAttr: [E] samples with cause_event == -1
    double thresholds[MCDP_MAX_THRESHOLDS];
    double hist_lo, hist_scale;
    double max_delay;
    int64_t n, ld;
    uint32_t ldb8, ldb4;                // ld * 8, ld * 4 (row strides in bytes, < 2^32)
    uint32_t smem_tab_off, smem_ring_off;  // dynamic shared memory: byte offsets of the table pool and of the chunk rings
    uint32_t smem_stat_off;                // quad kernel, reduced modes: byte offset of the statistics staging areas
    int32_t n_levels, n_orphans, n_dists, tab_pool_len;
    int32_t n_thresholds, n_bins, E, n_chunks;
    int32_t seed0;
    PhiloxKeys keys;  // round keys of Philox key word 0 (stream_key + r * 0x9E3779B9)
    int32_t warps_per_group;
    int32_t batches_per_group;  // always 1 (kept for the layout of launch_shape's report)
    int32_t cluster_size;       // quad kernel: CTAs per thread-block cluster that share one sample group (1 = no cluster)
};

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Warp totals of s and q with ONE butterfly: after the first exchange the lower half-warp carries s, the upper
// half-warp q.  Lanes 0-15 return the total of s, lanes 16-31 the total of q.
__device__ __forceinline__ double warp_sum2(double s, double q, int lane) {
    const bool upper = lane & 16;
    const double give = upper ? s : q, keep = upper ? q : s;
    double v = keep + __shfl_xor_sync(0xFFFFFFFFu, give, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Per-event statistics that are sums over the warp's samples (reduced modes).  `s` / `q`: this lane's sum of delays
// and of squared delays; `late_packed`: this lane's exceedance counts, 8 bits per threshold (a warp holds at most 128
// samples, so the fields of the warp total cannot carry into each other).  One butterfly, one REDUX, and the atomics
// of the thresholds issued by different lanes of one instruction.
template <typename P>
__device__ __forceinline__ void flush_sums(const P& p, uint32_t ev, int lane, double s, double q, uint32_t late_packed) {
    if (p.sum || p.sumsq) {
        const double t = warp_sum2(s, q, lane);
        if (lane == 0 && p.sum) atomicAdd(p.sum + ev, t);
        if (lane == 16 && p.sumsq) atomicAdd(p.sumsq + ev, t);
    }
    if (p.late) {
        const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, late_packed);
        if (lane < p.n_thresholds) {
            const uint32_t cnt = (tot >> (8 * lane)) & 0xFFu;
            if (cnt) atomicAdd(p.late + size_t(lane) * p.E + ev, (unsigned long long)cnt);
        }
    }
}

// std::min(a, b) of the reference build: (b < a) ? b : a  -- NOT fmin (NaN / signed-zero differ).
__device__ __forceinline__ double ref_min(double a, double b) { return (b < a) ? b : a; }

// One relaxation of the recurrence (_core.cpp:341-346) without materialising the clamp.  The reference keeps
//     lat <- t  if  t >= lat,   t = std::min(a, ub),   a = realized[src] + duration,
// and closes the event with realized = std::min(lat, ub).  Here `run` holds the UNCLAMPED arrival of the entry that
// last won, with the invariant lat == std::min(run, ub) (bit for bit):
//   * a >= ub:  t is ub (or a itself when they compare equal) and t >= lat always holds because lat <= ub: the entry
//               wins, and std::min(a, ub) reproduces t, signed zeros included;
//   * a <  ub:  t is a; lat is run when run <= ub (the test is a >= run) and ub otherwise (a >= ub fails, and so does
//               a >= run since a < ub < run);
//   * a NaN:    both compares fail, as t >= lat does in the reference (t is NaN); a NaN `run` or ub can only come from
//               a NaN `earliest`, and then nothing ever wins in either formulation.
// Two compares (the second folds the first in: DSETP.GE.OR) and predicated moves instead of compare + two selects for
// the clamp and compare + three selects for the update.  Bit-exact parity incl. NaN / inf / signed zeros is pinned by
// tests/test_gpu_parity.py::test_injected_special_values and the pair kernel, which keeps the literal form.
__device__ __forceinline__ void relax(double a, double ub, double& run, int& cause, int src) {
    const bool take = (a >= run) | (a >= ub);
    run = take ? a : run;
    cause = take ? src : cause;
}

// out[c][r] = in[r][c]  (in: rows x cols with row stride in_ld; out: cols x rows, stride out_ld).
// 1-D grid of 32x32 tiles (either extent can exceed the 65535 limit of grid.y).
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, int64_t in_ld, int64_t rows,
                                                        int64_t cols, T* __restrict__ out, int64_t out_ld,
                                                        int64_t tiles_c) {
    __shared__ T tile[32][33];
    const int64_t tile_r = int64_t(blockIdx.x) / tiles_c, tile_c = int64_t(blockIdx.x) - tile_r * tiles_c;
    const int64_t c0 = tile_c * 32, r0 = tile_r * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t r = r0 + ty + j, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + j][tx] = in[r * in_ld + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t c = c0 + ty + j, r = r0 + tx;
        if (r < rows && c < cols) out[c * out_ld + r] = tile[tx][ty + j];
    }
}

}  // namespace mcdp
