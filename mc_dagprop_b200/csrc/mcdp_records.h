// mcdp_records.h -- POD records of the device-resident plan (shared by the host plan
// compiler and the sm_100a kernels).  See DESIGN.md section 3 for the layout rationale.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace mcdp {

constexpr uint32_t kNoDist = 0xFFFFFFFFu;
constexpr uint32_t kNoAct = 0xFFFFFFFFu;  // precedence entry whose activity index has no duration row

// One event in evaluation order.  `row` is the event's row in the realized/cause arrays
// (== event id in full/injected mode; a recycled scratch slot in reduced mode).  `first_src_row`
// repeats the source row of the first precedence entry so that its realized row can be requested
// as soon as the event record arrives, without waiting for the entry record.
struct alignas(16) EventRec {
    uint32_t row;
    uint32_t event;
    uint32_t pred_begin;
    uint32_t fan_in;
    double earliest;
    uint32_t first_src_row;
    uint32_t pad;
};
static_assert(sizeof(EventRec) == 32, "EventRec must be 32 bytes");

// One precedence entry (src event --activity--> this event), in the caller's order.  Everything the
// sampler needs to dispatch is in the record itself (`meta`, `tab_off`), so the kind switch and the
// table lookup do not wait on a dependent DistRec load.
//   meta = kind << 29 | guide_log2 << 24 | scan << 23 | table_len      (kind 5 = no distribution;
//          scan = 1 when a guide bucket may hold more than one cumulative boundary)
//   table block in the pool: [guide: 2^g u32][cp: len f64][values: len f64][thr: len u32 (integer boundaries for
//   32-bit draws, padded to 8 bytes)]; for table kinds
//   `tab_off` is the BYTE offset of the guide and `dist` the BYTE offset of cp (values follow cp)
struct alignas(16) PredRec {
    uint32_t src_row;   // row of the source event's realized time; in full/injected mode also the value of cause_event
    uint32_t act;
    double base;        // Activity.minimal_duration
    uint32_t meta;
    uint32_t tab_off;
    uint32_t next_src_row; // src_row of the FOLLOWING entry of the same event (gather prefetch), else 0
    uint32_t dist;      // index into DistRec[] (constant / exponential / gamma), byte offset of cp (tables), kNoDist = none
};
static_assert(sizeof(PredRec) == 32, "PredRec must be 32 bytes");

// Generator contract v2, width rules (restated by the oracle): which draws take one 32-bit word of a QUAD block
// instead of 64 bits of a PAIR block
constexpr uint32_t kQuadTableMaxLen = 4096u;  // empirical tables of at most this many entries (then guide_log2 <= 16)
// Exponentials always start from one 32-bit word w: u = (w + 1/2) 2^-32.  In the top 2^-20 of the unit interval
// (w >= kExpTailWord), where the inverse CDF is steep, a second word w' refines the same draw to 64 bits:
// 1 - u = ((2^32 - w) - (w' + 1/2) 2^-32) 2^-32 -- the tail keeps the resolution a 64-bit draw would have.
constexpr uint32_t kExpTailWord = 0xFFFFF000u;

constexpr uint32_t kKindNone = 5u;   // precedence entry without a distribution (duration = base)
constexpr uint32_t kKindEvent = 6u;  // chunk stream: event header unit
constexpr uint32_t kKindEnd = 7u;    // chunk stream: end of chunk (meta bit0 set: continuation header)
constexpr uint32_t kNoRow = 0xFFFFFFFFu;
__host__ __device__ inline uint32_t pack_meta(uint32_t kind, uint32_t guide_log2, uint32_t len, uint32_t scan = 0) {
    return (kind << 29) | (guide_log2 << 24) | (scan << 23) | (len & 0x7FFFFFu);
}
__host__ __device__ inline uint32_t guide_doubles(uint32_t guide_log2) {
    const uint32_t g = 1u << guide_log2;
    return g >= 2u ? g / 2u : 1u;
}

// ---- chunk stream (full / injected sweep) -------------------------------------------------------
// The evaluation-ordered stream the sweep kernel consumes: 32-byte units grouped into 512-byte
// chunks (kChunkUnits units).  A chunk is the unit of work a warp pulls (one bulk copy into its
// shared-memory ring, cp.async.bulk + mbarrier) and never straddles a topological level:
//     [EVENT header][its precedence entries ...][EVENT header][...] ... [END]
// An event whose 1 + fan_in units do not fit starts a fresh chunk and continues through
// `remaining` continuation chunks ([CONT header][entries ...]) that the same warp processes; a
// warp that is handed a continuation chunk by the level cursor skips it.
// Precedence-entry units are PredRec as is; `next_src_row` names the source row of the next entry
// unit the same warp will process (next event of the chunk included), kNoRow when there is none.
constexpr int kChunkUnits = 16;
constexpr int kChunkBytes = kChunkUnits * 32;
struct alignas(16) HeaderUnit {
    uint32_t row;            // realized / cause row of the event
    uint32_t event;          // event id
    double earliest;
    uint32_t meta;           // kKindEvent << 29, or kKindEnd << 29 | 1 for a continuation header
    uint32_t remaining;      // continuation chunks that still follow this chunk's last event
    uint32_t first_src_row;  // chunk-opening header: source row of the chunk's first entry (else kNoRow)
    uint32_t pad;            // 1 = forward: the first entry's source is the event this warp closed just before
};
static_assert(sizeof(HeaderUnit) == 32, "HeaderUnit must be 32 bytes");
union ChunkUnit {
    PredRec pred;
    HeaderUnit head;
};
static_assert(sizeof(ChunkUnit) == 32, "ChunkUnit must be 32 bytes");

// Distribution parameters, one per activity_type.
//   CONSTANT     p0 = factor
//   EXPONENTIAL  p0 = lambda (mean), p1 = max_scale, p2 = F = 1 - exp(-max_scale/lambda);
//                p3 = 1 - F = exp(-max_scale/lambda); flags bit1 = F < 2^-10 (series instead of log, no tail refinement)
//   GAMMA        p0 = shape, p1 = scale, p2 = max_scale, p3 = d = shape' - 1/3,
//                p4 = c = 1/sqrt(9 d), p5 = 1/shape, p6 = d * scale; flags bit0 = shape < 1 (boost)
//   EMP_ABS/REL  tab_len entries in the pool block at tab_off (see PredRec)
struct alignas(16) DistRec {
    int32_t kind;
    int32_t tab_len;
    int32_t tab_off;
    int32_t guide_log2;
    int32_t flags;      // bit0 gamma shape < 1, bit1 exponential series, bit2 table needs the scan loop,
                        // bit3 gamma with 2*shape in {1..6, 8} (exact transformation: pad0 = floor(shape), pad1 = half term)
    int32_t pad0, pad1, pad2;
    double p[8];
};
static_assert(sizeof(DistRec) == 96, "DistRec must be 96 bytes");

}  // namespace mcdp
