// mcdp_records.h -- POD records of the device-resident plan (shared by the host plan
// compiler and the sm_100a kernels).  See DESIGN.md section 3 for the layout rationale.
#pragma once
#include <cstdint>

namespace mcdp {

constexpr uint32_t kNoDist = 0xFFFFFFFFu;
constexpr uint32_t kNoAct = 0xFFFFFFFFu;  // precedence entry whose activity index has no duration row

// One event in evaluation order.  `row` is the event's row in the realized/cause arrays
// (== event id in full/injected mode; a recycled scratch slot in reduced mode).
// `ub` = earliest + max_delay is rounded once on the host: the same IEEE add the reference
// performs per sample at _core.cpp:334.
struct alignas(16) EventRec {
    uint32_t row;
    uint32_t event;
    uint32_t pred_begin;
    uint32_t fan_in;
    double earliest;
    double ub;
};
static_assert(sizeof(EventRec) == 32, "EventRec must be 32 bytes");

// One precedence entry (src event --activity--> this event), in the caller's order.
struct alignas(16) PredRec {
    uint32_t src_row;
    uint32_t act;
    double base;        // Activity.minimal_duration
    uint32_t dist;      // index into DistRec[], kNoDist = duration is `base`
    uint32_t src_event; // value written to cause_event
    uint32_t pad0, pad1;
};
static_assert(sizeof(PredRec) == 32, "PredRec must be 32 bytes");

// An activity that no precedence entry references: sampled and written, never propagated.
struct alignas(16) OrphanRec {
    uint32_t act;
    uint32_t dist;
    double base;
};
static_assert(sizeof(OrphanRec) == 16, "OrphanRec must be 16 bytes");

// Distribution parameters, one per activity_type.
//   CONSTANT     p0 = factor
//   EXPONENTIAL  p0 = lambda (mean), p1 = max_scale, p2 = F = 1 - exp(-max_scale/lambda);
//                flags bit1 = F < 2^-10 (series instead of log)
//   GAMMA        p0 = shape, p1 = scale, p2 = max_scale, p3 = d = shape' - 1/3,
//                p4 = c = 1/sqrt(9 d), p5 = 1/shape, p6 = d * scale; flags bit0 = shape < 1 (boost)
//   EMP_ABS/REL  tab_len entries: cumulative at tab_pool[tab_off .. +len), values at
//                tab_pool[tab_off+len .. +2 len); guide table guide_pool[guide_off .. + 2^guide_log2)
struct alignas(16) DistRec {
    int32_t kind;
    int32_t tab_len;
    int32_t tab_off;
    int32_t guide_off;
    int32_t guide_log2;
    int32_t flags;
    int32_t pad0, pad1;
    double p[8];
};
static_assert(sizeof(DistRec) == 96, "DistRec must be 96 bytes");

}  // namespace mcdp
