// mcdp_plan.hpp -- host-side plan compiler (replaces Simulator::Simulator, reference
// src/mc_dagprop/monte_carlo/_core.cpp:193-307).  Pure C++, no CUDA.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mcdp_b200.h"
#include "mcdp_records.h"

namespace mcdp {

struct HostPlan {
    int32_t E = 0, A = 0;
    int64_t P = 0;
    double max_delay = 0.0;
    int32_t n_levels = 0;
    int32_t max_fan_in = 0;
    int32_t max_level_width = 0;

    // evaluation-ordered stream (rows = event ids)
    std::vector<EventRec> events;
    std::vector<PredRec> preds;
    std::vector<int32_t> level_begin;  // [n_levels + 1] positions into events
    std::vector<PredRec> orphans;      // activities no precedence entry references (src fields unused)

    // chunk stream of the full / injected sweep (rows = event ids), see mcdp_records.h
    std::vector<ChunkUnit> units;           // n_chunks * kChunkUnits
    std::vector<int32_t> chunk_level_begin;  // [n_levels + 1] positions into chunks
    int32_t n_chunks = 0;

    // distributions
    std::vector<DistRec> dists;
    std::vector<int32_t> dist_types;   // activity_type of dists[i]
    std::vector<double> tab_pool;      // [guide u32 x 2^g][cp][values] blocks

    // per-activity view in index order (reference-stream compatibility sampler)
    std::vector<double> act_base;
    std::vector<uint32_t> act_dist;

    // introspection
    std::vector<int32_t> order;        // event id at each position
    std::vector<int32_t> level_of_pos; // level of each position

    // reduced mode: realized scratch rows recycled once every consumer of a value has run
    std::vector<uint32_t> slot_of_event;
    int32_t n_slots = 0;
};

// Packs evaluation-ordered records into the chunk stream (mcdp_records.h); also used for the reduced mode's
// slot-row variant of the records.  `dense`: no level alignment (for launches in which one warp walks the
// whole stream in order); chunk_level_begin is then {0, n_chunks}.
void build_chunk_stream(const std::vector<EventRec>& events, const std::vector<PredRec>& preds,
                        const std::vector<int32_t>& level_begin, int32_t n_levels, bool dense, std::vector<ChunkUnit>& units,
                        std::vector<int32_t>& chunk_level_begin);

// Returns false and fills `err` (the reference's std::runtime_error texts where it has one).
bool compile_plan(const mcdp_graph_desc& g, const mcdp_dists_desc& d, HostPlan& out, std::string& err);

}  // namespace mcdp
