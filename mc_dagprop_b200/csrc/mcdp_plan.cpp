// mcdp_plan.cpp -- host plan compiler.  What Simulator::Simulator does once per propagator
// (reference _core.cpp:193-307), re-thought for the device: instead of a CSR indexed by event id
// plus a separate order vector, the output is ONE evaluation-ordered stream of 32-byte event and
// predecessor records that the kernel reads strictly sequentially, grouped into topological
// LEVELS (longest-path layering) so that several warps can split a level between them.
#include "mcdp_plan.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <unordered_map>

namespace mcdp {

namespace {

bool build_dists(const mcdp_dists_desc& d, HostPlan& out, std::unordered_map<int32_t, int32_t>& type_to_dist,
                 std::string& err) {
    // A later entry for an activity_type replaces the earlier one (dist_map_[t] = ..., _core.cpp:154-158).
    std::unordered_map<int32_t, int32_t> last_entry;
    std::vector<int32_t> type_order;
    for (int32_t t = 0; t < d.n_dists; ++t) {
        if (d.dist_type[t] == -1) {  // _core.cpp:196-198
            err = "Activity type -1 is reserved for no delay";
            return false;
        }
        if (!last_entry.count(d.dist_type[t])) type_order.push_back(d.dist_type[t]);
        last_entry[d.dist_type[t]] = t;
    }
    for (int32_t type : type_order) {
        const int32_t t = last_entry[type];
        DistRec r{};
        r.kind = d.kind[t];
        r.tab_off = -1;
        const double p0 = d.p0[t], p1 = d.p1[t], p2 = d.p2[t];
        switch (r.kind) {
            case MCDP_DIST_CONSTANT:
                r.p[0] = p0;
                break;
            case MCDP_DIST_EXPONENTIAL: {
                // The reference redraws while x > max_scale (_core.cpp:85-87) and never returns for
                // parameters that make acceptance impossible; those are rejected here instead.
                if (!(p0 > 0.0) || !std::isfinite(p0)) {
                    err = "add_exponential: lambda_ must be positive and finite";
                    return false;
                }
                if (!(p1 >= 0.0)) {
                    err = "add_exponential: max_scale must be non-negative";
                    return false;
                }
                r.p[0] = p0;
                r.p[1] = p1;
                r.p[2] = std::isinf(p1) ? 1.0 : -std::expm1(-p1 / p0);  // P(x <= max_scale)
                if (r.p[2] < 0x1p-10) r.flags |= 2;
                r.p[3] = std::isinf(p1) ? 0.0 : std::exp(-p1 / p0);  // 1 - F, for the refined tail (contract v2)
                r.p[7] = 1.0 / p0;  // rate, as ExponentialDist's constructor computes it (_core.cpp:81)
                break;
            }
            case MCDP_DIST_GAMMA: {
                if (!(p0 > 0.0) || !std::isfinite(p0) || !(p1 > 0.0) || !std::isfinite(p1)) {
                    err = "add_gamma: shape and scale must be positive and finite";
                    return false;
                }
                if (!(p2 >= 0.0)) {
                    err = "add_gamma: max_scale must be non-negative";
                    return false;
                }
                const double malpha = p0 < 1.0 ? p0 + 1.0 : p0;  // libstdc++ random.tcc:2339
                const double a1 = malpha - 1.0 / 3.0;
                r.p[0] = p0;
                r.p[1] = p1;
                r.p[2] = p2;
                r.p[3] = a1;
                r.p[4] = 1.0 / std::sqrt(9.0 * a1);
                r.p[5] = 1.0 / p0;
                r.p[6] = a1 * p1;
                r.flags = p0 < 1.0 ? 1 : 0;
                {
                    // 2 * shape in {1..6, 8}: sum of floor(shape) exponentials (+ half a squared normal) is exact
                    // and needs at most the four 32-bit words of one Philox block
                    const double twice = 2.0 * p0;
                    if (twice == std::floor(twice) && twice >= 1.0 && twice <= 8.0 && twice != 7.0) {
                        r.flags |= 8;
                        r.pad0 = int32_t(std::floor(p0));
                        r.pad1 = int32_t(twice) & 1;
                    }
                }
                break;
            }
            case MCDP_DIST_EMP_ABS:
            case MCDP_DIST_EMP_REL: {
                const int64_t n = d.tab_off[t + 1] - d.tab_off[t];
                if (n <= 0) {
                    err = "empirical distribution needs at least one value";
                    return false;
                }
                if (n >= (int64_t(1) << 23)) {
                    err = "empirical distribution table too large";
                    return false;
                }
                const double* vals = d.tab_values + d.tab_off[t];
                const double* w = d.tab_weights + d.tab_off[t];
                r.tab_len = int32_t(n);
                // Guide resolution 2^g: the smallest power of two (>= len, <= 16 x len) for which no bucket
                // [j/G, (j+1)/G) holds more than one cumulative boundary -- then the guide entry plus one
                // compare-and-step IS the lower bound and the device skips the scan loop.  Tables that
                // do not reach that within the cap keep the loop (flags bit2 / meta scan bit).
                int g = 1;  // >= 1 so that the device's `hi >> (32 - g)` is always a valid shift
                bool scan = false;
                if (n >= 2) {
                    double sum0 = 0.0;
                    for (int64_t i = 0; i < n; ++i) sum0 += w[i];
                    int g0 = 0;
                    while ((int64_t(1) << g0) < n) ++g0;
                    const int gmax = std::min(g0 + 4, 20);
                    g = g0;
                    for (;; ++g) {
                        // boundaries are the partial sums; a bucket is clean when consecutive boundaries
                        // never share floor(cp * G)
                        const double G = double(int64_t(1) << g);
                        bool clean = sum0 > 0.0;
                        double acc = 0.0, prev_bucket = -1.0;
                        for (int64_t i = 0; i + 1 < n && clean; ++i) {  // the last boundary is 1.0, outside every bucket
                            acc += w[i] / sum0;
                            const double b = std::floor(acc * G);
                            if (b == prev_bucket) clean = false;
                            prev_bucket = b;
                        }
                        if (clean) break;
                        if (g >= gmax) {
                            scan = true;
                            break;
                        }
                    }
                }
                if (scan) r.flags |= 4;
                r.guide_log2 = g;
                const size_t gd = guide_doubles(uint32_t(g));
                r.tab_off = int32_t(out.tab_pool.size());
                out.tab_pool.resize(out.tab_pool.size() + gd + 2 * size_t(n) + (size_t(n) + 1) / 2, 0.0);
                uint32_t* guide = reinterpret_cast<uint32_t*>(out.tab_pool.data() + r.tab_off);
                double* cp = out.tab_pool.data() + r.tab_off + gd;
                double* v = cp + n;
                uint32_t* thr = reinterpret_cast<uint32_t*>(v + n);
                std::copy(vals, vals + n, v);
                // Integer form of the boundaries for draws from ONE 32-bit word w (tables of <= kQuadTableMaxLen entries):
                // with u = (w + 1/2) 2^-32,  cp[i] < u  <=>  w > cp[i] 2^32 - 1/2  <=>  w > thr[i] = floor(cp[i] 2^32 - 1/2)
                // exactly, so the device compares integers and never forms u.  Boundaries below 2^-33 (thr would be -1)
                // can never be the lower bound of a representable u: thr 0 and the guide skips them.
                auto fill_thresholds = [&] {
                    for (int64_t i = 0; i < n; ++i) {
                        const long double x = static_cast<long double>(cp[i]) * 4294967296.0L - 0.5L;
                        thr[i] = x < 0.0L ? 0u : uint32_t(std::min<long double>(std::floor(x), 4294967295.0L));
                    }
                };
                if (n < 2) {
                    // std::discrete_distribution with < 2 weights always returns index 0
                    // (libstdc++ random.tcc:2660-2664,2703-2704).
                    cp[0] = 1.0;
                    guide[0] = guide[1] = 0;
                    fill_thresholds();
                    break;
                }
                // libstdc++ random.tcc:2655-2678: normalise, partial sums, last = 1 -- same
                // operation order, so the table is bit-identical to the reference's _M_cp.
                double sum = 0.0;
                for (int64_t i = 0; i < n; ++i) {
                    if (!(w[i] >= 0.0)) {
                        err = "empirical distribution weights must be non-negative";
                        return false;
                    }
                    sum += w[i];
                }
                if (!(sum > 0.0) || !std::isfinite(sum)) {
                    err = "empirical distribution weights must have a positive finite sum";
                    return false;
                }
                double acc = 0.0;
                for (int64_t i = 0; i < n; ++i) {
                    const double pr = w[i] / sum;
                    acc = (i == 0) ? pr : acc + pr;
                    cp[i] = acc;
                }
                cp[n - 1] = 1.0;
                // guide[j] = first i with cp[i] >= j / G (tables drawn from 32-bit words: and >= 2^-33, the smallest u such a
                // draw can form)
                const int64_t G = int64_t(1) << g;
                int64_t i = 0;
                if (n <= int64_t(kQuadTableMaxLen))
                    while (i + 1 < n && cp[i] < 0x1p-33) ++i;
                for (int64_t j = 0; j < G; ++j) {
                    const double edge = double(j) / double(G);
                    while (cp[i] < edge) ++i;
                    guide[j] = uint32_t(i);
                }
                fill_thresholds();
                break;
            }
            default:
                err = "unknown distribution kind";
                return false;
        }
        type_to_dist[type] = int32_t(out.dists.size());
        out.dists.push_back(r);
        out.dist_types.push_back(type);
    }
    return true;
}


// Gather-prefetch links of a chunk stream: every entry unit names the source row of the next
// entry unit its warp will process (the next event of the chunk and continuation chunks
// included); the chunk-opening header names the first entry's.
// A dense stream (one warp walks all chunks in order, chunks span levels) may put an event right
// behind one of its own predecessors.  Its first entry then must not be requested while that
// predecessor is still open: the link is cut and the header takes the value over from the
// registers of the event the warp has just closed (`forward`).
void link_chunk_gathers(std::vector<ChunkUnit>& units, bool dense) {
    constexpr size_t U = size_t(kChunkUnits);
    auto kind_of = [](const ChunkUnit& u) { return u.pred.meta >> 29; };
    const size_t n_chunks = units.size() / U;
    uint32_t prev_chunk_last_row = kNoRow;  // row of the event the previous chunk closed last
    for (size_t c = 0; c < n_chunks; ++c) {
        ChunkUnit* ch = &units[c * U];
        const bool cont_in = kind_of(ch[0]) == kKindEnd;  // continuation header: the previous chunk links into this one
        int last_pred = -1, last_pred_head = -1;
        int head = -1;  // header unit of the event being walked
        uint32_t last_row = kNoRow;
        for (size_t u = 0; u < U; ++u) {
            const uint32_t k = kind_of(ch[u]);
            if (u > 0 && k == kKindEnd) break;
            if (k >= kKindEvent) {
                head = int(u);
                last_row = ch[u].head.row;
                continue;
            }
            const uint32_t src = ch[u].pred.src_row;
            ch[u].pred.next_src_row = kNoRow;
            // the unit that requests this entry's source row one step early
            const int link_head = last_pred >= 0 ? last_pred_head : 0;
            bool linked = false;
            if (dense && head != link_head && !(last_pred < 0 && cont_in)) {
                // rows of the events that are still open or not yet started when the early request would be issued
                bool hazard = false;
                uint32_t before = kNoRow;  // row of the event closed right before `head`
                for (int v = link_head; v < head; ++v)
                    if (kind_of(ch[v]) >= kKindEvent) {
                        hazard = hazard || ch[v].head.row == src;
                        before = ch[v].head.row;
                    }
                if (hazard) {
                    if (src == before) ch[head].head.pad = 1u;  // forward from the registers of the event just closed
                    else ch[head].head.first_src_row = src;     // request at the header, after the close
                    linked = true;
                }
            }
            if (!linked) {
                if (last_pred >= 0) {
                    ch[last_pred].pred.next_src_row = src;
                } else if (cont_in) {
                    units[c * U - 1].pred.next_src_row = src;  // last entry unit of the previous (full) chunk
                } else if (dense && head == 0 && src == prev_chunk_last_row) {
                    ch[0].head.pad = 1u;  // the previous chunk's last event is still in this warp's registers
                } else {
                    ch[0].head.first_src_row = src;
                }
            }
            last_pred = int(u);
            last_pred_head = head;
        }
        if (last_row != kNoRow) prev_chunk_last_row = last_row;
    }
}

// Packs the evaluation-ordered records into the chunk stream of mcdp_records.h.  Within a level
// the chunks shrink towards the end (guided self-scheduling: a chunk takes about 1/32 of the units
// still to go, at least 6), so that the warps which split a level reach its barrier within an
// event or two of each other while most of the level is still handed out as full 512-byte chunks.
void build_chunk_stream_impl(const std::vector<EventRec>& events, const std::vector<PredRec>& preds,
                             const std::vector<int32_t>& level_begin_in, int32_t n_levels_in, bool dense,
                             std::vector<ChunkUnit>& units, std::vector<int32_t>& chunk_level_begin) {
    units.clear();
    // dense: one span over all events (no level alignment, full chunks); for launches in which one warp walks
    // the whole stream
    const std::vector<int32_t> one_span = {0, int32_t(events.size())};
    const std::vector<int32_t>& level_begin = dense ? one_span : level_begin_in;
    const int32_t n_levels = dense ? (events.empty() ? 0 : 1) : n_levels_in;
    chunk_level_begin.assign(size_t(n_levels) + 1, 0);
    constexpr uint32_t U = uint32_t(kChunkUnits);
    auto end_unit = [] {
        ChunkUnit u{};
        u.head.meta = kKindEnd << 29;
        u.head.first_src_row = kNoRow;
        return u;
    };
    auto header = [&](const EventRec& ev, bool cont, uint32_t remaining) {
        ChunkUnit u{};
        u.head.row = ev.row;
        u.head.event = ev.event;
        u.head.earliest = ev.earliest;
        u.head.meta = cont ? ((kKindEnd << 29) | 1u) : (kKindEvent << 29);
        u.head.remaining = remaining;
        u.head.first_src_row = kNoRow;
        return u;
    };
    auto pad_chunk = [&] {
        while (units.size() % U) units.push_back(end_unit());
    };
    for (int32_t l = 0; l < n_levels; ++l) {
        const int32_t pb = level_begin[l], pe = level_begin[size_t(l) + 1];
        int64_t units_left = 0;
        for (int32_t p = pb; p < pe; ++p) units_left += 1 + int64_t(events[p].fan_in);
        uint32_t pos = 0;        // units used in the open chunk
        uint32_t budget = U;     // units the open chunk may take
        for (int32_t p = pb; p < pe; ++p) {
            const EventRec& ev = events[p];
            const uint32_t need = 1u + ev.fan_in;
            if (need > U) {  // long event: own run of chunks, 15 entries each
                pad_chunk();
                const uint32_t per = U - 1u;
                const uint32_t n_ch = (ev.fan_in + per - 1u) / per;
                for (uint32_t j = 0; j < n_ch; ++j) {
                    units.push_back(header(ev, j > 0, n_ch - 1u - j));
                    for (uint32_t k = j * per; k < std::min(ev.fan_in, (j + 1u) * per); ++k) {
                        ChunkUnit u{};
                        u.pred = preds[ev.pred_begin + k];
                        units.push_back(u);
                    }
                }
                pad_chunk();
                pos = 0;
            } else {
                if (pos > 0 && pos + need > budget) {
                    pad_chunk();
                    pos = 0;
                }
                if (pos == 0) budget = dense ? U : uint32_t(std::min<int64_t>(U, std::max<int64_t>(6, units_left / 32)));
                units.push_back(header(ev, false, 0));
                for (uint32_t k = 0; k < ev.fan_in; ++k) {
                    ChunkUnit u{};
                    u.pred = preds[ev.pred_begin + k];
                    units.push_back(u);
                }
                pos += need;
                if (pos >= U) pos = 0;  // exactly full
            }
            units_left -= need;
        }
        pad_chunk();
        chunk_level_begin[size_t(l) + 1] = int32_t(units.size() / U);
    }
    link_chunk_gathers(units, dense);
}

}  // namespace

void build_chunk_stream(const std::vector<EventRec>& events, const std::vector<PredRec>& preds,
                        const std::vector<int32_t>& level_begin, int32_t n_levels, bool dense, std::vector<ChunkUnit>& units,
                        std::vector<int32_t>& chunk_level_begin) {
    build_chunk_stream_impl(events, preds, level_begin, n_levels, dense, units, chunk_level_begin);
}

bool compile_plan(const mcdp_graph_desc& g, const mcdp_dists_desc& d, HostPlan& out, std::string& err) {
    out = HostPlan{};
    if (g.n_events < 0 || g.n_act_entries < 0 || g.n_prec_entries < 0 || d.n_dists < 0) {
        err = "negative element count";
        return false;
    }
    std::unordered_map<int32_t, int32_t> type_to_dist;
    if (!build_dists(d, out, type_to_dist, err)) return false;
    if (g.max_delay < 0.0) {  // _core.cpp:199-201
        err = "max_delay must be non-negative";
        return false;
    }
    if (std::isnan(g.max_delay)) {
        err = "max_delay must not be NaN";
        return false;
    }
    const int32_t E = g.n_events;
    out.E = E;
    out.max_delay = g.max_delay;

    // activities: link_count = max idx + 1, gaps are zero-duration links without a distribution
    // (_core.cpp:213-229)
    int32_t max_idx = -1;
    for (int32_t i = 0; i < g.n_act_entries; ++i) {
        if (g.act_idx[i] < 0) {
            err = "Activity.idx must be non-negative";
            return false;
        }
        max_idx = std::max(max_idx, g.act_idx[i]);
    }
    const int32_t A = max_idx + 1;
    out.A = A;
    std::vector<double> base(size_t(A), 0.0);
    std::vector<uint32_t> act_dist(size_t(A), kNoDist);
    for (int32_t i = 0; i < g.n_act_entries; ++i) {
        const int32_t a = g.act_idx[i];
        base[a] = g.act_base[i];
        auto it = type_to_dist.find(g.act_type[i]);
        act_dist[a] = it == type_to_dist.end() ? kNoDist : uint32_t(it->second);
    }

    out.act_base = base;
    out.act_dist = act_dist;

    // precedence: the last entry for a target wins (preds_by_target[tgt] = entry.second, _core.cpp:240);
    // indices are bounds-checked here -- the reference does not (undefined behaviour there).
    std::vector<int32_t> entry_of(size_t(E), -1);
    for (int32_t i = 0; i < g.n_prec_entries; ++i) {
        const int32_t tgt = g.prec_target[i];
        if (tgt < 0 || tgt >= E) {
            err = "precedence_list: target event index out of range";
            return false;
        }
        if (g.prec_off[i + 1] < g.prec_off[i]) {
            err = "precedence_list: offsets must be non-decreasing";
            return false;
        }
        for (int64_t k = g.prec_off[i]; k < g.prec_off[i + 1]; ++k) {
            if (g.pred_src[k] < 0 || g.pred_src[k] >= E) {
                err = "precedence_list: predecessor event index out of range";
                return false;
            }
            // An index >= activity_count() reads past actual_durations_ in the reference (its own
            // LargeScaleTest, test_simulator.py:176-199, relies on that read being 0.0): accepted as
            // a zero-duration link without a duration row.  Negative indices are rejected.
            if (g.pred_act[k] < 0) {
                err = "precedence_list: activity index out of range";
                return false;
            }
        }
        entry_of[tgt] = i;
    }

    // successor CSR over the winning entries + in-degrees
    std::vector<int64_t> succ_off(size_t(E) + 1, 0);
    std::vector<int32_t> indeg(size_t(E), 0);
    int64_t P = 0;
    for (int32_t e = 0; e < E; ++e) {
        const int32_t en = entry_of[e];
        if (en < 0) continue;
        const int64_t n = g.prec_off[en + 1] - g.prec_off[en];
        if (n > std::numeric_limits<int32_t>::max()) {
            err = "precedence_list: fan-in too large";
            return false;
        }
        indeg[e] = int32_t(n);
        P += n;
        for (int64_t k = g.prec_off[en]; k < g.prec_off[en + 1]; ++k) succ_off[size_t(g.pred_src[k]) + 1]++;
    }
    if (P > int64_t(0xFFFFFFF0u)) {
        err = "too many precedence entries";
        return false;
    }
    out.P = P;
    for (int32_t e = 0; e < E; ++e) succ_off[size_t(e) + 1] += succ_off[e];
    std::vector<int32_t> succ(size_t(std::max<int64_t>(P, 1)));
    {
        std::vector<int64_t> pos(succ_off.begin(), succ_off.end() - 1);
        for (int32_t e = 0; e < E; ++e) {
            const int32_t en = entry_of[e];
            if (en < 0) continue;
            for (int64_t k = g.prec_off[en]; k < g.prec_off[en + 1]; ++k) succ[size_t(pos[g.pred_src[k]]++)] = e;
        }
    }

    // Kahn (_core.cpp:248-264) carrying the longest-path level of every event
    std::vector<int32_t> level(size_t(E), 0);
    std::vector<int32_t> queue;
    queue.reserve(size_t(E));
    for (int32_t e = 0; e < E; ++e)
        if (indeg[e] == 0) queue.push_back(e);
    for (size_t qh = 0; qh < queue.size(); ++qh) {
        const int32_t n = queue[qh];
        for (int64_t k = succ_off[n]; k < succ_off[size_t(n) + 1]; ++k) {
            const int32_t dst = succ[size_t(k)];
            level[dst] = std::max(level[dst], level[n] + 1);
            if (--indeg[dst] == 0) queue.push_back(dst);
        }
    }
    if (int32_t(queue.size()) != E) {
        err = "Invalid DAG: cycle detected in precedence list";
        return false;
    }
    int32_t n_levels = 0;
    for (int32_t e = 0; e < E; ++e) n_levels = std::max(n_levels, level[e] + 1);
    out.n_levels = n_levels;

    // evaluation order: by level, ascending event id inside a level (counting sort)
    out.level_begin.assign(size_t(n_levels) + 1, 0);
    for (int32_t e = 0; e < E; ++e) out.level_begin[size_t(level[e]) + 1]++;
    for (int32_t l = 0; l < n_levels; ++l) {
        out.max_level_width = std::max(out.max_level_width, out.level_begin[size_t(l) + 1]);
        out.level_begin[size_t(l) + 1] += out.level_begin[l];
    }
    out.order.resize(size_t(E));
    out.level_of_pos.resize(size_t(E));
    std::vector<int32_t> pos_of_event(static_cast<size_t>(E), 0);
    {
        std::vector<int32_t> cursor(out.level_begin.begin(), out.level_begin.end() - (n_levels ? 1 : 0));
        for (int32_t e = 0; e < E; ++e) {
            const int32_t p = cursor[level[e]]++;
            out.order[p] = e;
            out.level_of_pos[p] = level[e];
            pos_of_event[e] = p;
        }
    }

    auto fill_dist = [&out](PredRec& pr, uint32_t di) {
        pr.dist = di;
        if (di == kNoDist) return;
        const DistRec& r = out.dists[di];
        const bool table = r.kind == MCDP_DIST_EMP_ABS || r.kind == MCDP_DIST_EMP_REL;
        pr.meta = pack_meta(uint32_t(r.kind), table ? uint32_t(r.guide_log2) : 0u, table ? uint32_t(r.tab_len) : 0u,
                            (r.flags & 4) ? 1u : 0u);
        if (table) {
            // byte offsets into the pool: guide block, and the cumulative array right behind it
            pr.tab_off = uint32_t(r.tab_off) * 8u;
            pr.dist = (uint32_t(r.tab_off) + guide_doubles(uint32_t(r.guide_log2))) * 8u;
        } else {
            pr.tab_off = 0u;
        }
    };

    // the stream
    out.events.resize(size_t(E));
    out.preds.resize(size_t(P));
    std::vector<uint32_t> act_refs(size_t(A), 0);
    uint32_t cursor = 0;
    for (int32_t p = 0; p < E; ++p) {
        const int32_t e = out.order[p];
        EventRec& ev = out.events[p];
        ev.row = uint32_t(e);
        ev.event = uint32_t(e);
        ev.pred_begin = cursor;
        ev.earliest = g.earliest[e];
        ev.first_src_row = 0;
        ev.pad = 0;
        const int32_t en = entry_of[e];
        uint32_t fan = 0;
        if (en >= 0) {
            for (int64_t k = g.prec_off[en]; k < g.prec_off[en + 1]; ++k, ++fan) {
                PredRec& pr = out.preds[cursor + fan];
                pr.src_row = uint32_t(g.pred_src[k]);
                pr.next_src_row = k + 1 < g.prec_off[en + 1] ? uint32_t(g.pred_src[k + 1]) : 0u;
                pr.meta = pack_meta(kKindNone, 0, 0);
                pr.tab_off = 0;
                pr.dist = kNoDist;
                if (g.pred_act[k] >= A) {
                    pr.act = kNoAct;
                    pr.base = 0.0;
                    continue;
                }
                pr.act = uint32_t(g.pred_act[k]);
                pr.base = base[pr.act];
                fill_dist(pr, act_dist[pr.act]);
                act_refs[pr.act]++;
            }
        }
        ev.fan_in = fan;
        if (fan) ev.first_src_row = out.preds[cursor].src_row;
        cursor += fan;
        out.max_fan_in = std::max<int32_t>(out.max_fan_in, int32_t(fan));
    }
    for (int32_t a = 0; a < A; ++a) {
        if (act_refs[a] != 0) continue;
        PredRec pr{};
        pr.act = uint32_t(a);
        pr.base = base[a];
        pr.meta = pack_meta(kKindNone, 0, 0);
        pr.dist = kNoDist;
        fill_dist(pr, act_dist[a]);
        out.orphans.push_back(pr);
    }

    build_chunk_stream(out.events, out.preds, out.level_begin, out.n_levels, false, out.units, out.chunk_level_begin);
    out.n_chunks = int32_t(out.units.size() / size_t(kChunkUnits));

    // reduced-mode scratch slots: a slot is released once the LEVEL of the value's last consumer
    // has completed, so warps that split a level never race on a recycled row.
    {
        std::vector<int32_t> last_level(static_cast<size_t>(E), 0);
        for (int32_t e = 0; e < E; ++e) last_level[e] = level[e];
        for (int32_t e = 0; e < E; ++e) {
            const int32_t en = entry_of[e];
            if (en < 0) continue;
            for (int64_t k = g.prec_off[en]; k < g.prec_off[en + 1]; ++k)
                last_level[g.pred_src[k]] = std::max(last_level[g.pred_src[k]], level[e]);
        }
        std::vector<std::vector<int32_t>> release(static_cast<size_t>(n_levels));
        for (int32_t e = 0; e < E; ++e) release[size_t(last_level[e])].push_back(e);
        out.slot_of_event.assign(size_t(E), 0);
        std::vector<uint32_t> free_slots;
        uint32_t next_slot = 0;
        for (int32_t l = 0; l < n_levels; ++l) {
            for (int32_t p = out.level_begin[l]; p < out.level_begin[size_t(l) + 1]; ++p) {
                uint32_t s;
                if (!free_slots.empty()) {
                    s = free_slots.back();
                    free_slots.pop_back();
                } else {
                    s = next_slot++;
                }
                out.slot_of_event[out.order[p]] = s;
            }
            for (int32_t e : release[size_t(l)]) free_slots.push_back(out.slot_of_event[e]);
        }
        out.n_slots = int32_t(next_slot);
    }
    return true;
}

}  // namespace mcdp
