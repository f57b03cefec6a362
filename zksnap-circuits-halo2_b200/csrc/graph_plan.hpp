// graph_plan.hpp — host-side lowering of a halo2 GraphEvaluator (zkb_graph, include/zkb200.h) to the device program run by
// graph.cuh.  Host-only; shared by the product (poly.cu) and the CPU emulator (hostemu.cu).
//
// halo2-axiom's `plonk/evaluation.rs` (un-vendored) numbers its intermediates in SSA fashion — one per calculation — so a
// gate set with a few hundred calculations would need a few hundred 32-byte values per row.  The lowering
//   * folds every scalar source (constants, challenges, beta, gamma, theta, y) into one table,
//   * folds the three column arrays into one table of distinct polynomials (a handle used twice is read through one pointer),
//   * turns rotations into row offsets (rotation * rot_scale).rem_euclid(isize),
//   * and assigns intermediates to SLOTS by liveness (a slot is released after the last read of its intermediate), so the
//     per-row state is the graph's maximum number of simultaneously live values — what the kernel keeps in shared memory;
//     a value read only by the very next instruction never gets a slot: it stays in registers (halo2-base's gate is one such
//     chain from the first product to the Horner fold and needs no shared memory at all).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/zkb200.h"

namespace zkb {

// device instruction: x = op | G_NOSTORE | slot << 8, y / z / w = operand words
// operand word: kind << 29 | index   (kind: 0 scalar table, 1 slot, 2 column query = (polynomial, row offset) pair, 3 previous
//                                     value, 4 the previous instruction's result, still in registers)
constexpr uint32_t G_SCALAR = 0, G_SLOT = 1, G_POLY = 2, G_PREV = 3, G_ACC = 4;
constexpr uint32_t G_KIND_SHIFT = 29, G_MAX_INDEX = 1u << 20, G_MAX_ROT = 1u << 10;
constexpr uint32_t G_NOSTORE = 0x80;   // the result is consumed by the next instruction only (or is the row's result): no slot
constexpr uint32_t G_RESULT_ZERO = 0xFFFFFFFFu;

struct GraphInstrWord {
    uint32_t op_slot, a, b, c;
};

struct GraphPlan {
    std::vector<GraphInstrWord> prog;
    std::vector<uint64_t> scalars;       // 4 u64 each
    std::vector<uint64_t> poly_handles;  // distinct, in first-use order
    std::vector<uint32_t> rot_off;       // per rotation index: (rotation * rot_scale) mod isize
    std::vector<std::pair<uint32_t, uint32_t>> queries;  // distinct (polynomial, row offset) pairs the program reads
    uint32_t nslots = 0;
    uint32_t result_slot = G_RESULT_ZERO;  // 0: the row's result is the last instruction's; G_RESULT_ZERO: empty graph
    bool uses_prev = false;
};

inline uint32_t graph_num_operands(uint32_t op) {
    switch (op) {
        case ZKB_CALC_ADD: case ZKB_CALC_SUB: case ZKB_CALC_MUL: return 2;
        case ZKB_CALC_SQUARE: case ZKB_CALC_DOUBLE: case ZKB_CALC_NEGATE: case ZKB_CALC_STORE: return 1;
        case ZKB_CALC_MUL_ADD: return 3;
        default: return 0;
    }
}

// Copy propagation ahead of the lowering.  Upstream emits one `Store` per column query and starts every Horner with a Store of
// its start value; on the device a Store is a round trip through shared memory.
//   * Store(x) into an intermediate that is written once (and is not the graph's result) makes that intermediate an alias of x
//     when x is a scalar, a column query, the previous value or another write-once intermediate: readers use x directly;
//   * Store(x) into t immediately followed by t = t * f + p (the first Horner step) becomes t = x * f + p.
// The values of all remaining intermediates are unchanged.  Malformed graphs are passed through for graph_lower to report.
inline std::vector<zkb_calculation> graph_simplify(const zkb_graph& g) {
    const size_t nc = g.num_calculations;
    const uint32_t ni = g.num_intermediates;
    std::vector<zkb_calculation> out;
    if (!g.calculations) return out;
    std::vector<uint32_t> writes(ni, 0);
    for (size_t i = 0; i < nc; ++i) {
        if (g.calculations[i].target >= ni || graph_num_operands(g.calculations[i].op) == 0)
            return std::vector<zkb_calculation>(g.calculations, g.calculations + nc);
        ++writes[g.calculations[i].target];
    }
    std::vector<int8_t> has_alias(ni, 0), written(ni, 0);
    std::vector<zkb_value_source> alias(ni);
    auto resolve = [&](zkb_value_source s) {
        if (s.kind == ZKB_SRC_INTERMEDIATE && s.index < ni && has_alias[s.index]) return alias[s.index];
        return s;
    };
    auto reads = [](const zkb_value_source& s, uint32_t t) { return s.kind == ZKB_SRC_INTERMEDIATE && s.index == t; };
    out.reserve(nc);
    for (size_t i = 0; i < nc; ++i) {
        zkb_calculation c = g.calculations[i];
        const uint32_t nop = graph_num_operands(c.op);
        zkb_value_source* s[3] = {&c.a, &c.b, &c.c};
        for (uint32_t k = 0; k < nop; ++k) *s[k] = resolve(*s[k]);
        for (uint32_t k = nop; k < 3; ++k) *s[k] = zkb_value_source{ZKB_SRC_CONSTANT, 0, 0};
        if (c.op == ZKB_CALC_STORE && i + 1 < nc) {
            const bool src_stable = c.a.kind != ZKB_SRC_INTERMEDIATE || (c.a.index < ni && writes[c.a.index] == 1 && written[c.a.index]);
            if (writes[c.target] == 1 && src_stable) {
                has_alias[c.target] = 1;
                written[c.target] = 1;
                alias[c.target] = c.a;
                continue;
            }
            const zkb_calculation& nx = g.calculations[i + 1];
            if (nx.op == ZKB_CALC_MUL_ADD && nx.target == c.target && reads(nx.a, c.target) && !reads(nx.b, c.target) &&
                !reads(nx.c, c.target) && !reads(c.a, c.target)) {
                zkb_calculation m = nx;
                m.a = c.a;
                m.b = resolve(m.b);
                m.c = resolve(m.c);
                out.push_back(m);
                written[c.target] = 1;
                ++i;
                continue;
            }
        }
        out.push_back(c);
        written[c.target] = 1;
    }
    return out;
}

// A WINDOW of rows (one rank's share of the extended domain): every column buffer holds halo_lo rows of the previous
// shard, the shard's own rows, and halo_hi rows of the next one, so that rotated rows are read without wrapping.
struct GraphWindow {
    bool on = false;
    uint64_t halo_lo = 0, halo_hi = 0;
};

// Returns "" on success, otherwise what is wrong with the graph / inputs.  isize: rows of the whole domain (wrapping mode) or of
// the window.
inline std::string graph_lower(const zkb_graph& g, const zkb_graph_inputs& in, uint64_t isize, GraphPlan& plan,
                               const GraphWindow& win = GraphWindow()) {
    plan = GraphPlan();
    if (!win.on && (isize == 0 || (isize & (isize - 1)) || isize > (1ull << 28))) return "the extended domain size must be a power of two <= 2^28";
    if (win.on && (isize == 0 || isize > (1ull << 28) || win.halo_lo > (1u << 20) || win.halo_hi > (1u << 20))) return "bad row window";
    if (g.num_calculations && !g.calculations) return "calculations is NULL";
    if (g.num_constants && !g.constants) return "constants is NULL";
    if (g.num_rotations && !g.rotations) return "rotations is NULL";
    if (g.num_rotations > G_MAX_ROT) return "too many rotations";
    if (g.num_intermediates > G_MAX_INDEX || g.num_calculations > (1u << 24)) return "graph too large";
    if ((in.num_fixed && !in.fixed) || (in.num_advice && !in.advice) || (in.num_instance && !in.instance) ||
        (in.num_challenges && !in.challenges))
        return "a column / challenge array is NULL";

    std::vector<int8_t> rot_in_halo(g.num_rotations, 1);
    for (size_t r = 0; r < g.num_rotations; ++r) {
        int64_t o = (int64_t)g.rotations[r] * (int64_t)in.rot_scale;
        if (win.on) {  // window: index = halo_lo + row + offset, never wraps; a rotation outside the halo is an error if it is read
            if (o < -(int64_t)win.halo_lo || o > (int64_t)win.halo_hi) rot_in_halo[r] = 0;
            o += (int64_t)win.halo_lo;
            if (o < 0) o = 0;
        } else {
            o %= (int64_t)isize;
            if (o < 0) o += (int64_t)isize;
        }
        plan.rot_off.push_back((uint32_t)o);
    }

    std::map<std::pair<uint32_t, uint32_t>, uint32_t> scalar_ix;  // (kind, index) -> table entry
    std::map<uint64_t, uint32_t> poly_ix;
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> query_ix;
    auto scalar = [&](uint32_t kind, uint32_t index, const uint64_t* v) {
        auto key = std::make_pair(kind, index);
        auto it = scalar_ix.find(key);
        if (it != scalar_ix.end()) return it->second;
        uint32_t id = (uint32_t)(plan.scalars.size() / 4);
        plan.scalars.insert(plan.scalars.end(), v, v + 4);
        scalar_ix[key] = id;
        return id;
    };

    const std::vector<zkb_calculation> calcs = graph_simplify(g);
    const size_t nc = calcs.size();
    const uint32_t ni = g.num_intermediates;

    // ---- SSA: every write of an intermediate makes a new VALUE; operands become device words or value ids ------------------------
    struct Node {
        uint32_t op;
        uint32_t word[3];   // device operand word, valid when val[k] < 0
        int32_t val[3];     // value id read, or -1
        uint32_t nop;
        bool live;
    };
    std::vector<Node> nodes(nc);           // node i defines value i
    std::vector<int32_t> cur_val(ni, -1);  // value currently held by each intermediate
    std::string err;
    auto fail = [&](const char* what, size_t i) { if (err.empty()) err = std::string(what) + " (calculation " + std::to_string(i) + ")"; return 0u; };
    auto operand = [&](const zkb_value_source& s, size_t i) -> uint32_t {
        switch (s.kind) {
            case ZKB_SRC_CONSTANT:
                if (s.index >= g.num_constants) return fail("constant out of range", i);
                return G_SCALAR << G_KIND_SHIFT | scalar(s.kind, s.index, g.constants + 4 * (size_t)s.index);
            case ZKB_SRC_CHALLENGE:
                if (s.index >= in.num_challenges) return fail("challenge out of range", i);
                return G_SCALAR << G_KIND_SHIFT | scalar(s.kind, s.index, in.challenges + 4 * (size_t)s.index);
            case ZKB_SRC_BETA: if (!in.beta) return fail("beta is NULL", i); return G_SCALAR << G_KIND_SHIFT | scalar(s.kind, 0, in.beta);
            case ZKB_SRC_GAMMA: if (!in.gamma) return fail("gamma is NULL", i); return G_SCALAR << G_KIND_SHIFT | scalar(s.kind, 0, in.gamma);
            case ZKB_SRC_THETA: if (!in.theta) return fail("theta is NULL", i); return G_SCALAR << G_KIND_SHIFT | scalar(s.kind, 0, in.theta);
            case ZKB_SRC_Y: if (!in.y) return fail("y is NULL", i); return G_SCALAR << G_KIND_SHIFT | scalar(s.kind, 0, in.y);
            case ZKB_SRC_FIXED: case ZKB_SRC_ADVICE: case ZKB_SRC_INSTANCE: {
                const uint64_t* cols = s.kind == ZKB_SRC_FIXED ? in.fixed : s.kind == ZKB_SRC_ADVICE ? in.advice : in.instance;
                const size_t ncols = s.kind == ZKB_SRC_FIXED ? in.num_fixed : s.kind == ZKB_SRC_ADVICE ? in.num_advice : in.num_instance;
                if (s.index >= ncols) return fail("column out of range", i);
                if (s.rotation >= g.num_rotations) return fail("rotation index out of range", i);
                if (!rot_in_halo[s.rotation]) return fail("rotation reaches outside the window's halo", i);
                const uint64_t h = cols[s.index];
                auto it = poly_ix.find(h);
                uint32_t id;
                if (it == poly_ix.end()) {
                    id = (uint32_t)plan.poly_handles.size();
                    plan.poly_handles.push_back(h);
                    poly_ix[h] = id;
                } else id = it->second;
                if (id >= G_MAX_INDEX) return fail("too many polynomials", i);
                const auto key = std::make_pair(id, plan.rot_off[s.rotation]);
                auto qt = query_ix.find(key);
                uint32_t qid;
                if (qt == query_ix.end()) {
                    qid = (uint32_t)plan.queries.size();
                    plan.queries.push_back(key);
                    query_ix[key] = qid;
                } else qid = qt->second;
                if (qid >= G_MAX_INDEX) return fail("too many column queries", i);
                return G_POLY << G_KIND_SHIFT | qid;
            }
            case ZKB_SRC_PREVIOUS: plan.uses_prev = true; return G_PREV << G_KIND_SHIFT;
            default: return fail("unknown value source", i);
        }
    };
    for (size_t i = 0; i < nc; ++i) {
        const zkb_calculation& c = calcs[i];
        Node& n = nodes[i];
        n.op = c.op;
        n.nop = graph_num_operands(c.op);
        n.live = false;
        if (n.nop == 0) return "unknown calculation at " + std::to_string(i);
        if (c.target >= ni) return "calculation " + std::to_string(i) + " writes an intermediate out of range";
        const zkb_value_source* s[3] = {&c.a, &c.b, &c.c};
        for (uint32_t k = 0; k < 3; ++k) { n.word[k] = 0; n.val[k] = -1; }
        for (uint32_t k = 0; k < n.nop; ++k) {
            if (s[k]->kind == ZKB_SRC_INTERMEDIATE) {
                if (s[k]->index >= ni) return "calculation " + std::to_string(i) + " reads an intermediate out of range";
                if (cur_val[s[k]->index] < 0) return "intermediate read before it is written (calculation " + std::to_string(i) + ")";
                n.val[k] = cur_val[s[k]->index];
            } else {
                n.word[k] = operand(*s[k], i);
            }
        }
        if (!err.empty()) return err;
        cur_val[c.target] = (int32_t)i;
    }
    if (nc == 0) return "";
    const int32_t result_val = (int32_t)nc - 1;

    // ---- dead values (never reach the result) are not computed --------------------------------------------------------------------
    nodes[result_val].live = true;
    for (size_t i = nc; i-- > 0;)
        if (nodes[i].live)
            for (uint32_t k = 0; k < nodes[i].nop; ++k)
                if (nodes[i].val[k] >= 0) nodes[nodes[i].val[k]].live = true;
    plan.uses_prev = false;
    for (size_t i = 0; i < nc; ++i)
        if (nodes[i].live)
            for (uint32_t k = 0; k < nodes[i].nop; ++k)
                if (nodes[i].val[k] < 0 && nodes[i].word[k] >> G_KIND_SHIFT == G_PREV) plan.uses_prev = true;

    // ---- order + slots.  emit(order) assigns slots by liveness: a value's slot is released at its last reader, before the reader's
    // own result is placed (operands are in registers before the store).  Two orders are tried: the graph's own, and a list
    // schedule that, among the ready calculations, prefers the one releasing the most values (ties: the graph's order) — upstream
    // computes every gate first and folds them with Horner afterwards, which keeps one value per gate alive; folding each gate as
    // soon as it is complete needs only the deepest gate's values.  The order with fewer slots is used.
    std::vector<uint32_t> uses(nc, 0);
    for (size_t i = 0; i < nc; ++i)
        if (nodes[i].live)
            for (uint32_t k = 0; k < nodes[i].nop; ++k)
                if (nodes[i].val[k] >= 0) ++uses[nodes[i].val[k]];
    auto emit = [&](const std::vector<uint32_t>& order, std::vector<GraphInstrWord>* prog) {
        std::vector<uint32_t> left(uses);
        std::vector<int32_t> slot_of(nc, -1);   // -2: forwarded in registers to the next instruction
        std::vector<uint32_t> free_slots;
        uint32_t nslots = 0;
        for (size_t p = 0; p < order.size(); ++p) {
            const uint32_t i = order[p];
            const Node& n = nodes[i];
            GraphInstrWord w{0, 0, 0, 0};
            uint32_t* dst[3] = {&w.a, &w.b, &w.c};
            for (uint32_t k = 0; k < n.nop; ++k) {
                if (n.val[k] < 0) *dst[k] = n.word[k];
                else if (slot_of[n.val[k]] == -2) *dst[k] = G_ACC << G_KIND_SHIFT;
                else *dst[k] = G_SLOT << G_KIND_SHIFT | (uint32_t)slot_of[n.val[k]];
            }
            for (uint32_t k = 0; k < n.nop; ++k)
                if (n.val[k] >= 0 && --left[n.val[k]] == 0 && slot_of[n.val[k]] >= 0) free_slots.push_back((uint32_t)slot_of[n.val[k]]);
            // all readers of this value sit in the next instruction (or it is the row's result): keep it in registers
            bool forward = (int32_t)i == result_val;
            if (!forward && p + 1 < order.size()) {
                uint32_t in_next = 0;
                const Node& nx = nodes[order[p + 1]];
                for (uint32_t k = 0; k < nx.nop; ++k) in_next += nx.val[k] == (int32_t)i;
                forward = in_next == uses[i];
            }
            if (forward) {
                slot_of[i] = -2;
                w.op_slot = n.op | G_NOSTORE;
            } else {
                if (!free_slots.empty()) { slot_of[i] = (int32_t)free_slots.back(); free_slots.pop_back(); }
                else slot_of[i] = (int32_t)nslots++;
                w.op_slot = n.op | (uint32_t)slot_of[i] << 8;
            }
            if (prog) prog->push_back(w);
        }
        return nslots;
    };
    std::vector<uint32_t> natural;
    for (size_t i = 0; i < nc; ++i)
        if (nodes[i].live) natural.push_back((uint32_t)i);
    std::vector<uint32_t> best = natural;
    uint32_t best_slots = emit(natural, nullptr);
    if (natural.size() <= 8192 && best_slots > 1) {
        std::vector<uint32_t> left(uses), pending(nc, 0), sched;
        std::vector<std::vector<uint32_t>> readers(nc);
        for (uint32_t i : natural) {
            const Node& n = nodes[i];
            for (uint32_t k = 0; k < n.nop; ++k)
                if (n.val[k] >= 0) {
                    bool dup = false;
                    for (uint32_t q = 0; q < k; ++q) dup |= n.val[q] == n.val[k];
                    if (!dup) { ++pending[i]; readers[n.val[k]].push_back(i); }
                }
        }
        std::vector<uint32_t> ready;
        for (uint32_t i : natural)
            if (pending[i] == 0) ready.push_back(i);
        while (!ready.empty()) {
            size_t pick = 0;
            int pick_score = -100;
            for (size_t r = 0; r < ready.size(); ++r) {
                const Node& n = nodes[ready[r]];
                int kills = 0;
                for (uint32_t k = 0; k < n.nop; ++k)
                    if (n.val[k] >= 0) {
                        uint32_t same = 0;
                        for (uint32_t q = 0; q < n.nop; ++q) same += n.val[q] == n.val[k];
                        bool first = true;
                        for (uint32_t q = 0; q < k; ++q) first &= n.val[q] != n.val[k];
                        if (first && left[n.val[k]] == same) ++kills;
                    }
                if (kills > pick_score || (kills == pick_score && ready[r] < ready[pick])) { pick = r; pick_score = kills; }
            }
            const uint32_t i = ready[pick];
            ready.erase(ready.begin() + (long)pick);
            sched.push_back(i);
            for (uint32_t k = 0; k < nodes[i].nop; ++k)
                if (nodes[i].val[k] >= 0) --left[nodes[i].val[k]];
            for (uint32_t r : readers[i])
                if (--pending[r] == 0) ready.push_back(r);
        }
        if (sched.size() == natural.size()) {
            const uint32_t s2 = emit(sched, nullptr);
            if (s2 < best_slots) { best = sched; best_slots = s2; }
        }
    }
    plan.prog.reserve(best.size());
    plan.nslots = emit(best, &plan.prog);
    plan.result_slot = 0;
    if (plan.nslots >= (1u << 20)) return "too many live intermediates";
    return "";
}

}  // namespace zkb
