// graph.cuh — quotient evaluation: halo2's GraphEvaluator run for every row of the extended domain.
//
// Restates (on the device) the row loop of halo2-axiom `plonk/evaluation.rs::evaluate_h` over a compiled `GraphEvaluator`
// (source un-vendored; the field is exact, so the values are fixed by the calculations).  One thread owns one row and runs the
// lowered program of graph_plan.hpp: operands come from the scalar table, from a polynomial at a rotated row, from the row's
// previous value or from a slot; results go to a slot.  Slots live in shared memory as two 16-byte planes per slot with the
// thread index innermost, so every access is a conflict-free 128-bit LDS/STS; the program is staged in shared memory once per
// CTA and read uniformly.  Slot values are kept lazy (< 2r; fp_*_lazy), the stored result is canonical.  HBM traffic is one
// 32-byte read per polynomial — rotated rows of one column fall in lines the neighbouring threads fetch anyway — plus one
// 32-byte write (ncu: DRAM bytes = algorithmic bytes), but already halo2-base's gate q (a + b c - d), three products per row for
// 128 bytes, keeps the multiply pipe 84 % busy at a third of the HBM rate: like the NTT, this kernel is integer-bound.
#pragma once
#include "field.cuh"
#include "graph_plan.hpp"

namespace zkb {

struct GraphQuery {
    const uint4* poly;   // isize elements
    uint64_t off;        // (rotation * rot_scale) mod isize, or halo_lo + rotation * rot_scale in a row window
};

struct GraphArgs {
    const uint4* prog;          // ninstr device instructions (global copy; the kernel stages them in shared memory)
    uint32_t ninstr;
    uint32_t result_slot;       // G_RESULT_ZERO: empty graph, the row's value is zero; otherwise the last instruction's result
    const uint4* scalars;       // scalar table, 32 B each
    const GraphQuery* queries;  // distinct (polynomial, row offset) pairs
    uint4* values;              // nrows elements: previous value in, result out
    uint64_t nrows;             // rows evaluated
    uint64_t mask;              // whole domain: nrows - 1 (rotated rows wrap); row window: all ones (offsets include the halo)
};

ZKB_HD Fr graph_operand(const GraphArgs& g, uint32_t w, uint64_t idx, const uint4* slots, uint32_t stride, uint32_t lane, const Fr& acc) {
    const uint32_t kind = w >> G_KIND_SHIFT, ix = w & (G_MAX_INDEX - 1);
    switch (kind) {
        case G_SCALAR: return fr_load2(g.scalars, ix);
        case G_SLOT: return fr_from_u4(slots[(2 * ix) * stride + lane], slots[(2 * ix + 1) * stride + lane]);
        case G_POLY: {
            const GraphQuery q = g.queries[ix];
            return fr_load2(q.poly, (idx + q.off) & g.mask);
        }
        case G_PREV: return fr_load2(g.values, idx);
        default: return acc;
    }
}

// prog: where the thread reads instructions from (shared memory on the device); slots: the CTA's slot planes,
// stride = threads per CTA, lane = thread index in the CTA.
ZKB_HD void graph_eval_thread(const GraphArgs& g, const uint4* prog, uint64_t idx, uint4* slots, uint32_t stride, uint32_t lane) {
    if (idx >= g.nrows) return;
    Fr acc = Fr::zero();   // the previous instruction's result
    for (uint32_t i = 0; i < g.ninstr; ++i) {
        const uint4 ins = prog[i];
        const uint32_t op = ins.x & 0x7F, dst = ins.x >> 8;
        const Fr a = graph_operand(g, ins.y, idx, slots, stride, lane, acc);
        Fr r;
        if (op == ZKB_CALC_MUL || op == ZKB_CALC_SQUARE || op == ZKB_CALC_MUL_ADD) {
            const Fr b = op == ZKB_CALC_SQUARE ? a : graph_operand(g, ins.z, idx, slots, stride, lane, acc);
            r = fp_mul_lazy(a, b);
            if (op == ZKB_CALC_MUL_ADD) r = fp_add_lazy(r, graph_operand(g, ins.w, idx, slots, stride, lane, acc));
        } else if (op == ZKB_CALC_ADD || op == ZKB_CALC_DOUBLE) {
            r = fp_add_lazy(a, op == ZKB_CALC_DOUBLE ? a : graph_operand(g, ins.z, idx, slots, stride, lane, acc));
        } else if (op == ZKB_CALC_SUB) {
            r = fp_sub_lazy(a, graph_operand(g, ins.z, idx, slots, stride, lane, acc));
        } else if (op == ZKB_CALC_NEGATE) {
            r = fp_sub_lazy(Fr::zero(), a);
        } else {
            r = a;  // ZKB_CALC_STORE
        }
        if (!(ins.x & G_NOSTORE)) {
            slots[(2 * dst) * stride + lane] = make_uint4(r.l[0], r.l[1], r.l[2], r.l[3]);
            slots[(2 * dst + 1) * stride + lane] = make_uint4(r.l[4], r.l[5], r.l[6], r.l[7]);
        }
        acc = r;
    }
    // the row's result is the last instruction's (every other value is one of its ancestors); zero for an empty graph
    fr_store2(g.values, idx, g.result_slot == G_RESULT_ZERO ? Fr::zero() : fp_canon(acc));
}

// threads per CTA for a graph with `nslots` slots: the widest CTA whose slots fit in shared memory next to the program
inline uint32_t graph_cta_threads(uint32_t nslots, uint32_t ninstr, size_t smem_limit) {
    for (uint32_t t = 128; t >= 32; t >>= 1)
        if ((size_t)nslots * 32 * t + (size_t)ninstr * 16 <= smem_limit) return t;
    return 0;
}

}  // namespace zkb
