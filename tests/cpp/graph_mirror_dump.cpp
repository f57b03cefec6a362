// Builds the same graphs as tests/test_cpp_mirror.py::test_cpp_graph_builders_match_the_python_mirror with the C++ host mirror
// (include/zkb200_halo2.hpp) and prints them: "graph <name>", "const <4 hex limbs>", "rot <i>", "calc <11 u32>".  CPU only.
#include <cstdio>

#include "zkb200_halo2.hpp"

using namespace halo2;
using E = Expression;

static void dump(const char* name, const GraphEvaluator& g) {
    std::printf("graph %s %u\n", name, g.num_intermediates);
    for (const auto& c : g.constants) std::printf("const %016llx %016llx %016llx %016llx\n", (unsigned long long)c[0], (unsigned long long)c[1], (unsigned long long)c[2], (unsigned long long)c[3]);
    for (int32_t r : g.rotations) std::printf("rot %d\n", r);
    for (const auto& c : g.calculations)
        std::printf("calc %u %u %u %u %u %u %u %u %u %u %u\n", c.op, c.target, c.a.kind, c.a.index, c.a.rotation, c.b.kind, c.b.index, c.b.rotation, c.c.kind,
                    c.c.index, c.c.rotation);
}

static E::Ptr gate(uint32_t adv, uint32_t sel) {  // halo2-base: q (a + b c - d)
    auto q = [&](int r) { return E::query(E::Advice, adv, r); };
    return E::product(E::query(E::Fixed, sel, 0), E::sum(E::sum(q(0), E::product(q(1), q(2))), E::neg(q(3))));
}

int main() {
    const Fr five = fr::from_u64(5), seven = fr::from_u64(7);
    // custom gates: the gate twice on different columns, plus expressions that hit every special case of add_expression
    std::vector<E::Ptr> gates = {
        gate(0, 0), gate(1, 1),
        E::sum(E::scaled(E::query(E::Instance, 0, -1), seven), E::neg(E::constant_(five))),
        E::product(E::constant_(fr::from_u64(2)), E::product(E::query(E::Advice, 0, 0), E::query(E::Advice, 0, 0))),
        E::sum(E::constant_(Fr{0, 0, 0, 0}), E::product(E::constant_(fr::ONE), E::challenge(1))),
        E::sum(E::neg(E::query(E::Fixed, 1, 2)), E::query(E::Advice, 2, 0)),
        E::scaled(E::query(E::Advice, 1, 1), fr::ONE)};
    dump("custom_gates", custom_gates_graph(gates));
    // permutation: 5 advice columns in sets of 2; fixed 0 l_0, 1 l_last, 2 l_active, 3 X coset, 4.. sigmas; advice 5.. the z cosets
    std::vector<ColumnRef> cols, sigmas, zs;
    for (uint32_t i = 0; i < 5; ++i) { cols.push_back({ZKB_SRC_ADVICE, i}); sigmas.push_back({ZKB_SRC_FIXED, 4 + i}); }
    for (uint32_t i = 0; i < 3; ++i) zs.push_back({ZKB_SRC_ADVICE, 5 + i});
    dump("permutation", permutation_graph(cols, 2, -4, {ZKB_SRC_FIXED, 0}, {ZKB_SRC_FIXED, 1}, {ZKB_SRC_FIXED, 2}, {ZKB_SRC_FIXED, 3}, sigmas, zs));
    // lookup: inputs (a0, a1 * a0(wX)), table (f3, 3 f3(w^-1 X)); advice 2 z, 3 a', 4 s'
    std::vector<E::Ptr> in = {E::query(E::Advice, 0, 0), E::product(E::query(E::Advice, 1, 0), E::query(E::Advice, 0, 1))};
    std::vector<E::Ptr> tab = {E::query(E::Fixed, 3, 0), E::scaled(E::query(E::Fixed, 3, -1), fr::from_u64(3))};
    dump("lookup", lookup_graph(in, tab, {ZKB_SRC_FIXED, 0}, {ZKB_SRC_FIXED, 1}, {ZKB_SRC_FIXED, 2}, {ZKB_SRC_ADVICE, 2}, {ZKB_SRC_ADVICE, 3}, {ZKB_SRC_ADVICE, 4}));
    // QuotientEvaluator: construction compiles the gate graph (evaluate_h itself needs a GPU and is exercised there)
    PermutationArgument perm{cols, 2, -4};
    QuotientEvaluator q(gates, &perm, {LookupArgument{in, tab}});
    (void)q;
    return 0;
}
