// Compiles against include/zkb200_halo2.hpp and links libzkb200.so.
//   no argument : CPU-only checks (domain geometry, omega constant, loud failure without a device)
//   "gpu"       : SURVEY.md §8c KAT best_fft([1..8], omega(3), 3) limbs + coeff_to_extended/extended_to_coeff round trip
#include <cstdio>
#include <cstring>

#include "zkb200_halo2.hpp"

static int fail(const char* m) { std::printf("FAIL: %s\n", m); return 1; }

// Montgomery form of small integers: x * R mod r computed by repeated doubling of R is overkill here; the test
// uses the library-provided omega plus vectors whose expected Montgomery limbs are known constants.
int main(int argc, char** argv) {
    halo2::EvaluationDomain d(4, 13);
    if (d.extended_k() != 15 || d.get_quotient_poly_degree() != 3) return fail("domain geometry");
    halo2::Fr w3;
    if (zkb_fr_omega(3, w3.data()) != ZKB_OK) return fail("zkb_fr_omega");
    if (argc < 2 || std::strcmp(argv[1], "gpu") != 0) {
        if (zkb_device_count() == 0) {
            try {
                std::vector<halo2::Fr> a(8);
                halo2::best_fft(a, w3, 3);
                return fail("best_fft must throw without a CUDA device");
            } catch (const std::runtime_error&) {
            }
        }
        std::printf("cpu ok\n");
        return 0;
    }
    // Montgomery limbs of 1..8 (a * 2^256 mod r), precomputed by oracle/pyref.py
    const uint64_t in[8][4] = {
        {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
        {0x592c68389ffffff6ULL, 0x6df8ed2b3ec19a53ULL, 0xccdd46def0f28c5cULL, 0x1c14ef83340fbe5eULL},
        {0x05c29c54effffff1ULL, 0xa4f563c0de22677dULL, 0x334bea4e696bd28aULL, 0x2a1f6744ce179d8eULL},
        {0x6e76dadd4fffffebULL, 0xb3bdf20e03c9c415ULL, 0xe16a48076063c05bULL, 0x07c5909386eddc93ULL},
        {0x1b0d0ef99fffffe6ULL, 0xeaba68a3a32a913fULL, 0x47d8eb76d8dd0689ULL, 0x15d0085520f5bbc3ULL},
        {0xc7a34315efffffe1ULL, 0x21b6df39428b5e68ULL, 0xae478ee651564cb8ULL, 0x23da8016bafd9af2ULL},
        {0x3057819e4fffffdbULL, 0x307f6d866832bb01ULL, 0x5c65ec9f484e3a89ULL, 0x0180a96573d3d9f8ULL},
        {0xdcedb5ba9fffffd6ULL, 0x677be41c0793882aULL, 0xc2d4900ec0c780b7ULL, 0x0f8b21270ddbb927ULL}};
    std::vector<halo2::Fr> a(8);
    for (int i = 0; i < 8; ++i) std::memcpy(a[i].data(), in[i], 32);
    halo2::best_fft(a, w3, 3);
    const uint64_t want1[4] = {0x1069f4287460cb5fULL, 0xea22dd8c9b017fc5ULL, 0xc9cdfb2b2395711eULL, 0x2758db28a5c09fddULL};
    if (std::memcmp(a[1].data(), want1, 32) != 0) return fail("best_fft KAT out[1]");
    halo2::EvaluationDomain d2(4, 3);
    std::vector<halo2::Fr> c(8);
    for (int i = 0; i < 8; ++i) std::memcpy(c[i].data(), in[i], 32);
    auto ext = d2.coeff_to_extended(c);
    if (ext.size() != 32) return fail("extended_len");
    auto back = d2.extended_to_coeff(ext);
    if (back.size() != 24) return fail("extended_to_coeff truncation");
    for (int i = 0; i < 8; ++i)
        if (std::memcmp(back[i].data(), in[i], 32) != 0) return fail("coset round trip");
    for (int i = 8; i < 24; ++i)
        for (int j = 0; j < 4; ++j)
            if (back[i][j]) return fail("coset round trip tail");
    // ParamsKZG::setup on the device + commit in both bases: commit(coeffs) == commit_lagrange(evaluations)
    {
        halo2::Fr s;
        std::memcpy(s.data(), in[6], 32);  // "sampled" scalar 7
        halo2::ParamsKZG params = halo2::ParamsKZG::setup(3, s);
        auto g = params.get_g();
        const uint64_t gen_x[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};  // R mod p
        if (std::memcmp(g[0].data(), gen_x, 32) != 0) return fail("g[0] must be the generator (1, 2)");
        halo2::G1 cm = params.commit(c);
        std::vector<halo2::Fr> ev = c;
        halo2::best_fft(ev, w3, 3);  // coeff_to_lagrange
        halo2::G1 cl = params.commit_lagrange(ev);
        if (std::memcmp(cm.data(), cl.data(), 96) != 0) return fail("commit(coeffs) != commit_lagrange(evals)");
        std::vector<halo2::G1> pts{cm, cl};
        auto aff = halo2::batch_normalize(pts);
        if (std::memcmp(aff[0].data(), cm.data(), 64) != 0) return fail("batch_normalize of a z = R point");
    }
    // GraphEvaluator over resident polynomials: values = advice[0](row) * advice[0](row + 1), rot_scale 1, against the
    // element-wise product of the column with its rotated copy (zkb_poly_mul)
    {
        halo2::Polynomial col(c);
        std::vector<halo2::Fr> rot(8);
        for (int i = 0; i < 8; ++i) rot[i] = c[(i + 1) % 8];
        halo2::Polynomial want(c), rotp(rot);
        want.mul(rotp);
        halo2::GraphEvaluator g;
        const uint32_t r0 = g.add_rotation(0), r1 = g.add_rotation(1);
        auto x = g.add_calculation(ZKB_CALC_STORE, halo2::ValueSource::Advice(0, r0));
        g.add_calculation(ZKB_CALC_MUL, x, halo2::ValueSource::Advice(0, r1));
        halo2::Polynomial values = halo2::Polynomial::zeros(8);
        g.evaluate(values, {}, {&col}, {}, {}, {}, 1);
        if (values.to_vec() != want.to_vec()) return fail("GraphEvaluator product of rotations");
        // Horner from the previous value: values = values * y + col  (y = 2)
        halo2::GraphEvaluator h;
        h.add_horner(halo2::ValueSource::PreviousValue(), {halo2::ValueSource::Advice(0, h.add_rotation(0))}, halo2::ValueSource::Y());
        halo2::Fr two;
        std::memcpy(two.data(), in[1], 32);
        halo2::GraphEvaluator::Scalars sc;
        sc.y = &two;
        halo2::Polynomial v2(c);                        // previous = col -> 2 col + col = 3 col
        h.evaluate(v2, {}, {&col}, {}, {}, sc, 1);
        halo2::GraphEvaluator t;                        // 3 * col through a constant
        halo2::Fr three;
        std::memcpy(three.data(), in[2], 32);
        t.add_calculation(ZKB_CALC_MUL, t.add_constant(three), halo2::ValueSource::Advice(0, t.add_rotation(0)));
        halo2::Polynomial v3 = halo2::Polynomial::zeros(8);
        t.evaluate(v3, {}, {&col}, {}, {}, {}, 1);
        if (v2.to_vec() != v3.to_vec()) return fail("GraphEvaluator Horner from the previous value");
        try {
            halo2::Polynomial small = halo2::Polynomial::zeros(4);
            g.evaluate(small, {}, {&col}, {}, {}, {}, 1);
            return fail("a column of another size must throw");
        } catch (const std::runtime_error&) {
        }
    }
    // round-2 interface: lookup permute_expression_pair, blinding-row write, periodic multiply, polynomial file -> HBM, device list
    {
        if (halo2::bound_devices().empty()) return fail("bound_devices after use");
        // input (1, 2, 2, 1, 3, 1, .., ..) against table (3, 1, 2, 2, 1, 1, .., ..), 6 usable rows:
        //   A' = 1 1 1 2 2 3;  first occurrences at rows 0, 3, 5 -> S' = 1 . . 2 . 3; leftovers 1 1 2 go to rows 4, 2, 1 (popped from the end)
        auto fr = [&](int v) { halo2::Fr x; std::memcpy(x.data(), in[v - 1], 32); return x; };
        std::vector<halo2::Fr> inp{fr(1), fr(2), fr(2), fr(1), fr(3), fr(1), fr(8), fr(8)}, tab{fr(3), fr(1), fr(2), fr(2), fr(1), fr(1), fr(7), fr(7)};
        halo2::Polynomial pi(inp), pt(tab);
        auto pr = pi.permute_expression_pair(pt, 6);
        auto a = pr.first.to_vec(), t = pr.second.to_vec();
        const int want_a[6] = {1, 1, 1, 2, 2, 3}, want_t[6] = {1, 2, 1, 2, 1, 3};
        for (int i = 0; i < 6; ++i)
            if (a[i] != fr(want_a[i]) || t[i] != fr(want_t[i])) return fail("permute_expression_pair");
        if (a[6] != halo2::Fr{0, 0, 0, 0} || t[7] != halo2::Fr{0, 0, 0, 0}) return fail("rows past usable_rows must stay zero");
        try {
            std::vector<halo2::Fr> bad = inp;
            bad[0] = fr(5);
            halo2::Polynomial pb(bad);
            pb.permute_expression_pair(pt, 6);
            return fail("an input value outside the table must throw (ConstraintSystemFailure)");
        } catch (const std::runtime_error&) {
        }
        pr.first.write(6, {fr(4), fr(5)});
        auto a2 = pr.first.to_vec();
        if (a2[6] != fr(4) || a2[7] != fr(5) || a2[5] != fr(3)) return fail("Polynomial::write");
        // periodic multiply by (1, 2): element i times (i even ? 1 : 2) == col + (col masked to odd rows)
        halo2::Polynomial pm(inp);
        pm.mul_periodic({fr(1), fr(2)});
        auto m = pm.to_vec();
        for (int i = 0; i < 8; ++i)
            if (m[i] != (i & 1 ? halo2::fr::add(inp[i], inp[i]) : inp[i])) return fail("mul_periodic");
        // raw limbs from a file straight into a handle
        const char* path = "/tmp/zkb200_cpp_poly.bin";
        if (FILE* f = std::fopen(path, "wb")) {
            std::fwrite("HEAD", 1, 4, f);
            std::fwrite(inp[0].data(), 32, inp.size(), f);
            std::fclose(f);
        } else return fail("cannot write the temporary polynomial file");
        halo2::Polynomial pf = halo2::Polynomial::load_file(path, 4, 8);
        if (pf.to_vec() != inp) return fail("Polynomial::load_file");
        std::remove(path);
        try {
            halo2::Polynomial::load_file(path, 0, 8);
            return fail("a missing file must throw");
        } catch (const std::runtime_error&) {
        }
    }
    std::printf("gpu ok\n");
    return 0;
}
