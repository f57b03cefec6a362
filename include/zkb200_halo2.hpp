// zkb200_halo2.hpp — C++ host-side mirror of the halo2-axiom interface on the zksnap hot path, over the C ABI in
// zkb200.h.  The reference's host language is Rust (absent from the build image), so this header is the compiled
// host layer that mirrors the same names, argument meaning and error behaviour (a failed call throws, where the
// Rust callers `.unwrap()` / `.expect()` — /root/reference/aggregator/src/wrapper.rs:107-108,137):
//
//   halo2_proofs::arithmetic::best_multiexp / best_fft
//   halo2_proofs::poly::EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended, extended_to_coeff}
//   halo2_proofs::poly::kzg::commitment::ParamsKZG::{commit, commit_lagrange}
//   halo2_proofs::plonk::evaluation::GraphEvaluator::{add_rotation, add_constant, add_calculation, evaluate}  (row loop of
//       evaluate_h over polynomials resident in HBM)
//
// Types are the in-memory layouts of halo2curves::bn256 (Montgomery limbs).
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "zkb200.h"

namespace halo2 {

using Fr = std::array<uint64_t, 4>;        // bn256::Fr
using G1Affine = std::array<uint64_t, 8>;  // bn256::G1Affine (x, y); identity (0, 0)
using G1 = std::array<uint64_t, 12>;       // bn256::G1 (x, y, z) Jacobian

inline void check(int rc, const char* what) {
    if (rc != ZKB_OK) throw std::runtime_error(std::string(what) + ": " + zkb_last_error());
}

// arithmetic::best_multiexp(coeffs, bases) -> G1
inline G1 best_multiexp(const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
    if (coeffs.size() != bases.size()) throw std::invalid_argument("assertion failed: coeffs.len() == bases.len()");
    G1 out{};
    check(zkb_msm_g1(coeffs.empty() ? nullptr : coeffs[0].data(), bases.empty() ? nullptr : bases[0].data(), coeffs.size(), out.data()),
          "best_multiexp");
    return out;
}

// arithmetic::best_fft(a, omega, log_n)
inline void best_fft(std::vector<Fr>& a, const Fr& omega, uint32_t log_n) {
    if (a.size() != (size_t(1) << log_n)) throw std::invalid_argument("assertion failed: a.len() == 1 << log_n");
    check(zkb_ntt_fr(a[0].data(), omega.data(), log_n), "best_fft");
}

// group::Curve::batch_normalize(&[G1], &mut [G1Affine])
inline std::vector<G1Affine> batch_normalize(const std::vector<G1>& p) {
    std::vector<G1Affine> q(p.size());
    check(zkb_g1_batch_normalize(p.empty() ? nullptr : p[0].data(), p.size(), q.empty() ? nullptr : q[0].data()), "batch_normalize");
    return q;
}

// poly::EvaluationDomain<Fr>
class EvaluationDomain {
   public:
    EvaluationDomain(uint32_t j, uint32_t k) : k_(k), quotient_poly_degree_(j - 1) {
        extended_k_ = k;
        while ((uint64_t(1) << extended_k_) < (uint64_t(1) << k) * quotient_poly_degree_) ++extended_k_;
    }
    uint32_t k() const { return k_; }
    uint32_t extended_k() const { return extended_k_; }
    size_t extended_len() const { return size_t(1) << extended_k_; }
    uint64_t get_quotient_poly_degree() const { return quotient_poly_degree_; }
    Fr get_omega() const { Fr w; check(zkb_fr_omega(k_, w.data()), "get_omega"); return w; }
    Fr get_extended_omega() const { Fr w; check(zkb_fr_omega(extended_k_, w.data()), "get_extended_omega"); return w; }

    std::vector<Fr> lagrange_to_coeff(std::vector<Fr> a) const {
        if (a.size() != (size_t(1) << k_)) throw std::invalid_argument("assertion failed: a.values.len() == 1 << self.k");
        check(zkb_lagrange_to_coeff(a[0].data(), k_), "lagrange_to_coeff");
        return a;
    }
    std::vector<Fr> coeff_to_extended(const std::vector<Fr>& a) const {
        if (a.size() != (size_t(1) << k_)) throw std::invalid_argument("assertion failed: a.values.len() == 1 << self.k");
        std::vector<Fr> out(extended_len());
        check(zkb_coeff_to_extended(a[0].data(), out[0].data(), k_, extended_k_), "coeff_to_extended");
        return out;
    }
    std::vector<Fr> extended_to_coeff(std::vector<Fr> a) const {
        if (a.size() != extended_len()) throw std::invalid_argument("assertion failed: a.values.len() == self.extended_len()");
        check(zkb_extended_to_coeff(a[0].data(), k_, extended_k_), "extended_to_coeff");
        a.resize((size_t(1) << k_) * quotient_poly_degree_);
        return a;
    }

   private:
    uint32_t k_, extended_k_;
    uint64_t quotient_poly_degree_;
};

// poly::kzg::commitment::ParamsKZG — the MSM-facing part: g and g_lagrange resident in HBM
class ParamsKZG {
   public:
    ParamsKZG(uint32_t k, const std::vector<G1Affine>& g, const std::vector<G1Affine>& g_lagrange) : k_(k) {
        if (g.size() != (size_t(1) << k) || g_lagrange.size() != g.size()) throw std::invalid_argument("SRS must hold 2^k points");
        check(zkb_srs_register(g[0].data(), g.size(), &h_g_), "ParamsKZG (g)");
        check(zkb_srs_register(g_lagrange[0].data(), g_lagrange.size(), &h_gl_), "ParamsKZG (g_lagrange)");
    }
    // ParamsKZG::setup(k, rng), G1 side: `s` is the scalar the caller sampled; g and g_lagrange are generated in HBM
    static ParamsKZG setup(uint32_t k, const Fr& s) {
        ParamsKZG p(k);
        check(zkb_kzg_setup_resident(k, s.data(), &p.h_g_, &p.h_gl_), "ParamsKZG::setup");
        return p;
    }
    ParamsKZG(ParamsKZG&& o) noexcept : k_(o.k_), h_g_(o.h_g_), h_gl_(o.h_gl_) { o.h_g_ = o.h_gl_ = 0; }
    std::vector<G1Affine> get_g() const { return download(h_g_); }
    std::vector<G1Affine> get_g_lagrange() const { return download(h_gl_); }
    ParamsKZG(const ParamsKZG&) = delete;
    ParamsKZG& operator=(const ParamsKZG&) = delete;
    ~ParamsKZG() {
        if (h_g_) zkb_srs_release(h_g_);
        if (h_gl_) zkb_srs_release(h_gl_);
    }
    uint32_t k() const { return k_; }
    // commit(&Polynomial<Fr, Coeff>, Blind) — the blind is unused under KZG
    G1 commit(const std::vector<Fr>& poly) const { return msm(h_g_, poly); }
    G1 commit_lagrange(const std::vector<Fr>& poly) const { return msm(h_gl_, poly); }

   private:
    explicit ParamsKZG(uint32_t k) : k_(k) {}
    std::vector<G1Affine> download(uint64_t h) const {
        std::vector<G1Affine> out(size_t(1) << k_);
        check(zkb_srs_download(h, out[0].data(), out.size()), "get_g");
        return out;
    }
    G1 msm(uint64_t h, const std::vector<Fr>& poly) const {
        G1 out{};
        check(zkb_msm_g1_srs(h, poly.empty() ? nullptr : poly[0].data(), poly.size(), out.data()), "commit");
        return out;
    }
    uint32_t k_;
    uint64_t h_g_ = 0, h_gl_ = 0;
};

// A polynomial / column of evaluations resident in HBM (zkb_poly_*): uploaded once, used by handle.
class Polynomial {
   public:
    explicit Polynomial(const std::vector<Fr>& values) { check(zkb_poly_upload(values.empty() ? nullptr : values[0].data(), values.size(), &h_), "Polynomial"); }
    static Polynomial zeros(size_t n) { Polynomial p; check(zkb_poly_alloc(n, &p.h_), "Polynomial::zeros"); return p; }
    Polynomial(Polynomial&& o) noexcept : h_(o.h_) { o.h_ = 0; }
    Polynomial(const Polynomial&) = delete;
    Polynomial& operator=(const Polynomial&) = delete;
    ~Polynomial() { if (h_) zkb_poly_free(h_); }
    uint64_t handle() const { return h_; }
    size_t len() const { size_t n = 0; check(zkb_poly_len(h_, &n), "len"); return n; }
    std::vector<Fr> to_vec() const {
        std::vector<Fr> out(len());
        check(zkb_poly_download(h_, out.empty() ? nullptr : out[0].data(), out.size()), "to_vec");
        return out;
    }
    void mul(const Polynomial& other) { check(zkb_poly_mul(h_, other.h_), "mul"); }
    Polynomial slice(size_t offset, size_t n) const { Polynomial p; check(zkb_poly_slice(h_, offset, n, &p.h_), "slice"); return p; }

   private:
    Polynomial() = default;
    uint64_t h_ = 0;
};

// plonk::evaluation::{ValueSource, Calculation, GraphEvaluator}.  `evaluate` is evaluate_h's row loop
//   values[idx] = graph.evaluate(&mut data, fixed, advice, instance, challenges, &beta, &gamma, &theta, &y, &values[idx], idx, rot_scale, isize)
// for every idx, on the device.  Horner(start, parts, factor) = Store(start) + one MulAdd per part (add_horner).
struct ValueSource {
    static zkb_value_source Constant(uint32_t i) { return {ZKB_SRC_CONSTANT, i, 0}; }
    static zkb_value_source Intermediate(uint32_t i) { return {ZKB_SRC_INTERMEDIATE, i, 0}; }
    static zkb_value_source Fixed(uint32_t col, uint32_t rot) { return {ZKB_SRC_FIXED, col, rot}; }
    static zkb_value_source Advice(uint32_t col, uint32_t rot) { return {ZKB_SRC_ADVICE, col, rot}; }
    static zkb_value_source Instance(uint32_t col, uint32_t rot) { return {ZKB_SRC_INSTANCE, col, rot}; }
    static zkb_value_source Challenge(uint32_t i) { return {ZKB_SRC_CHALLENGE, i, 0}; }
    static zkb_value_source Beta() { return {ZKB_SRC_BETA, 0, 0}; }
    static zkb_value_source Gamma() { return {ZKB_SRC_GAMMA, 0, 0}; }
    static zkb_value_source Theta() { return {ZKB_SRC_THETA, 0, 0}; }
    static zkb_value_source Y() { return {ZKB_SRC_Y, 0, 0}; }
    static zkb_value_source PreviousValue() { return {ZKB_SRC_PREVIOUS, 0, 0}; }
};

class GraphEvaluator {
   public:
    std::vector<Fr> constants;
    std::vector<int32_t> rotations;
    std::vector<zkb_calculation> calculations;
    uint32_t num_intermediates = 0;

    uint32_t add_rotation(int32_t rotation) {
        for (size_t i = 0; i < rotations.size(); ++i)
            if (rotations[i] == rotation) return uint32_t(i);
        rotations.push_back(rotation);
        return uint32_t(rotations.size() - 1);
    }
    zkb_value_source add_constant(const Fr& c) {
        for (size_t i = 0; i < constants.size(); ++i)
            if (constants[i] == c) return ValueSource::Constant(uint32_t(i));
        constants.push_back(c);
        return ValueSource::Constant(uint32_t(constants.size() - 1));
    }
    // op: ZKB_CALC_{ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, STORE}; an identical earlier calculation is reused
    zkb_value_source add_calculation(uint32_t op, zkb_value_source a, zkb_value_source b = {ZKB_SRC_CONSTANT, 0, 0}) {
        auto same = [](const zkb_value_source& x, const zkb_value_source& y) { return x.kind == y.kind && x.index == y.index && x.rotation == y.rotation; };
        for (size_t i : made_by_add_calculation_) {  // a Horner's Store is not a candidate: its target is rewritten by the steps
            const auto& c = calculations[i];
            if (c.op == op && same(c.a, a) && same(c.b, b)) return ValueSource::Intermediate(c.target);
        }
        made_by_add_calculation_.push_back(calculations.size());
        calculations.push_back({op, num_intermediates, a, b, {ZKB_SRC_CONSTANT, 0, 0}});
        return ValueSource::Intermediate(num_intermediates++);
    }
    zkb_value_source add_horner(zkb_value_source start, const std::vector<zkb_value_source>& parts, zkb_value_source factor) {
        const uint32_t t = num_intermediates++;
        calculations.push_back({ZKB_CALC_STORE, t, start, {ZKB_SRC_CONSTANT, 0, 0}, {ZKB_SRC_CONSTANT, 0, 0}});
        for (const auto& p : parts) calculations.push_back({ZKB_CALC_MUL_ADD, t, ValueSource::Intermediate(t), factor, p});
        return ValueSource::Intermediate(t);
    }

    std::vector<size_t> made_by_add_calculation_;

    struct Scalars {
        const Fr* beta = nullptr;
        const Fr* gamma = nullptr;
        const Fr* theta = nullptr;
        const Fr* y = nullptr;
    };
    void evaluate(Polynomial& values, const std::vector<const Polynomial*>& fixed, const std::vector<const Polynomial*>& advice,
                  const std::vector<const Polynomial*>& instance, const std::vector<Fr>& challenges, const Scalars& sc,
                  int32_t rot_scale) const {
        auto handles = [](const std::vector<const Polynomial*>& v) {
            std::vector<uint64_t> h;
            for (auto* p : v) h.push_back(p->handle());
            return h;
        };
        const std::vector<uint64_t> f = handles(fixed), a = handles(advice), in = handles(instance);
        zkb_graph g{calculations.data(), calculations.size(), num_intermediates, constants.empty() ? nullptr : constants[0].data(),
                    constants.size(), rotations.data(), rotations.size()};
        zkb_graph_inputs inp{f.data(), f.size(), a.data(), a.size(), in.data(), in.size(),
                             challenges.empty() ? nullptr : challenges[0].data(), challenges.size(),
                             sc.beta ? sc.beta->data() : nullptr, sc.gamma ? sc.gamma->data() : nullptr,
                             sc.theta ? sc.theta->data() : nullptr, sc.y ? sc.y->data() : nullptr, rot_scale};
        check(zkb_graph_evaluate(&g, &inp, values.handle()), "GraphEvaluator::evaluate");
    }
};

}  // namespace halo2
