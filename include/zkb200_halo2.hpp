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
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <utility>

#include "zkb200.h"

namespace halo2 {

using Fr = std::array<uint64_t, 4>;        // bn256::Fr
using G1Affine = std::array<uint64_t, 8>;  // bn256::G1Affine (x, y); identity (0, 0)
using G1 = std::array<uint64_t, 12>;       // bn256::G1 (x, y, z) Jacobian

inline void check(int rc, const char* what) {
    if (rc != ZKB_OK) throw std::runtime_error(std::string(what) + ": " + zkb_last_error());
}

// Process set-up: bind the library to `devices` (CUDA ordinals; empty = ZKB_DEVICES or the current device).  With several devices
// the ONE prover process (create_proof, /root/reference/aggregator/src/wrapper.rs:129-137) reaches all of them behind the same
// functions below: commits are sharded by SRS point range, batches by column, one large transform over NVLink peer memory.
inline void init(const std::vector<int>& devices = {}) {
    check(zkb_init(devices.empty() ? nullptr : devices.data(), (int)devices.size()), "init");
}
// the calling thread acts on `device` alone from now on (one host thread per device keeps resident work everywhere)
inline void bind_thread_to_device(int device) { check(zkb_thread_bind_device(device), "bind_thread_to_device"); }
inline std::vector<int> bound_devices() {
    int d[8];
    const int n = zkb_bound_devices(d, 8);
    return std::vector<int>(d, d + (n > 8 ? 8 : n));
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

// ---- host arithmetic on bn256::Fr Montgomery limbs: only for the constants of a graph (negated constants, powers of DELTA) ------
namespace fr {
constexpr Fr MOD = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
constexpr Fr R2 = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
constexpr Fr ONE = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
constexpr Fr DELTA = {0x9a0c322befd78855ull, 0x46e82d14249b563cull, 0x5983a663e0b0b7a7ull, 0x22ab452baaa111adull};  // 7^(2^28)
constexpr uint64_t INV = 0xc2e1f593efffffffull;  // -r^-1 mod 2^64
inline bool geq(const Fr& a, const Fr& b) {
    for (int i = 3; i >= 0; --i)
        if (a[i] != b[i]) return a[i] > b[i];
    return true;
}
inline Fr sub_raw(const Fr& a, const Fr& b) {
    Fr r;
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned __int128 d = (unsigned __int128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    return r;
}
inline Fr add(const Fr& a, const Fr& b) {
    Fr r;
    unsigned __int128 carry = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned __int128 t = (unsigned __int128)a[i] + b[i] + carry;
        r[i] = (uint64_t)t;
        carry = t >> 64;
    }
    return geq(r, MOD) ? sub_raw(r, MOD) : r;  // 2r < 2^256: no carry out
}
inline Fr neg(const Fr& a) { return a == Fr{0, 0, 0, 0} ? a : sub_raw(MOD, a); }
inline Fr mul(const Fr& a, const Fr& b) {  // CIOS
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (unsigned __int128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * INV;
        c = (unsigned __int128)m * MOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (unsigned __int128)m * MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r = {t[0], t[1], t[2], t[3]};
    return (t[4] || geq(r, MOD)) ? sub_raw(r, MOD) : r;
}
inline Fr from_u64(uint64_t v) { return mul(Fr{v, 0, 0, 0}, R2); }
}  // namespace fr

// plonk::Expression without selectors (substituted before evaluate_h runs)
struct Expression {
    enum Kind { Constant, Fixed, Advice, Instance, Challenge, Negated, Sum, Product, Scaled } kind;
    Fr constant{};           // Constant: the value; Scaled: the factor (Montgomery)
    uint32_t index = 0;      // column / challenge
    int32_t rotation = 0;
    std::shared_ptr<const Expression> a, b;
    using Ptr = std::shared_ptr<const Expression>;
    static Ptr constant_(const Fr& c) { auto e = std::make_shared<Expression>(); e->kind = Constant; e->constant = c; return e; }
    static Ptr query(Kind k, uint32_t column, int32_t rot) { auto e = std::make_shared<Expression>(); e->kind = k; e->index = column; e->rotation = rot; return e; }
    static Ptr challenge(uint32_t i) { auto e = std::make_shared<Expression>(); e->kind = Challenge; e->index = i; return e; }
    static Ptr neg(Ptr x) { auto e = std::make_shared<Expression>(); e->kind = Negated; e->a = std::move(x); return e; }
    static Ptr sum(Ptr x, Ptr y) { auto e = std::make_shared<Expression>(); e->kind = Sum; e->a = std::move(x); e->b = std::move(y); return e; }
    static Ptr product(Ptr x, Ptr y) { auto e = std::make_shared<Expression>(); e->kind = Product; e->a = std::move(x); e->b = std::move(y); return e; }
    static Ptr scaled(Ptr x, const Fr& f) { auto e = std::make_shared<Expression>(); e->kind = Scaled; e->a = std::move(x); e->constant = f; return e; }
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
    // ProvingKey::read(.., RawBytesUnchecked), polynomial by polynomial: n raw Montgomery Fr from the file straight into HBM
    // (/root/reference/aggregator/src/wrapper.rs:970-988)
    static Polynomial load_file(const std::string& path, uint64_t offset, size_t n) {
        Polynomial p;
        check(zkb_poly_load_file(path.c_str(), offset, n, &p.h_), "Polynomial::load_file");
        return p;
    }
    // self[offset ..] = values: the random blinding rows at the end of a product column computed on the device
    void write(size_t offset, const std::vector<Fr>& values) {
        check(zkb_poly_write(h_, offset, values.empty() ? nullptr : values[0].data(), values.size()), "write");
    }
    // EvaluationDomain::divide_by_vanishing_poly from the 2^(extended_k - k) distinct values of 1 / (X^n - 1) on the coset
    void mul_periodic(const std::vector<Fr>& table) {
        check(zkb_poly_mul_periodic(h_, table.empty() ? nullptr : table[0].data(), (uint32_t)table.size()), "mul_periodic");
    }
    // plonk::lookup::prover::permute_expression_pair(input = *this, table) on the usable rows -> (A', S'); throws where upstream
    // returns Error::ConstraintSystemFailure (an input value that does not occur in the table)
    std::pair<Polynomial, Polynomial> permute_expression_pair(const Polynomial& table, size_t usable_rows) const {
        Polynomial a, t;
        check(zkb_lookup_permute_expression_pair(h_, table.h_, usable_rows, &a.h_, &t.h_), "permute_expression_pair");
        return {std::move(a), std::move(t)};
    }
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


    // GraphEvaluator::default(): constants [0, 1, 2]
    static GraphEvaluator with_default_constants() {
        GraphEvaluator g;
        g.constants = {Fr{0, 0, 0, 0}, fr::ONE, fr::add(fr::ONE, fr::ONE)};
        return g;
    }
    bool is_constant(const zkb_value_source& v, const Fr& c) const { return v.kind == ZKB_SRC_CONSTANT && constants[v.index] == c; }
    static bool less(const zkb_value_source& x, const zkb_value_source& y) {
        if (x.kind != y.kind) return x.kind < y.kind;
        if (x.index != y.index) return x.index < y.index;
        return x.rotation < y.rotation;
    }
    static bool same_source(const zkb_value_source& x, const zkb_value_source& y) { return x.kind == y.kind && x.index == y.index && x.rotation == y.rotation; }
    // GraphEvaluator::add_expression with upstream's special cases (0, 1, 2, a + (-b) -> Sub, x * x -> Square, ordered operands);
    // needs the default constants
    zkb_value_source add_expression(const Expression& e) {
        const Fr zero{0, 0, 0, 0}, one = fr::ONE, two = fr::add(fr::ONE, fr::ONE);
        switch (e.kind) {
            case Expression::Constant: return add_constant(e.constant);
            case Expression::Fixed: return add_calculation(ZKB_CALC_STORE, ValueSource::Fixed(e.index, add_rotation(e.rotation)));
            case Expression::Advice: return add_calculation(ZKB_CALC_STORE, ValueSource::Advice(e.index, add_rotation(e.rotation)));
            case Expression::Instance: return add_calculation(ZKB_CALC_STORE, ValueSource::Instance(e.index, add_rotation(e.rotation)));
            case Expression::Challenge: return add_calculation(ZKB_CALC_STORE, ValueSource::Challenge(e.index));
            case Expression::Negated: {
                if (e.a->kind == Expression::Constant) return add_constant(fr::neg(e.a->constant));
                const auto r = add_expression(*e.a);
                return is_constant(r, zero) ? r : add_calculation(ZKB_CALC_NEGATE, r);
            }
            case Expression::Sum: {
                if (e.b->kind == Expression::Negated || e.a->kind == Expression::Negated) {
                    const bool b_neg = e.b->kind == Expression::Negated;
                    const auto rp = add_expression(b_neg ? *e.a : *e.b);
                    const auto rn = add_expression(b_neg ? *e.b->a : *e.a->a);
                    if (is_constant(rp, zero)) return add_calculation(ZKB_CALC_NEGATE, rn);
                    if (is_constant(rn, zero)) return rp;
                    return add_calculation(ZKB_CALC_SUB, rp, rn);
                }
                const auto ra = add_expression(*e.a), rb = add_expression(*e.b);
                if (is_constant(ra, zero)) return rb;
                if (is_constant(rb, zero)) return ra;
                return less(rb, ra) ? add_calculation(ZKB_CALC_ADD, rb, ra) : add_calculation(ZKB_CALC_ADD, ra, rb);
            }
            case Expression::Product: {
                const auto ra = add_expression(*e.a), rb = add_expression(*e.b);
                if (is_constant(ra, zero) || is_constant(rb, zero)) return ValueSource::Constant(0);
                if (is_constant(ra, one)) return rb;
                if (is_constant(rb, one)) return ra;
                if (is_constant(ra, two)) return add_calculation(ZKB_CALC_DOUBLE, rb);
                if (is_constant(rb, two)) return add_calculation(ZKB_CALC_DOUBLE, ra);
                if (same_source(ra, rb)) return add_calculation(ZKB_CALC_SQUARE, ra);
                return less(rb, ra) ? add_calculation(ZKB_CALC_MUL, rb, ra) : add_calculation(ZKB_CALC_MUL, ra, rb);
            }
            case Expression::Scaled: {
                if (e.constant == zero) return ValueSource::Constant(0);
                if (e.constant == one) return add_expression(*e.a);
                const auto cst = add_constant(e.constant);
                return add_calculation(ZKB_CALC_MUL, add_expression(*e.a), cst);
            }
        }
        throw std::logic_error("unknown expression");
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

// The custom gates, the permutation argument and one lookup argument as graphs, folded from the previous value with y in
// upstream's term order (the mirror of zksnap-circuits-halo2_b200/evaluation.py, which documents the formulas).
struct ColumnRef {
    uint32_t kind;   // ZKB_SRC_FIXED / ZKB_SRC_ADVICE / ZKB_SRC_INSTANCE
    uint32_t index;
    zkb_value_source at(uint32_t rot_idx) const { return {kind, index, rot_idx}; }
};

inline GraphEvaluator custom_gates_graph(const std::vector<Expression::Ptr>& gate_polys) {
    GraphEvaluator g = GraphEvaluator::with_default_constants();
    std::vector<zkb_value_source> parts;
    for (const auto& e : gate_polys) parts.push_back(g.add_expression(*e));
    g.add_horner(ValueSource::PreviousValue(), parts, ValueSource::Y());
    return g;
}

inline GraphEvaluator permutation_graph(const std::vector<ColumnRef>& columns, size_t chunk_len, int32_t last_rotation, ColumnRef l0,
                                        ColumnRef l_last, ColumnRef l_active, ColumnRef x_coset, const std::vector<ColumnRef>& sigmas,
                                        const std::vector<ColumnRef>& zs) {
    GraphEvaluator g = GraphEvaluator::with_default_constants();
    const uint32_t r0 = g.add_rotation(0), r1 = g.add_rotation(1), rl = g.add_rotation(last_rotation);
    const size_t nsets = (columns.size() + chunk_len - 1) / chunk_len;
    if (zs.size() != nsets || sigmas.size() != columns.size()) throw std::invalid_argument("permutation_graph: one sigma per column, one z per set");
    const auto beta = ValueSource::Beta(), gamma = ValueSource::Gamma(), one = ValueSource::Constant(1);
    auto c = [&](uint32_t op, zkb_value_source a, zkb_value_source b = {ZKB_SRC_CONSTANT, 0, 0}) { return g.add_calculation(op, a, b); };
    std::vector<zkb_value_source> terms;
    terms.push_back(c(ZKB_CALC_MUL, c(ZKB_CALC_SUB, one, zs[0].at(r0)), l0.at(r0)));
    const auto zl = zs.back().at(r0);
    terms.push_back(c(ZKB_CALC_MUL, c(ZKB_CALC_SUB, c(ZKB_CALC_SQUARE, zl), zl), l_last.at(r0)));
    for (size_t i = 1; i < nsets; ++i) terms.push_back(c(ZKB_CALC_MUL, c(ZKB_CALC_SUB, zs[i].at(r0), zs[i - 1].at(rl)), l0.at(r0)));
    const auto bx = c(ZKB_CALC_MUL, beta, x_coset.at(r0));
    Fr delta_pow = fr::ONE;
    size_t j = 0;
    for (size_t i = 0; i < nsets; ++i) {
        auto left = zs[i].at(r1), right = zs[i].at(r0);
        for (size_t k = i * chunk_len; k < columns.size() && k < (i + 1) * chunk_len; ++k, ++j) {
            const auto v = columns[k].at(r0);
            left = c(ZKB_CALC_MUL, left, c(ZKB_CALC_ADD, c(ZKB_CALC_ADD, v, c(ZKB_CALC_MUL, beta, sigmas[j].at(r0))), gamma));
            right = c(ZKB_CALC_MUL, right, c(ZKB_CALC_ADD, c(ZKB_CALC_ADD, v, c(ZKB_CALC_MUL, bx, g.add_constant(delta_pow))), gamma));
            delta_pow = fr::mul(delta_pow, fr::DELTA);
        }
        terms.push_back(c(ZKB_CALC_MUL, c(ZKB_CALC_SUB, left, right), l_active.at(r0)));
    }
    g.add_horner(ValueSource::PreviousValue(), terms, ValueSource::Y());
    return g;
}

inline GraphEvaluator lookup_graph(const std::vector<Expression::Ptr>& input_exprs, const std::vector<Expression::Ptr>& table_exprs, ColumnRef l0,
                                   ColumnRef l_last, ColumnRef l_active, ColumnRef z, ColumnRef permuted_input, ColumnRef permuted_table) {
    GraphEvaluator g = GraphEvaluator::with_default_constants();
    const auto theta = ValueSource::Theta(), beta = ValueSource::Beta(), gamma = ValueSource::Gamma(), one = ValueSource::Constant(1);
    auto c = [&](uint32_t op, zkb_value_source a, zkb_value_source b = {ZKB_SRC_CONSTANT, 0, 0}) { return g.add_calculation(op, a, b); };
    auto compress = [&](const std::vector<Expression::Ptr>& exprs) {
        std::vector<zkb_value_source> parts;
        for (const auto& e : exprs) parts.push_back(g.add_expression(*e));
        return g.add_horner(ValueSource::Constant(0), parts, theta);
    };
    const auto A = compress(input_exprs), S = compress(table_exprs);
    const uint32_t r0 = g.add_rotation(0), r1 = g.add_rotation(1), rm1 = g.add_rotation(-1);
    const auto zz = z.at(r0), zw = z.at(r1), ap = permuted_input.at(r0), apm = permuted_input.at(rm1), sp = permuted_table.at(r0);
    const auto d = c(ZKB_CALC_SUB, ap, sp);
    // one statement per calculation wherever two operands are calculations themselves: C++ leaves the evaluation order of
    // function arguments open, and the numbering of the intermediates must not depend on the compiler
    std::vector<zkb_value_source> terms;
    terms.push_back(c(ZKB_CALC_MUL, c(ZKB_CALC_SUB, one, zz), l0.at(r0)));
    terms.push_back(c(ZKB_CALC_MUL, c(ZKB_CALC_SUB, c(ZKB_CALC_SQUARE, zz), zz), l_last.at(r0)));
    const auto l1 = c(ZKB_CALC_MUL, zw, c(ZKB_CALC_ADD, ap, beta));
    const auto l2 = c(ZKB_CALC_ADD, sp, gamma);
    const auto left = c(ZKB_CALC_MUL, l1, l2);
    const auto r1_ = c(ZKB_CALC_MUL, zz, c(ZKB_CALC_ADD, A, beta));
    const auto r2_ = c(ZKB_CALC_ADD, S, gamma);
    const auto right = c(ZKB_CALC_MUL, r1_, r2_);
    terms.push_back(c(ZKB_CALC_MUL, c(ZKB_CALC_SUB, left, right), l_active.at(r0)));
    terms.push_back(c(ZKB_CALC_MUL, d, l0.at(r0)));
    terms.push_back(c(ZKB_CALC_MUL, c(ZKB_CALC_MUL, d, c(ZKB_CALC_SUB, ap, apm)), l_active.at(r0)));
    g.add_horner(ValueSource::PreviousValue(), terms, ValueSource::Y());
    return g;
}

// plonk::evaluation::Evaluator::evaluate_h over resident cosets: custom gates, then the permutation argument, then every lookup,
// each folded into `values` with y (the mirror of evaluation.QuotientEvaluator).  The auxiliary cosets are appended to the column
// lists: fixed + {l_0, l_last, l_active, X coset, sigmas...}, advice + {z_i...} resp. advice + {z, a', s'} per lookup.
struct PermutationArgument {
    std::vector<ColumnRef> columns;
    size_t chunk_len = 2;          // cs.degree() - 2
    int32_t last_rotation = -1;    // -(blinding_factors + 1)
};
struct LookupArgument {
    std::vector<Expression::Ptr> input_expressions, table_expressions;
};
struct LookupCosets {
    const Polynomial* product;
    const Polynomial* permuted_input;
    const Polynomial* permuted_table;
};
class QuotientEvaluator {
   public:
    QuotientEvaluator(const std::vector<Expression::Ptr>& gate_polys, const PermutationArgument* permutation, std::vector<LookupArgument> lookups)
        : has_gates_(!gate_polys.empty()), custom_gates_(gate_polys.empty() ? GraphEvaluator() : custom_gates_graph(gate_polys)),
          has_permutation_(permutation != nullptr), permutation_(permutation ? *permutation : PermutationArgument()), lookups_(std::move(lookups)) {}

    void evaluate_h(Polynomial& values, std::vector<const Polynomial*> fixed, std::vector<const Polynomial*> advice,
                    const std::vector<const Polynomial*>& instance, const std::vector<Fr>& challenges, const GraphEvaluator::Scalars& sc,
                    int32_t rot_scale, const Polynomial& l0, const Polynomial& l_last, const Polynomial& l_active, const Polynomial* x_coset,
                    const std::vector<const Polynomial*>& sigma_cosets, const std::vector<const Polynomial*>& permutation_product_cosets,
                    const std::vector<LookupCosets>& lookup_cosets) const {
        if (has_gates_) custom_gates_.evaluate(values, fixed, advice, instance, challenges, sc, rot_scale);
        const uint32_t nf = uint32_t(fixed.size()), na = uint32_t(advice.size());
        std::vector<const Polynomial*> fx = fixed;
        fx.insert(fx.end(), {&l0, &l_last, &l_active});
        const ColumnRef L0{ZKB_SRC_FIXED, nf}, LL{ZKB_SRC_FIXED, nf + 1}, LA{ZKB_SRC_FIXED, nf + 2};
        if (has_permutation_) {
            const size_t ncols = permutation_.columns.size(), nsets = (ncols + permutation_.chunk_len - 1) / permutation_.chunk_len;
            if (!x_coset || sigma_cosets.size() != ncols || permutation_product_cosets.size() != nsets)
                throw std::invalid_argument("evaluate_h: the permutation needs the X coset, one sigma coset per column and one product coset per set");
            std::vector<ColumnRef> sig, zs;
            for (uint32_t j = 0; j < ncols; ++j) sig.push_back({ZKB_SRC_FIXED, nf + 4 + j});
            for (uint32_t i = 0; i < nsets; ++i) zs.push_back({ZKB_SRC_ADVICE, na + i});
            const GraphEvaluator g = permutation_graph(permutation_.columns, permutation_.chunk_len, permutation_.last_rotation, L0, LL, LA,
                                                       {ZKB_SRC_FIXED, nf + 3}, sig, zs);
            std::vector<const Polynomial*> f2 = fx, a2 = advice;
            f2.push_back(x_coset);
            f2.insert(f2.end(), sigma_cosets.begin(), sigma_cosets.end());
            a2.insert(a2.end(), permutation_product_cosets.begin(), permutation_product_cosets.end());
            g.evaluate(values, f2, a2, instance, challenges, sc, rot_scale);
        }
        if (lookup_cosets.size() != lookups_.size()) throw std::invalid_argument("evaluate_h: one coset triple per lookup");
        for (size_t n = 0; n < lookups_.size(); ++n) {
            const GraphEvaluator g = lookup_graph(lookups_[n].input_expressions, lookups_[n].table_expressions, L0, LL, LA, {ZKB_SRC_ADVICE, na},
                                                  {ZKB_SRC_ADVICE, na + 1}, {ZKB_SRC_ADVICE, na + 2});
            std::vector<const Polynomial*> a2 = advice;
            a2.insert(a2.end(), {lookup_cosets[n].product, lookup_cosets[n].permuted_input, lookup_cosets[n].permuted_table});
            g.evaluate(values, fx, a2, instance, challenges, sc, rot_scale);
        }
    }

   private:
    bool has_gates_;
    GraphEvaluator custom_gates_;
    bool has_permutation_;
    PermutationArgument permutation_;
    std::vector<LookupArgument> lookups_;
};

}  // namespace halo2
