// halo2_b200.hpp -- C++ host-side mirror of the reference's Rust interface for the
// polynomial-commitment hot path, over the C ABI in include/b2pcs.h.
//
// The reference is Rust (no toolchain in the build image), so the host side above the ABI is
// written in C++ with the SAME names, argument meaning and error behaviour as
//   halo2_proofs/src/arithmetic.rs      best_multiexp, best_multiexp_gpu_cond, gpu_multiexp*,
//                                       best_fft, gpu_fft, gpu_ifft            (:309-554)
//   halo2_proofs/src/poly/domain.rs     EvaluationDomain                       (:24-423)
//   halo2_proofs/src/poly/commitment.rs Params::{commit, commit_lagrange, commit_lagrange_and_ifft,
//                                       commit_lagrange_with_bound}            (:23-29, :129-222)
// Where the reference panics (.unwrap()/.expect()/assert_eq!) these functions throw
// std::runtime_error.  Vectors are moved in and out exactly where the reference moves the
// Polynomial's Vec.  Header-only; link with -lb2pcs.  No CPU fallback.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b2pcs.h"

namespace halo2_b200 {

struct Fr { uint64_t l[4]; };                 // Montgomery form, as pairing::bn256::Fr in memory
struct G1Affine { uint64_t x[4], y[4]; };     // (0,0) = identity
struct G1 { uint64_t x[4], y[4], z[4]; };     // Jacobian; engine results are normalised (z = 1)

inline void check(int rc, const char* what) {
    if (rc != B2_OK) throw std::runtime_error(std::string(what) + ": " + b2_last_error());
}

// ---------------------------------------------------------------------------------------------
// host-side Fr scalar arithmetic: only for the dozen constants EvaluationDomain::new derives
// (poly/domain.rs:44-149 computes them on the host as well)
namespace fr {
typedef unsigned __int128 u128;
static const uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t INV = 0xc2e1f593efffffffULL;
static const Fr ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const Fr R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const uint32_t S = 28;  // Fr::S

inline bool geq_p(const uint64_t* t) {
    for (int i = 3; i >= 0; i--) {
        if (t[i] > P[i]) return true;
        if (t[i] < P[i]) return false;
    }
    return true;
}
inline void sub_p(uint64_t* t) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)t[i] - P[i] - (uint64_t)br;
        t[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
inline Fr mul(const Fr& a, const Fr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * INV;
        c = (u128)m * P[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq_p(t)) sub_p(t);
    Fr r; std::memcpy(r.l, t, 32); return r;
}
inline Fr sub(const Fr& a, const Fr& b) {
    uint64_t t[4]; u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a.l[i] - b.l[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)t[i] + P[i]; t[i] = (uint64_t)c; c >>= 64; } }
    Fr r; std::memcpy(r.l, t, 32); return r;
}
inline Fr square(const Fr& a) { return mul(a, a); }
inline Fr from_u64(uint64_t v) { Fr x = {{v, 0, 0, 0}}; return mul(x, R2); }   // Fr::from(u64)
inline bool eq(const Fr& a, const Fr& b) { return std::memcmp(a.l, b.l, 32) == 0; }
inline Fr pow_vartime(const Fr& a, uint64_t e) {
    Fr acc = ONE;
    for (int i = 63; i >= 0; i--) { acc = square(acc); if ((e >> i) & 1) acc = mul(acc, a); }
    return acc;
}
inline Fr invert(const Fr& a) {  // a^(r-2)
    uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
    Fr acc = ONE;
    for (int i = 255; i >= 0; i--) { acc = square(acc); if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, a); }
    return acc;
}
// Fr::root_of_unity(): 7^((r-1)/2^28), order 2^28
inline Fr root_of_unity() {
    Fr x = {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};
    return mul(x, R2);
}
// default Fr::ZETA = 7^((r-1)/3); the pinned crate's choice is not recoverable from the tree,
// so EvaluationDomain takes zeta as a parameter
inline Fr zeta_default() {
    Fr x = {{0x8b17ea66b99c90ddULL, 0x5bfc41088d8daaa7ULL, 0xb3c4d79d41a91758ULL, 0x0ULL}};
    return mul(x, R2);
}
}  // namespace fr

// ---------------------------------------------------------------------------------------------
// arithmetic.rs entry points
class Srs {  // bases resident in HBM (replaces the per-call upload of arithmetic.rs:349-360)
  public:
    Srs() = default;
    Srs(const G1Affine* bases, size_t n, bool precompute = true) : n_(n) {
        check(b2_srs_register(bases, n, sizeof(G1Affine), &h_), "b2_srs_register");
        if (precompute) check(b2_srs_precompute(h_, 0), "b2_srs_precompute");
    }
    // bases[i] = [k_i] G for n Montgomery-form scalars already on the device (b2_srs_from_scalars_dev)
    static Srs from_scalars_dev(const void* d_scalars, size_t n, bool precompute = true) {
        Srs s;
        s.n_ = n;
        check(b2_srs_from_scalars_dev(d_scalars, n, &s.h_), "b2_srs_from_scalars_dev");
        if (precompute) check(b2_srs_precompute(s.h_, 0), "b2_srs_precompute");
        return s;
    }
    std::vector<G1Affine> read() const {
        std::vector<G1Affine> out(n_);
        check(b2_srs_read(h_, 0, n_, out.data()), "b2_srs_read");
        return out;
    }
    Srs(const Srs&) = delete;
    Srs& operator=(const Srs&) = delete;
    Srs(Srs&& o) noexcept : h_(o.h_), n_(o.n_) { o.h_ = 0; }
    Srs& operator=(Srs&& o) noexcept { std::swap(h_, o.h_); std::swap(n_, o.n_); return *this; }
    ~Srs() { if (h_) b2_srs_free(h_); }
    b2_handle_t handle() const { return h_; }
    size_t len() const { return n_; }
  private:
    b2_handle_t h_ = 0;
    size_t n_ = 0;
};

inline G1 identity() {
    G1 r; std::memset(&r, 0, sizeof r);
    const uint64_t one[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
    std::memcpy(r.y, one, 32);
    return r;
}
// arithmetic.rs:334-367
inline G1 gpu_multiexp_single_gpu_with_bound(const Fr* coeffs, size_t n, const Srs& bases, size_t max_bits) {
    if (n > bases.len()) throw std::runtime_error("coeffs.len() > bases.len()");
    if (max_bits == 0 || n == 0) return identity();
    G1 out;
    check(b2_msm(bases.handle(), 0, coeffs, n, (uint32_t)max_bits, &out), "b2_msm");
    return out;
}
// arithmetic.rs:370-372, 413-440 (the cross-GPU split lives in the launcher: one process per GPU)
inline G1 gpu_multiexp(const Fr* coeffs, size_t n, const Srs& bases) {
    return gpu_multiexp_single_gpu_with_bound(coeffs, n, bases, 254);
}
// arithmetic.rs:465-492 with host bases (uploaded for this call); panics on length mismatch
inline G1 best_multiexp(const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
    if (coeffs.size() != bases.size()) throw std::runtime_error("assert_eq!(coeffs.len(), bases.len())");
    G1 out;
    check(b2_best_multiexp(coeffs.data(), bases.data(), coeffs.size(), &out), "b2_best_multiexp");
    return out;
}
// arithmetic.rs:442-458
inline G1 best_multiexp_gpu_cond(const Fr* coeffs, size_t n, const Srs& bases) {
    if (n == 0) return identity();
    return gpu_multiexp(coeffs, n, bases);
}
// arithmetic.rs:375-410
inline G1 gpu_multiexp_bound_and_fft(std::vector<Fr>& coeffs, const Srs& bases, size_t max_bits, const Fr& omega,
                                     const Fr& divisor, uint32_t log_n) {
    if (coeffs.size() != ((size_t)1 << log_n)) throw std::runtime_error("coeffs.len() != 1 << log_n");
    G1 out;
    check(b2_msm_and_ifft(bases.handle(), coeffs.data(), (uint32_t)max_bits, &omega, &divisor, log_n, &out),
          "b2_msm_and_ifft");
    return out;
}
// arithmetic.rs:546-554 / 495-512
inline void best_fft(std::vector<Fr>& a, const Fr& omega, uint32_t log_n) {
    if (a.size() != ((size_t)1 << log_n)) throw std::runtime_error("assert_eq!(n, 1 << log_n)");
    check(b2_best_fft(a.data(), &omega, log_n), "b2_best_fft");
}
// arithmetic.rs:515-534
inline void gpu_ifft(std::vector<Fr>& a, const Fr& omega, uint32_t log_n, const Fr& divisor) {
    if (a.size() != ((size_t)1 << log_n)) throw std::runtime_error("assert_eq!(n, 1 << log_n)");
    check(b2_gpu_ifft(a.data(), &omega, log_n, &divisor), "b2_gpu_ifft");
}

// arithmetic.rs:707-735
inline Fr eval_polynomial(const std::vector<Fr>& poly, const Fr& point) {
    if (poly.empty()) return Fr{{0, 0, 0, 0}};
    Fr out;
    check(b2_eval_polynomial(poly.data(), poly.size(), &point, &out), "b2_eval_polynomial");
    return out;
}
// arithmetic.rs:752-773: (a(X) - a(b)) / (X - b)
inline std::vector<Fr> kate_division(const std::vector<Fr>& a, const Fr& b) {
    if (a.size() < 2) throw std::runtime_error("kate_division needs at least two coefficients");
    std::vector<Fr> q(a.size() - 1);
    check(b2_kate_division(a.data(), a.size(), &b, q.data()), "b2_kate_division");
    return q;
}
// poly/multiopen/gwc/prover.rs:47-56: poly_batch = poly_batch * v + poly over the polynomials opened at one point
inline std::vector<Fr> poly_combine(const std::vector<const std::vector<Fr>*>& polys, const Fr& v) {
    if (polys.empty()) throw std::runtime_error("poly_combine needs at least one polynomial");
    std::vector<const void*> ptrs;
    for (const auto* p : polys) {
        if (p->size() != polys[0]->size()) throw std::runtime_error("poly_combine: lengths differ");
        ptrs.push_back(p->data());
    }
    std::vector<Fr> out(polys[0]->size());
    check(b2_poly_combine(ptrs.data(), (uint32_t)ptrs.size(), out.size(), &v, out.data()), "b2_poly_combine");
    return out;
}

// ---------------------------------------------------------------------------------------------
// poly/domain.rs:24-149
class EvaluationDomain {
  public:
    uint64_t n; uint32_t k, extended_k;
    Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv;
    uint64_t quotient_poly_degree;
    Fr ifft_divisor, extended_ifft_divisor, barycentric_weight;
    std::vector<Fr> t_evaluations;

    EvaluationDomain(uint32_t j, uint32_t k_, const Fr& zeta = fr::zeta_default()) {
        quotient_poly_degree = j - 1;                                        // :46
        k = k_; n = 1ull << k;
        extended_k = k;
        while ((1ull << extended_k) < n * quotient_poly_degree) extended_k++;  // :56-59
        if (extended_k > fr::S) throw std::runtime_error("extended_k > Fr::S");
        extended_omega = fr::root_of_unity();
        for (uint32_t i = extended_k; i < fr::S; i++) extended_omega = fr::square(extended_omega);  // :66-68
        omega = extended_omega;
        for (uint32_t i = k; i < extended_k; i++) omega = fr::square(omega);  // :78-80
        g_coset = zeta;                                                        // :88
        g_coset_inv = fr::square(zeta);                                        // :89
        Fr orig = fr::pow_vartime(zeta, n), step = fr::pow_vartime(extended_omega, n), cur = orig;  // :95-96
        do { t_evaluations.push_back(cur); cur = fr::mul(cur, step); } while (!fr::eq(cur, orig));
        if (t_evaluations.size() != (1ull << (extended_k - k))) throw std::runtime_error("t_evaluations length");
        for (auto& t : t_evaluations) t = fr::invert(fr::sub(t, fr::ONE));     // :108-131
        ifft_divisor = fr::invert(fr::from_u64(1ull << k));
        extended_ifft_divisor = fr::invert(fr::from_u64(1ull << extended_k));
        barycentric_weight = fr::invert(fr::from_u64(n));
        extended_omega_inv = fr::invert(extended_omega);
        omega_inv = fr::invert(omega);
    }
    size_t extended_len() const { return (size_t)1 << extended_k; }

    // :233-266 (moves the Vec in and out, as the reference does)
    std::vector<Fr> lagrange_to_coeff(std::vector<Fr> a) const {
        if (a.size() != n) throw std::runtime_error("assert_eq!(a.values.len(), 1 << self.k)");
        gpu_ifft(a, omega_inv, k, ifft_divisor);
        return a;
    }
    std::vector<Fr> lagrange_to_coeff_st(std::vector<Fr> a) const { return lagrange_to_coeff(std::move(a)); }
    // :270-287
    std::vector<Fr> coeff_to_extended(const std::vector<Fr>& a) const {
        if (a.size() != n) throw std::runtime_error("assert_eq!(a.values.len(), 1 << self.k)");
        std::vector<Fr> out(extended_len());
        check(b2_coeff_to_extended(a.data(), out.data(), 1, k, extended_k, &g_coset, &g_coset_inv, &extended_omega),
              "b2_coeff_to_extended");
        return out;
    }
    // :328-350
    std::vector<Fr> extended_to_coeff(const std::vector<Fr>& a) const {
        if (a.size() != extended_len()) throw std::runtime_error("assert_eq!(a.values.len(), self.extended_len())");
        std::vector<Fr> out((size_t)(n * quotient_poly_degree));
        check(b2_extended_to_coeff(a.data(), out.data(), out.size(), extended_k, &g_coset, &g_coset_inv,
                                   &extended_omega_inv, &extended_ifft_divisor), "b2_extended_to_coeff");
        return out;
    }
    // :354-373
    std::vector<Fr> divide_by_vanishing_poly(std::vector<Fr> a) const {
        if (a.size() != extended_len()) throw std::runtime_error("assert_eq!(a.values.len(), self.extended_len())");
        check(b2_divide_by_vanishing_poly(a.data(), extended_k, t_evaluations.data(), (uint32_t)t_evaluations.size()),
              "b2_divide_by_vanishing_poly");
        return a;
    }
};

// ---------------------------------------------------------------------------------------------
// poly/commitment.rs:23-29, 129-222
class Params {
  public:
    uint32_t k; uint64_t n;
    Srs g, g_lagrange;
    Params(uint32_t k_, const std::vector<G1Affine>& g_, const std::vector<G1Affine>& gl_)
        : k(k_), n(1ull << k_), g(g_.data(), g_.size()), g_lagrange(gl_.data(), gl_.size()) {
        if (g_.size() != n || gl_.size() != n) throw std::runtime_error("g / g_lagrange must hold 2^k points");
    }
    Params(uint32_t k_, Srs&& g_, Srs&& gl_) : k(k_), n(1ull << k_), g(std::move(g_)), g_lagrange(std::move(gl_)) {
        if (g.len() != n || g_lagrange.len() != n) throw std::runtime_error("g / g_lagrange must hold 2^k points");
    }
    // Params::unsafe_setup (:56-124) on the device for a caller-chosen s: g[i] = [s^i] G (:63-83),
    // g_lagrange[i] = [(s^n - 1)/n * w^i / (s - w^i)] G (:85-112).  Scalars by prefix product, batch inversion and
    // element-wise passes on resident vectors; points by one fixed-base multiplication each.
    static Params unsafe_setup(uint32_t k, const Fr& s, bool precompute = true) {
        if (k > fr::S) throw std::runtime_error("assert!(k <= Fr::S)");
        const uint64_t n = 1ull << k;
        struct Dev {
            void* p = nullptr;
            explicit Dev(size_t bytes) { check(b2_dev_alloc(bytes, &p), "b2_dev_alloc"); }
            ~Dev() { if (p) b2_dev_free(p); }
        } cst(n * 32), pw(n * 32), den(n * 32);
        auto fill = [&](const Fr& v) {
            std::vector<Fr> host(n, v);
            check(b2_memcpy_h2d(cst.p, host.data(), n * 32), "b2_memcpy_h2d");
        };
        auto powers = [&](const Fr& base, void* out) {      // out[i] = base^i
            fill(base);
            check(b2_prefix_scan_dev(0, cst.p, n, &fr::ONE, nullptr, out, n, nullptr), "b2_prefix_scan_dev");
        };
        powers(s, pw.p);
        Srs g = Srs::from_scalars_dev(pw.p, n, precompute);
        Fr root = fr::root_of_unity();
        for (uint32_t i = k; i < fr::S; i++) root = fr::square(root);
        powers(root, pw.p);
        fill(s);
        check(b2_fr_vec_dev(2, cst.p, pw.p, n, den.p, nullptr), "b2_fr_vec_dev");       // s - w^i
        check(b2_batch_invert_dev(den.p, n, nullptr), "b2_batch_invert_dev");
        check(b2_fr_vec_dev(0, pw.p, den.p, n, den.p, nullptr), "b2_fr_vec_dev");       // w^i / (s - w^i)
        fill(fr::mul(fr::sub(fr::pow_vartime(s, n), fr::ONE), fr::invert(fr::from_u64(n))));
        check(b2_fr_vec_dev(0, den.p, cst.p, n, den.p, nullptr), "b2_fr_vec_dev");      // * (s^n - 1) / n
        Srs gl = Srs::from_scalars_dev(den.p, n, precompute);
        return Params(k, std::move(g), std::move(gl));
    }
    G1 commit(const std::vector<Fr>& poly) const {                 // :129-133
        if (g.len() < poly.size()) throw std::runtime_error("assert!(self.g.len() >= size)");
        return best_multiexp_gpu_cond(poly.data(), poly.size(), g);
    }
    G1 commit_lagrange(const std::vector<Fr>& poly) const {        // :138-142
        if (g.len() < poly.size()) throw std::runtime_error("assert!(self.g.len() >= size)");
        return best_multiexp_gpu_cond(poly.data(), poly.size(), g_lagrange);
    }
    // :199-222: zero scalars produce no bucket entries, so the CPU-side filtering is not needed
    G1 commit_lagrange_with_bound(const std::vector<Fr>& poly, size_t max_bits) const {
        return gpu_multiexp_single_gpu_with_bound(poly.data(), poly.size(), g_lagrange, max_bits);
    }
    // :144-170
    std::pair<std::vector<Fr>, G1> commit_lagrange_and_ifft(std::vector<Fr> poly, const Fr& omega_inv,
                                                            const Fr& ifft_divisor) const {
        G1 c = gpu_multiexp_bound_and_fft(poly, g_lagrange, 254, omega_inv, ifft_divisor, k);
        return {std::move(poly), c};
    }
};

}  // namespace halo2_b200
