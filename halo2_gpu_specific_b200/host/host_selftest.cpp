// host_selftest.cpp -- the reference's own unit tests for this path, run through the C++ mirror
// on the GPU: test_commit_lagrange (poly/commitment.rs:480-495), an iFFT consistency check in the
// spirit of test_rotate (poly/domain.rs:550-589), and the coset round trip.  Exit code 0 = pass.
#include <cstdio>
#include <vector>

#include "halo2_b200.hpp"

using namespace halo2_b200;

static Fr host_eval(const std::vector<Fr>& poly, const Fr& x) {  // arithmetic.rs:707-711, on the CPU
    Fr acc = {{0, 0, 0, 0}};
    for (size_t i = poly.size(); i-- > 0;) {
        acc = fr::mul(acc, x);
        // acc += poly[i]
        unsigned __int128 c = 0;
        uint64_t t[4];
        for (int j = 0; j < 4; j++) { c += (unsigned __int128)acc.l[j] + poly[i].l[j]; t[j] = (uint64_t)c; c >>= 64; }
        if (fr::geq_p(t)) fr::sub_p(t);
        for (int j = 0; j < 4; j++) acc.l[j] = t[j];
    }
    return acc;
}
static bool same_point(const G1& a, const G1& b) { return std::memcmp(&a, &b, sizeof a) == 0; }

int main() {
    if (b2_device_count() < 1) { std::printf("no CUDA device: %s\n", "skipping is not allowed"); return 2; }
    const uint32_t K = 6;
    const uint64_t n = 1ull << K;
    // Params::unsafe_setup (poly/commitment.rs:56-124) with a fixed s; [x]G through a 1-term MSM
    Fr s = fr::from_u64(0x123456789abcdef1ULL);
    G1Affine gen;
    { Fr one_q; (void)one_q; }
    const uint64_t q_one[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
    const uint64_t q_two[4] = {0xa6ba871b8b1e1b3aULL, 0x14f1d651eb8e167bULL, 0xccdd46def0f28c58ULL, 0x1c14ef83340fbe5eULL};
    std::memcpy(gen.x, q_one, 32);
    std::memcpy(gen.y, q_two, 32);
    std::vector<G1Affine> one_base(1, gen);
    auto mul_gen = [&](const Fr& x) {
        std::vector<Fr> c(1, x);
        G1 p = best_multiexp(c, one_base);
        G1Affine a; std::memcpy(a.x, p.x, 32); std::memcpy(a.y, p.y, 32);
        return a;
    };
    std::vector<G1Affine> g(n), gl(n);
    Fr cur = fr::ONE;
    for (uint64_t i = 0; i < n; i++) { g[i] = mul_gen(cur); cur = fr::mul(cur, s); }
    EvaluationDomain domain(1, K);
    Fr multiplier = fr::mul(fr::sub(fr::pow_vartime(s, n), fr::ONE), fr::invert(fr::from_u64(n)));
    for (uint64_t i = 0; i < n; i++) {
        Fr root_pow = fr::pow_vartime(domain.omega, i);
        Fr scalar = fr::mul(fr::mul(multiplier, root_pow), fr::invert(fr::sub(s, root_pow)));
        gl[i] = mul_gen(scalar);
    }
    Params params(K, g, gl);
    {   // Params::unsafe_setup on the device gives the same 2 * 2^k points as the loop above
        Params dev_params = Params::unsafe_setup(K, s);
        std::vector<G1Affine> dg = dev_params.g.read(), dgl = dev_params.g_lagrange.read();
        if (std::memcmp(dg.data(), g.data(), n * sizeof(G1Affine)) != 0) { std::printf("FAIL unsafe_setup g\n"); return 1; }
        if (std::memcmp(dgl.data(), gl.data(), n * sizeof(G1Affine)) != 0) { std::printf("FAIL unsafe_setup g_lagrange\n"); return 1; }
    }
    std::vector<Fr> a(n);
    for (uint64_t i = 0; i < n; i++) a[i] = fr::from_u64(i);
    std::vector<Fr> b = domain.lagrange_to_coeff(a);
    G1 c1 = params.commit_lagrange(a), c2 = params.commit(b);
    if (!same_point(c1, c2)) { std::printf("FAIL test_commit_lagrange\n"); return 1; }
    if (!same_point(params.commit_lagrange_with_bound(a, 6), c1)) { std::printf("FAIL with_bound\n"); return 1; }
    auto both = params.commit_lagrange_and_ifft(a, domain.omega_inv, domain.ifft_divisor);
    if (!same_point(both.second, c1) || std::memcmp(both.first.data(), b.data(), n * 32) != 0) {
        std::printf("FAIL commit_lagrange_and_ifft\n"); return 1;
    }
    // coefficient form evaluates back to the Lagrange values: b(omega^i) == a[i]
    for (uint64_t i = 0; i < n; i += 7) {
        Fr x = fr::pow_vartime(domain.omega, i);
        if (!fr::eq(host_eval(b, x), a[i])) { std::printf("FAIL iFFT consistency at %llu\n", (unsigned long long)i); return 1; }
    }
    // one opening of the multiopen argument (gwc/prover.rs:47-160): fold a and b by v, divide by (X - z):
    // w(r) * (r - z) + batch(z) == batch(r)
    {
        const Fr v = fr::from_u64(0x1234567), z = fr::from_u64(0x89abcdef), r = fr::from_u64(0x31415926);
        std::vector<Fr> batch = poly_combine({&a, &b}, v);
        for (uint64_t i = 0; i < n; i += 5) {   // batch[i] = a[i] * v + b[i]
            std::vector<Fr> two = {b[i], a[i]};
            if (!fr::eq(host_eval(two, v), batch[i])) { std::printf("FAIL poly_combine at %llu\n", (unsigned long long)i); return 1; }
        }
        std::vector<Fr> w = kate_division(batch, z);
        const Fr lhs = fr::mul(halo2_b200::eval_polynomial(w, r), fr::sub(r, z));
        const Fr rhs = fr::sub(halo2_b200::eval_polynomial(batch, r), halo2_b200::eval_polynomial(batch, z));
        if (!fr::eq(lhs, rhs)) { std::printf("FAIL kate_division identity\n"); return 1; }
        if (!fr::eq(halo2_b200::eval_polynomial(batch, r), host_eval(batch, r))) { std::printf("FAIL eval_polynomial\n"); return 1; }
    }
    // coset round trip with j = 5
    EvaluationDomain d5(5, K);
    std::vector<Fr> ext = d5.coeff_to_extended(b);
    std::vector<Fr> back = d5.extended_to_coeff(ext);
    if (back.size() != n * 4 || std::memcmp(back.data(), b.data(), n * 32) != 0) { std::printf("FAIL coset round trip\n"); return 1; }
    for (size_t i = n; i < back.size(); i++)
        if (back[i].l[0] | back[i].l[1] | back[i].l[2] | back[i].l[3]) { std::printf("FAIL coset tail\n"); return 1; }
    // error behaviour: length mismatch panics in the reference -> exception here
    bool threw = false;
    try { std::vector<Fr> bad(5); best_fft(bad, domain.omega, 2); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { std::printf("FAIL expected exception\n"); return 1; }
    std::printf("host_selftest ok\n");
    return 0;
}
