// fp.cuh -- 256-bit Montgomery field arithmetic for BN254 Fr / Fq on 8 x 32-bit limbs.
//
// Device-only, sm_100a.  Element layout in memory is the reference's: 4 x u64
// little-endian limbs in Montgomery form a*2^256 mod p (halo2_proofs/src/plonk/
// prover.rs:176 transmutes Fr to [u64;4]; arithmetic.rs:351,391,507 transmute the
// slices handed to the GPU), which on a little-endian machine is the same bytes as
// 8 x u32.
//
// Multiplication is a row-interleaved CIOS Montgomery product that keeps two
// carry-save accumulators: products a[j]*b[i] with even j land on (lo,hi) limb
// pairs that tile the even accumulator without overlap, odd j tile the odd one, so
// every 32x32->64 product is a mad.lo.cc/madc.hi.cc pair on one carry chain (ptxas
// fuses each pair into one IMAD.WIDE with carry in/out).  After each row the lowest
// limb is zeroed by adding m*p and the accumulators swap roles (a one-limb shift
// turns the odd alignment into the even one).  The modulus limbs are compile-time
// immediates.
#pragma once
#include <cstdint>

namespace b2 {

struct FrParams {
    static constexpr uint32_t p0 = 0xf0000001u, p1 = 0x43e1f593u, p2 = 0x79b97091u, p3 = 0x2833e848u,
                              p4 = 0x8181585du, p5 = 0xb85045b6u, p6 = 0xe131a029u, p7 = 0x30644e72u;
    static constexpr uint32_t inv = 0xefffffffu;  // -p^{-1} mod 2^32
    static constexpr uint32_t one0 = 0x4ffffffbu, one1 = 0xac96341cu, one2 = 0x9f60cd29u, one3 = 0x36fc7695u,
                              one4 = 0x7879462eu, one5 = 0x666ea36fu, one6 = 0x9a07df2fu, one7 = 0x0e0a77c1u;
    static constexpr uint32_t rr0 = 0xae216da7u, rr1 = 0x1bb8e645u, rr2 = 0xe35c59e3u, rr3 = 0x53fe3ab1u,
                              rr4 = 0x53bb8085u, rr5 = 0x8c49833du, rr6 = 0x7f4e44a5u, rr7 = 0x0216d0b1u;
};
struct FqParams {
    static constexpr uint32_t p0 = 0xd87cfd47u, p1 = 0x3c208c16u, p2 = 0x6871ca8du, p3 = 0x97816a91u,
                              p4 = 0x8181585du, p5 = 0xb85045b6u, p6 = 0xe131a029u, p7 = 0x30644e72u;
    static constexpr uint32_t inv = 0xe4866389u;
    static constexpr uint32_t one0 = 0xc58f0d9du, one1 = 0xd35d438du, one2 = 0xf5c70b3du, one3 = 0x0a78eb28u,
                              one4 = 0x7879462cu, one5 = 0x666ea36fu, one6 = 0x9a07df2fu, one7 = 0x0e0a77c1u;
    static constexpr uint32_t rr0 = 0x538afa89u, rr1 = 0xf32cfc5bu, rr2 = 0xd44501fbu, rr3 = 0xb5e71911u,
                              rr4 = 0x0a417ff6u, rr5 = 0x47ab1effu, rr6 = 0xcab8351fu, rr7 = 0x06d89f71u;
};

template <class P>
struct __align__(16) Fp {
    uint32_t v[8];

    __device__ __forceinline__ static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    __device__ __forceinline__ static Fp one() {
        Fp r;
        r.v[0] = P::one0; r.v[1] = P::one1; r.v[2] = P::one2; r.v[3] = P::one3;
        r.v[4] = P::one4; r.v[5] = P::one5; r.v[6] = P::one6; r.v[7] = P::one7;
        return r;
    }
    __device__ __forceinline__ static Fp r2() {
        Fp r;
        r.v[0] = P::rr0; r.v[1] = P::rr1; r.v[2] = P::rr2; r.v[3] = P::rr3;
        r.v[4] = P::rr4; r.v[5] = P::rr5; r.v[6] = P::rr6; r.v[7] = P::rr7;
        return r;
    }
    __device__ __forceinline__ bool is_zero() const {
        return (v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) == 0;
    }
    __device__ __forceinline__ bool operator==(const Fp& o) const {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= v[i] ^ o.v[i];
        return d == 0;
    }
    __device__ __forceinline__ bool operator!=(const Fp& o) const { return !(*this == o); }
};

// ---- 128-bit global memory access (two LDG.128 / STG.128 per element) ----------------
template <class P>
__device__ __forceinline__ Fp<P> fp_load(const void* ptr) {
    const uint4* q = reinterpret_cast<const uint4*>(ptr);
    uint4 lo = q[0], hi = q[1];
    Fp<P> r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
template <class P>
__device__ __forceinline__ Fp<P> fp_load_nc(const void* ptr) {  // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(ptr);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    Fp<P> r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
template <class P>
__device__ __forceinline__ void fp_store(void* ptr, const Fp<P>& a) {
    uint4* q = reinterpret_cast<uint4*>(ptr);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// ---- add / sub / neg ---------------------------------------------------------------
// r = a - p if a >= p else a   (a < 2p)
template <class P>
__device__ __forceinline__ void fp_reduce_once(uint32_t (&a)[8]) {
    uint32_t t[8], borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
          "=r"(borrow)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "n"(P::p0), "n"(P::p1), "n"(P::p2), "n"(P::p3), "n"(P::p4), "n"(P::p5), "n"(P::p6), "n"(P::p7));
    // borrow == 0xffffffff when a < p
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : t[i];
}

template <class P>
__device__ __forceinline__ Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    fp_reduce_once<P>(r.v);  // a + b < 2p < 2^255: no carry out of limb 7
    return r;
}

template <class P>
__device__ __forceinline__ Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // borrow = 0xffffffff if a < b: add p back (masked)
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7])
        : "r"(borrow & P::p0), "r"(borrow & P::p1), "r"(borrow & P::p2), "r"(borrow & P::p3),
          "r"(borrow & P::p4), "r"(borrow & P::p5), "r"(borrow & P::p6), "r"(borrow & P::p7));
    return r;
}

template <class P>
__device__ __forceinline__ Fp<P> fp_neg(const Fp<P>& a) {
    if (a.is_zero()) return a;
    Fp<P> p;
    p.v[0] = P::p0; p.v[1] = P::p1; p.v[2] = P::p2; p.v[3] = P::p3;
    p.v[4] = P::p4; p.v[5] = P::p5; p.v[6] = P::p6; p.v[7] = P::p7;
    Fp<P> r;
    asm("sub.cc.u32 %0, %8, %16;\n\t"
        "subc.cc.u32 %1, %9, %17;\n\t"
        "subc.cc.u32 %2, %10, %18;\n\t"
        "subc.cc.u32 %3, %11, %19;\n\t"
        "subc.cc.u32 %4, %12, %20;\n\t"
        "subc.cc.u32 %5, %13, %21;\n\t"
        "subc.cc.u32 %6, %14, %22;\n\t"
        "subc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(p.v[0]), "r"(p.v[1]), "r"(p.v[2]), "r"(p.v[3]), "r"(p.v[4]), "r"(p.v[5]), "r"(p.v[6]), "r"(p.v[7]),
          "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
    return r;
}

template <class P>
__device__ __forceinline__ Fp<P> fp_dbl(const Fp<P>& a) { return fp_add<P>(a, a); }

// ---- Montgomery multiplication -------------------------------------------------------
// Row 0: E = a[even]*b0 (positions 0..7), O = a[odd]*b0 (positions 1..8).
__device__ __forceinline__ void mm_row0(uint32_t (&E)[8], uint32_t (&O)[8], const uint32_t (&a)[8], uint32_t b) {
    asm("mul.lo.u32 %0, %16, %24;\n\t mul.hi.u32 %1, %16, %24;\n\t"
        "mul.lo.u32 %2, %18, %24;\n\t mul.hi.u32 %3, %18, %24;\n\t"
        "mul.lo.u32 %4, %20, %24;\n\t mul.hi.u32 %5, %20, %24;\n\t"
        "mul.lo.u32 %6, %22, %24;\n\t mul.hi.u32 %7, %22, %24;\n\t"
        "mul.lo.u32 %8, %17, %24;\n\t mul.hi.u32 %9, %17, %24;\n\t"
        "mul.lo.u32 %10, %19, %24;\n\t mul.hi.u32 %11, %19, %24;\n\t"
        "mul.lo.u32 %12, %21, %24;\n\t mul.hi.u32 %13, %21, %24;\n\t"
        "mul.lo.u32 %14, %23, %24;\n\t mul.hi.u32 %15, %23, %24;"
        : "=r"(E[0]), "=r"(E[1]), "=r"(E[2]), "=r"(E[3]), "=r"(E[4]), "=r"(E[5]), "=r"(E[6]), "=r"(E[7]),
          "=r"(O[0]), "=r"(O[1]), "=r"(O[2]), "=r"(O[3]), "=r"(O[4]), "=r"(O[5]), "=r"(O[6]), "=r"(O[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b));
}

// Row i >= 1.  On entry E is the accumulator aligned at the current position 0 and O is
// the previous row's even accumulator (its limb 0 is zero, limb 1 sits at position 0).
// Folds O[1] into E[0], shifts O down one 64-bit slot while adding a[odd]*b, adds
// a[even]*b to E, and sends E's carry-out to O[7] (position 8).
__device__ __forceinline__ void mm_row(uint32_t (&E)[8], uint32_t (&O)[8], const uint32_t (&a)[8], uint32_t b) {
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "madc.lo.cc.u32 %8, %17, %24, %10;\n\t madc.hi.cc.u32 %9, %17, %24, %11;\n\t"
        "madc.lo.cc.u32 %10, %19, %24, %12;\n\t madc.hi.cc.u32 %11, %19, %24, %13;\n\t"
        "madc.lo.cc.u32 %12, %21, %24, %14;\n\t madc.hi.cc.u32 %13, %21, %24, %15;\n\t"
        "madc.lo.cc.u32 %14, %23, %24, 0;\n\t madc.hi.u32 %15, %23, %24, 0;\n\t"
        "mad.lo.cc.u32 %0, %16, %24, %0;\n\t madc.hi.cc.u32 %1, %16, %24, %1;\n\t"
        "madc.lo.cc.u32 %2, %18, %24, %2;\n\t madc.hi.cc.u32 %3, %18, %24, %3;\n\t"
        "madc.lo.cc.u32 %4, %20, %24, %4;\n\t madc.hi.cc.u32 %5, %20, %24, %5;\n\t"
        "madc.lo.cc.u32 %6, %22, %24, %6;\n\t madc.hi.cc.u32 %7, %22, %24, %7;\n\t"
        "addc.u32 %15, %15, 0;"
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),
          "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b));
}

// Reduction step: m = E[0] * inv; O += p[odd]*m; E += p[even]*m (E[0] becomes 0);
// E's carry-out goes to O[7].
template <class P>
__device__ __forceinline__ void mm_redc(uint32_t (&E)[8], uint32_t (&O)[8]) {
    uint32_t m = E[0] * P::inv;
    asm("mad.lo.cc.u32 %8, %16, %18, %8;\n\t madc.hi.cc.u32 %9, %16, %18, %9;\n\t"
        "madc.lo.cc.u32 %10, %16, %20, %10;\n\t madc.hi.cc.u32 %11, %16, %20, %11;\n\t"
        "madc.lo.cc.u32 %12, %16, %22, %12;\n\t madc.hi.cc.u32 %13, %16, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %16, %24, %14;\n\t madc.hi.u32 %15, %16, %24, %15;\n\t"
        "mad.lo.cc.u32 %0, %16, %17, %0;\n\t madc.hi.cc.u32 %1, %16, %17, %1;\n\t"
        "madc.lo.cc.u32 %2, %16, %19, %2;\n\t madc.hi.cc.u32 %3, %16, %19, %3;\n\t"
        "madc.lo.cc.u32 %4, %16, %21, %4;\n\t madc.hi.cc.u32 %5, %16, %21, %5;\n\t"
        "madc.lo.cc.u32 %6, %16, %23, %6;\n\t madc.hi.cc.u32 %7, %16, %23, %7;\n\t"
        "addc.u32 %15, %15, 0;"
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),
          "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7])
        : "r"(m), "n"(P::p0), "n"(P::p1), "n"(P::p2), "n"(P::p3), "n"(P::p4), "n"(P::p5), "n"(P::p6),
          "n"(P::p7));
}

// r = a*b*2^-256 mod p, fully reduced.  Inputs < p.  (Interleaved product / reduction rows: 128 + 8 MACs.)
template <class P>
__device__ __forceinline__ Fp<P> fp_mul_cios(const Fp<P>& a, const Fp<P>& b) {
    uint32_t A[8], B[8];
    mm_row0(A, B, a.v, b.v[0]); mm_redc<P>(A, B);
    mm_row(B, A, a.v, b.v[1]);  mm_redc<P>(B, A);
    mm_row(A, B, a.v, b.v[2]);  mm_redc<P>(A, B);
    mm_row(B, A, a.v, b.v[3]);  mm_redc<P>(B, A);
    mm_row(A, B, a.v, b.v[4]);  mm_redc<P>(A, B);
    mm_row(B, A, a.v, b.v[5]);  mm_redc<P>(B, A);
    mm_row(A, B, a.v, b.v[6]);  mm_redc<P>(A, B);
    mm_row(B, A, a.v, b.v[7]);  mm_redc<P>(B, A);
    // last even accumulator is B (B[0] == 0, B[j] at position j-1); odd is A (A[j] at position j)
    Fp<P> r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]),
          "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]));
    fp_reduce_once<P>(r.v);
    return r;
}

// Second product of a fused pair: adds c[even]*d to E and c[odd]*d to O in place (same alignment as after
// mm_row0 / mm_row: E at positions 0..7, O at positions 1..8); E's carry-out goes to O[7].
__device__ __forceinline__ void mm_row_acc(uint32_t (&E)[8], uint32_t (&O)[8], const uint32_t (&c)[8], uint32_t d) {
    asm("mad.lo.cc.u32 %8, %17, %24, %8;\n\t madc.hi.cc.u32 %9, %17, %24, %9;\n\t"
        "madc.lo.cc.u32 %10, %19, %24, %10;\n\t madc.hi.cc.u32 %11, %19, %24, %11;\n\t"
        "madc.lo.cc.u32 %12, %21, %24, %12;\n\t madc.hi.cc.u32 %13, %21, %24, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %24, %14;\n\t madc.hi.u32 %15, %23, %24, %15;\n\t"
        "mad.lo.cc.u32 %0, %16, %24, %0;\n\t madc.hi.cc.u32 %1, %16, %24, %1;\n\t"
        "madc.lo.cc.u32 %2, %18, %24, %2;\n\t madc.hi.cc.u32 %3, %18, %24, %3;\n\t"
        "madc.lo.cc.u32 %4, %20, %24, %4;\n\t madc.hi.cc.u32 %5, %20, %24, %5;\n\t"
        "madc.lo.cc.u32 %6, %22, %24, %6;\n\t madc.hi.cc.u32 %7, %22, %24, %7;\n\t"
        "addc.u32 %15, %15, 0;"
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),
          "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7])
        : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]), "r"(d));
}

// r = (a*b + c*d) * 2^-256 mod p with ONE Montgomery reduction (one m*p row per 32 bits for both products):
// 192 + 8 multiply-accumulates instead of 256 + 16.  Inputs < p; the accumulators stay below 3p*2^32 and the
// result below p*(1 + 2p/2^256) < 1.38p, so one conditional subtraction finishes it.
template <class P>
__device__ __forceinline__ Fp<P> fp_mul2_add(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    uint32_t A[8], B[8];
    mm_row0(A, B, a.v, b.v[0]); mm_row_acc(A, B, c.v, d.v[0]); mm_redc<P>(A, B);
    mm_row(B, A, a.v, b.v[1]);  mm_row_acc(B, A, c.v, d.v[1]); mm_redc<P>(B, A);
    mm_row(A, B, a.v, b.v[2]);  mm_row_acc(A, B, c.v, d.v[2]); mm_redc<P>(A, B);
    mm_row(B, A, a.v, b.v[3]);  mm_row_acc(B, A, c.v, d.v[3]); mm_redc<P>(B, A);
    mm_row(A, B, a.v, b.v[4]);  mm_row_acc(A, B, c.v, d.v[4]); mm_redc<P>(A, B);
    mm_row(B, A, a.v, b.v[5]);  mm_row_acc(B, A, c.v, d.v[5]); mm_redc<P>(B, A);
    mm_row(A, B, a.v, b.v[6]);  mm_row_acc(A, B, c.v, d.v[6]); mm_redc<P>(A, B);
    mm_row(B, A, a.v, b.v[7]);  mm_row_acc(B, A, c.v, d.v[7]); mm_redc<P>(B, A);
    Fp<P> r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]),
          "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]));
    fp_reduce_once<P>(r.v);
    return r;
}
// r = a*b - c*d (same cost: c is negated first)
template <class P>
__device__ __forceinline__ Fp<P> fp_mul2_sub(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    return fp_mul2_add<P>(a, b, fp_neg<P>(c), d);
}

// The product the kernels use is the interleaved one above.  A separated-operand-scanning squaring (100 MACs) and a
// Karatsuba product (112 MACs) were generated and measured on B200 in round 1: 62.3 and 67.2 G products/s against 68.1
// for this one (ptxas turns the carry-sink additions into IMAD.X on the same pipe and the dependent addition chains
// grow from 182 to 295 / 284 instructions), so they were removed ("measured and rejected").
template <class P>
__device__ __forceinline__ Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) { return fp_mul_cios<P>(a, b); }
template <class P>
__device__ __forceinline__ Fp<P> fp_sqr(const Fp<P>& a) { return fp_mul<P>(a, a); }

// Montgomery form -> canonical integer (multiply by 1)
template <class P>
__device__ __forceinline__ Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> one = Fp<P>::zero();
    one.v[0] = 1;
    return fp_mul<P>(a, one);
}
template <class P>
__device__ __forceinline__ Fp<P> fp_to_mont(const Fp<P>& a) { return fp_mul<P>(a, Fp<P>::r2()); }

// a^(p-2) by square-and-multiply (setup / finalisation paths only)
template <class P>
__device__ __noinline__ Fp<P> fp_inv(const Fp<P>& a) {
    const uint32_t e[8] = {P::p0 - 2u, P::p1, P::p2, P::p3, P::p4, P::p5, P::p6, P::p7};
    Fp<P> acc = Fp<P>::one();
    for (int i = 255; i >= 0; i--) {
        acc = fp_sqr<P>(acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fp_mul<P>(acc, a);
    }
    return acc;
}

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

}  // namespace b2
