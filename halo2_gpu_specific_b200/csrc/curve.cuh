// curve.cuh -- BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic for the MSM kernels.
//
// Buckets are kept in extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2), which makes the mixed add of an affine SRS point 8M + 2S.  All
// operators are COMPLETE: the reference's CPU path uses the complete `Curve`
// operators inside buckets (halo2_proofs/src/arithmetic.rs:66-75), so P+P, P+(-P)
// and identity operands must give the group-law answer here too.
//
// Memory layouts (C ABI, include/b2pcs.h): affine = x||y (64 B, identity (0,0));
// Jacobian = X||Y||Z (96 B, identity Z = 0).
#pragma once
#include "fp.cuh"

namespace b2 {

struct Affine {
    Fq x, y;
    __device__ __forceinline__ bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct XYZZ {
    Fq x, y, zz, zzz;
    __device__ __forceinline__ bool is_identity() const { return zz.is_zero(); }
    __device__ __forceinline__ static XYZZ identity() {
        XYZZ r;
        r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
        return r;
    }
    __device__ __forceinline__ static XYZZ from_affine(const Affine& p) {
        XYZZ r;
        if (p.is_identity()) return identity();
        r.x = p.x; r.y = p.y; r.zz = Fq::one(); r.zzz = Fq::one();
        return r;
    }
};

#define FQ_MUL(a, b) fp_mul<FqParams>((a), (b))
#define FQ_SQR(a) fp_sqr<FqParams>((a))
#define FQ_ADD(a, b) fp_add<FqParams>((a), (b))
#define FQ_SUB(a, b) fp_sub<FqParams>((a), (b))
#define FQ_DBL(a) fp_dbl<FqParams>((a))

__device__ __forceinline__ Affine affine_load(const void* p) {
    Affine a;
    a.x = fp_load_nc<FqParams>(p);
    a.y = fp_load_nc<FqParams>(reinterpret_cast<const char*>(p) + 32);
    return a;
}
__device__ __forceinline__ XYZZ xyzz_load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    XYZZ r;
    r.x = fp_load<FqParams>(c); r.y = fp_load<FqParams>(c + 32);
    r.zz = fp_load<FqParams>(c + 64); r.zzz = fp_load<FqParams>(c + 96);
    return r;
}
__device__ __forceinline__ void xyzz_store(void* p, const XYZZ& a) {
    char* c = reinterpret_cast<char*>(p);
    fp_store<FqParams>(c, a.x); fp_store<FqParams>(c + 32, a.y);
    fp_store<FqParams>(c + 64, a.zz); fp_store<FqParams>(c + 96, a.zzz);
}

// doubling of an affine point (mdbl-2008-s-1, a = 0); p must not be the identity
__device__ __forceinline__ XYZZ xyzz_dbl_affine(const Affine& p) {
    XYZZ r;
    Fq u = FQ_DBL(p.y);
    Fq v = FQ_SQR(u);
    Fq w = FQ_MUL(u, v);
    Fq s = FQ_MUL(p.x, v);
    Fq xx = FQ_SQR(p.x);
    Fq m = FQ_ADD(FQ_DBL(xx), xx);
    r.x = FQ_SUB(FQ_SQR(m), FQ_DBL(s));
    r.y = fp_mul2_sub<FqParams>(m, FQ_SUB(s, r.x), w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
}

// dbl-2008-s-1, a = 0
__device__ __forceinline__ XYZZ xyzz_dbl(const XYZZ& p) {
    if (p.is_identity()) return p;
    XYZZ r;
    Fq u = FQ_DBL(p.y);
    Fq v = FQ_SQR(u);
    Fq w = FQ_MUL(u, v);
    Fq s = FQ_MUL(p.x, v);
    Fq xx = FQ_SQR(p.x);
    Fq m = FQ_ADD(FQ_DBL(xx), xx);
    r.x = FQ_SUB(FQ_SQR(m), FQ_DBL(s));
    r.y = fp_mul2_sub<FqParams>(m, FQ_SUB(s, r.x), w, p.y);
    r.zz = FQ_MUL(v, p.zz);
    r.zzz = FQ_MUL(w, p.zzz);
    return r;
}

// acc += p (mixed add, madd-2008-s), complete.  8M + 2S on the common path; the two products of Y3 share one
// Montgomery reduction (fp_mul2_sub), so the cost is 9.44 product-equivalents rather than 10.
__device__ __forceinline__ void xyzz_madd(XYZZ& acc, const Affine& p) {
    if (p.is_identity()) return;
    if (acc.is_identity()) {
        acc.x = p.x; acc.y = p.y; acc.zz = Fq::one(); acc.zzz = Fq::one();
        return;
    }
    Fq u2 = FQ_MUL(p.x, acc.zz);
    Fq s2 = FQ_MUL(p.y, acc.zzz);
    Fq pp_ = FQ_SUB(u2, acc.x);
    Fq r = FQ_SUB(s2, acc.y);
    if (pp_.is_zero()) {
        if (r.is_zero()) acc = xyzz_dbl_affine(p);   // same point: double
        else acc = XYZZ::identity();                 // inverse points
        return;
    }
    Fq pp = FQ_SQR(pp_);
    Fq ppp = FQ_MUL(pp_, pp);
    Fq q = FQ_MUL(acc.x, pp);
    Fq x3 = FQ_SUB(FQ_SUB(FQ_SQR(r), ppp), FQ_DBL(q));
    Fq y3 = fp_mul2_sub<FqParams>(r, FQ_SUB(q, x3), acc.y, ppp);   // one reduction for both products
    acc.x = x3;
    acc.y = y3;
    acc.zz = FQ_MUL(acc.zz, pp);
    acc.zzz = FQ_MUL(acc.zzz, ppp);
}

// acc += b (add-2008-s), complete.  12M + 2S.
__device__ __forceinline__ void xyzz_add(XYZZ& acc, const XYZZ& b) {
    if (b.is_identity()) return;
    if (acc.is_identity()) { acc = b; return; }
    Fq u1 = FQ_MUL(acc.x, b.zz);
    Fq u2 = FQ_MUL(b.x, acc.zz);
    Fq s1 = FQ_MUL(acc.y, b.zzz);
    Fq s2 = FQ_MUL(b.y, acc.zzz);
    Fq pp_ = FQ_SUB(u2, u1);
    Fq r = FQ_SUB(s2, s1);
    if (pp_.is_zero()) {
        if (r.is_zero()) acc = xyzz_dbl(acc);
        else acc = XYZZ::identity();
        return;
    }
    Fq pp = FQ_SQR(pp_);
    Fq ppp = FQ_MUL(pp_, pp);
    Fq q = FQ_MUL(u1, pp);
    Fq x3 = FQ_SUB(FQ_SUB(FQ_SQR(r), ppp), FQ_DBL(q));
    Fq y3 = fp_mul2_sub<FqParams>(r, FQ_SUB(q, x3), s1, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = FQ_MUL(FQ_MUL(acc.zz, b.zz), pp);
    acc.zzz = FQ_MUL(FQ_MUL(acc.zzz, b.zzz), ppp);
}

// Non-inlined variants for cold / low-parallelism paths (keeps code size down).
__device__ __noinline__ void xyzz_add_ni(XYZZ& acc, const XYZZ& b) { xyzz_add(acc, b); }
__device__ __noinline__ void xyzz_dbl_ni(XYZZ& acc) { acc = xyzz_dbl(acc); }

// acc = [k] acc for a small non-negative integer k (double-and-add, MSB first)
__device__ __forceinline__ void xyzz_mul_small(XYZZ& acc, uint32_t k) {
    if (k == 0) { acc = XYZZ::identity(); return; }
    XYZZ base = acc;
    int top = 31 - __clz(k);
    for (int i = top - 1; i >= 0; i--) {
        xyzz_dbl_ni(acc);
        if ((k >> i) & 1u) xyzz_add_ni(acc, base);
    }
}

// XYZZ -> affine (one field inversion); identity -> (0, 0)
__device__ __forceinline__ Affine xyzz_to_affine(const XYZZ& p) {
    Affine a;
    if (p.is_identity()) { a.x = Fq::zero(); a.y = Fq::zero(); return a; }
    Fq inv = fp_inv<FqParams>(FQ_MUL(p.zz, p.zzz));   // 1 / (ZZ * ZZZ)
    Fq izz = FQ_MUL(inv, p.zzz);                      // 1 / ZZ
    Fq izzz = FQ_MUL(inv, p.zz);                      // 1 / ZZZ
    a.x = FQ_MUL(p.x, izz);
    a.y = FQ_MUL(p.y, izzz);
    return a;
}

// Jacobian (X, Y, Z) -> XYZZ:  ZZ = Z^2, ZZZ = Z^3
__device__ __forceinline__ XYZZ xyzz_from_jacobian(const Fq& X, const Fq& Y, const Fq& Z) {
    XYZZ r;
    if (Z.is_zero()) return XYZZ::identity();
    r.x = X; r.y = Y;
    r.zz = FQ_SQR(Z);
    r.zzz = FQ_MUL(r.zz, Z);
    return r;
}

}  // namespace b2
