// probe_mixed.cuh -- two measurements that decide design questions of the MSM ("measured"):
//
//  1. mixed_probe_kernel: are the wide-integer multiplier (fmaheavy: IMAD.WIDE) and the fp64 FMA pipe of B200
//     independent?  Warps whose bit is set in `imad_mask` run carry-chained Montgomery products (or raw
//     IMAD.WIDE chains), warps in `dfma_mask` run DFMA chains, the rest exit.  Timing the three launches
//     (integer warps only, fp64 warps only, both) with the same per-warp iteration counts tells whether a
//     second, fp64-based field multiplication could run NEXT TO the integer one: T_both ~ max(T_i, T_d) if
//     the pipes and the issue slots allow it, T_i + T_d if they do not.
//
//  2. affine_batch_probe_kernel: the arithmetic of batched-affine bucket accumulation (Montgomery's trick),
//     with its real memory pattern: every thread adds B consecutive pairs of affine points; forward pass =
//     running product of the denominators x2 - x1 (prefix products parked in global memory, thread-minor so a
//     warp's accesses coalesce), ONE Fermat inversion per thread, backward pass = 6 products per addition
//     (denominator inverse, running inverse, lambda, lambda^2, lambda*(x1 - x3)) and the 64-byte result.  The
//     XYZZ mixed add of msm_accumulate_kernel costs 9.44 product-equivalents; this probe measures what the
//     6 + 380/B products actually buy once the two extra passes over the points are paid for.
#pragma once
#include "curve.cuh"

namespace b2 {

template <int ILP>
__global__ void __launch_bounds__(256) mixed_probe_kernel(uint4* sink, int iters_int, int iters_f64,
                                                          uint32_t imad_mask, uint32_t dfma_mask, int int_kind) {
    const uint32_t warp = threadIdx.x >> 5;
    if ((imad_mask >> warp) & 1u) {
        if (int_kind == 0) {   // Montgomery products: the instruction mix of the bignum kernels
            Fq x[ILP];
#pragma unroll
            for (int j = 0; j < ILP; j++) {
                x[j] = Fq::one();
                x[j].v[0] ^= threadIdx.x + j;
            }
            const Fq y = Fq::r2();
            for (int i = 0; i < iters_int; i++) {
#pragma unroll
                for (int j = 0; j < ILP; j++) x[j] = fp_mul_cios<FqParams>(x[j], y);
            }
            Fq acc = x[0];
#pragma unroll
            for (int j = 1; j < ILP; j++) acc = fp_add<FqParams>(acc, x[j]);
            if (acc.v[0] == 0x12345678u && acc.v[7] == 0x9abcdef0u) fp_store<FqParams>(sink, acc);
        } else {               // raw IMAD.WIDE chains (136 per "product" so that iteration counts compare)
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { lo[j] = threadIdx.x * 0x9E3779B9u + j; hi[j] = threadIdx.x + 17 * j + 1; }
            const uint32_t b = 0x9E3779B1u ^ iters_int;
            for (int i = 0; i < iters_int * ILP; i++) {
#pragma unroll
                for (int u = 0; u < 17; u++) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        asm volatile("{\n\t.reg .u32 x;\n\tmov.u32 x, %1;\n\tmad.lo.cc.u32 %0, x, %2, %0;\n\t"
                                     "madc.hi.u32 %1, x, %2, %1;\n\t}"
                                     : "+r"(lo[j]), "+r"(hi[j]) : "r"(b));
                    }
                }
            }
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) x += lo[j] ^ hi[j];
            if (x == 0x9abcdef0u) sink->x = x;
        }
        return;
    }
    if ((dfma_mask >> warp) & 1u) {
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = 1.0 + 1e-9 * (threadIdx.x + j);
        const double a = 1.0000001, b = 1e-7;
        for (int i = 0; i < iters_f64; i++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = fma(x[j], a, b);
            }
        }
        double acc = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) acc += x[j];
        if (acc == 123.456) *reinterpret_cast<double*>(sink) = acc;
    }
}

// pairs i in [t*B, (t+1)*B): out[i] = P[i] + Q[i] (affine, 64 B each).  prefix: nthreads*B field elements.
// Inputs of the probe are random points: the x2 == x1 cases (doubling / inverse pair) are detected and produce the
// identity (0, 0) here -- a product kernel would route them to the complete formulas -- so that the running product
// never becomes zero.
__global__ void __launch_bounds__(128, 4)
affine_batch_probe_kernel(const char* __restrict__ P, const char* __restrict__ Q, char* __restrict__ out,
                          uint4* __restrict__ prefix, uint32_t B, uint32_t nthreads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const size_t i0 = (size_t)t * B;
    Fq run = Fq::one();
#pragma unroll 1
    for (uint32_t j = 0; j < B; j++) {
        const Fq x1 = fp_load_nc<FqParams>(P + (i0 + j) * 64);
        const Fq x2 = fp_load_nc<FqParams>(Q + (i0 + j) * 64);
        Fq d = FQ_SUB(x2, x1);
        fp_store<FqParams>(prefix + 2 * ((size_t)j * nthreads + t), run);
        if (!d.is_zero()) run = FQ_MUL(run, d);
    }
    Fq inv = fp_inv<FqParams>(run);
#pragma unroll 1
    for (int j = (int)B - 1; j >= 0; j--) {
        const Affine p = affine_load(P + (i0 + j) * 64);
        const Affine q = affine_load(Q + (i0 + j) * 64);
        const Fq d = FQ_SUB(q.x, p.x);
        char* o = out + (i0 + j) * 64;
        if (d.is_zero()) {
            fp_store<FqParams>(o, Fq::zero());
            fp_store<FqParams>(o + 32, Fq::zero());
            continue;
        }
        const Fq pre = fp_load<FqParams>(prefix + 2 * ((size_t)j * nthreads + t));
        const Fq dinv = FQ_MUL(inv, pre);
        inv = FQ_MUL(inv, d);
        const Fq lam = FQ_MUL(FQ_SUB(q.y, p.y), dinv);
        const Fq x3 = FQ_SUB(FQ_SUB(FQ_SQR(lam), p.x), q.x);
        const Fq y3 = FQ_SUB(FQ_MUL(lam, FQ_SUB(p.x, x3)), p.y);
        fp_store<FqParams>(o, x3);
        fp_store<FqParams>(o + 32, y3);
    }
}

// the same additions with the XYZZ mixed add the MSM uses today (from_affine(P) + Q, no conversion back): the
// baseline the batch variant has to beat, on the same memory pattern.  out: 128-byte XYZZ per pair.
__global__ void __launch_bounds__(128, 4)
xyzz_pair_probe_kernel(const char* __restrict__ P, const char* __restrict__ Q, char* __restrict__ out, uint32_t B,
                       uint32_t nthreads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const size_t i0 = (size_t)t * B;
    XYZZ acc = XYZZ::identity();
#pragma unroll 1
    for (uint32_t j = 0; j < B; j++) {
        const Affine p = affine_load(P + (i0 + j) * 64);
        const Affine q = affine_load(Q + (i0 + j) * 64);
        xyzz_madd(acc, p);    // a running bucket, as in msm_accumulate_kernel: two mixed adds per pair
        xyzz_madd(acc, q);
    }
    xyzz_store(out + (size_t)t * 128, acc);
}

}  // namespace b2
