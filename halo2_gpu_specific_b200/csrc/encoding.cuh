// encoding.cuh -- G1 point compression / decompression for Params::read / Params::write
// (halo2_proofs/src/poly/commitment.rs:241-294: k || g || g_lagrange as 32-byte compressed points ||
// additional data).  The reference decompresses 2 * 2^k points on the CPU with `parallelize`
// (:262-273, one square root in Fq per point); here the file bytes go to the device as they are and the
// SRS never exists on the host in affine form.
//
// Encoding (GroupEncoding of the pinned pairing crate, [EXT] in SURVEY 8c -- restated from the
// pasta / pairing_bn256 convention and parametrised where it could differ): x as 32 little-endian bytes
// of the canonical integer; bit `sign_bit` (7 = top bit of byte 31) carries the parity of the canonical
// y; the identity is the all-zero string.
#pragma once
#include "curve.cuh"

namespace b2 {

// a^((q + 1) / 4): square root in Fq (q = 3 mod 4) when a is a quadratic residue
__device__ __noinline__ Fq fq_sqrt_candidate(const Fq& a) {
    // (q + 1) / 4, little-endian 32-bit limbs
    const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u,
                           0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
    Fq acc = Fq::one();
    for (int i = 253; i >= 0; i--) {
        acc = fp_sqr<FqParams>(acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fp_mul<FqParams>(acc, a);
    }
    return acc;
}

// status: 0 ok; first failing index + 1 is written to *bad with atomicMin-style race (any failing index)
__global__ void __launch_bounds__(128) g1_decompress_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                            unsigned long long n, uint32_t sign_bit,
                                                            unsigned long long* __restrict__ bad) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq x = fp_load<FqParams>(in + 2ull * i);
    const uint32_t sign = (x.v[7] >> (24 + sign_bit)) & 1u;
    x.v[7] &= ~(1u << (24 + sign_bit));
    Fq y = Fq::zero();
    bool ok = true;
    if (x.is_zero() && !sign) {
        // identity
    } else {
        // canonical x must be < q
        uint32_t t[8];
#pragma unroll
        for (int l = 0; l < 8; l++) t[l] = x.v[l];
        fp_reduce_once<FqParams>(t);
        bool reduced = false;
#pragma unroll
        for (int l = 0; l < 8; l++) reduced |= (t[l] != x.v[l]);
        if (reduced) {
            ok = false;
        } else {
            x = fp_to_mont<FqParams>(x);
            const Fq three = FQ_ADD(FQ_DBL(Fq::one()), Fq::one());
            const Fq rhs = FQ_ADD(FQ_MUL(FQ_SQR(x), x), three);
            y = fq_sqrt_candidate(rhs);
            if (FQ_SQR(y) != rhs) {
                ok = false;
            } else {
                const Fq yc = fp_from_mont<FqParams>(y);
                if ((yc.v[0] & 1u) != sign) y = fp_neg<FqParams>(y);
            }
        }
    }
    if (!ok) {
        atomicMin(bad, i + 1ull);
        x = Fq::zero();
        y = Fq::zero();
    }
    fp_store<FqParams>(out + 4ull * i, x);
    fp_store<FqParams>(out + 4ull * i + 2, y);
}

__global__ void __launch_bounds__(128) g1_compress_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                          unsigned long long n, uint32_t sign_bit) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fq x = fp_load<FqParams>(in + 4ull * i), y = fp_load<FqParams>(in + 4ull * i + 2);
    Fq r = Fq::zero();
    if (!(x.is_zero() && y.is_zero())) {
        r = fp_from_mont<FqParams>(x);
        const Fq yc = fp_from_mont<FqParams>(y);
        r.v[7] |= (yc.v[0] & 1u) << (24 + sign_bit);
    }
    fp_store<FqParams>(out + 2ull * i, r);
}

}  // namespace b2
