// chacha.cuh -- the ChaCha20 block function (RFC 8439, 2.3), plain 32-bit integer code that compiles for the device
// and for the host: the counter-based generator behind the vanishing argument's random polynomial
// (halo2_proofs/src/plonk/vanishing/prover.rs:48-63 draws 2n field elements and 2n indices from thread_rng, which is
// ChaCha keyed with 256 bits; here block j of a ChaCha20 stream keyed by 256 bits of the caller's rng is computed where
// coefficient j / 3 is made, so the polynomial never exists on the host).  host/chacha_selftest.cpp compiles this very
// file with g++ and is checked against RFC 8439's vector and an independent implementation on CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_CHACHA_FN __host__ __device__ __forceinline__
#else
#define B2_CHACHA_FN inline
#endif

namespace b2 {

B2_CHACHA_FN uint32_t chacha_rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

#define B2_CHACHA_QR(a, b, c, d)                                                                            \
    a += b; d ^= a; d = chacha_rotl(d, 16);                                                                 \
    c += d; b ^= c; b = chacha_rotl(b, 12);                                                                 \
    a += b; d ^= a; d = chacha_rotl(d, 8);                                                                  \
    c += d; b ^= c; b = chacha_rotl(b, 7);

// out = the 16 little-endian words of key stream block `counter` for `key` (8 little-endian words) and nonce words
// n0, n1, n2: state = "expand 32-byte k" | key | counter | nonce, 20 rounds, + state
B2_CHACHA_FN void chacha20_block(const uint32_t* key, uint32_t counter, uint32_t n0, uint32_t n1, uint32_t n2,
                                 uint32_t* out) {
    const uint32_t s0 = 0x61707865u, s1 = 0x3320646eu, s2 = 0x79622d32u, s3 = 0x6b206574u;
    uint32_t x0 = s0, x1 = s1, x2 = s2, x3 = s3, x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3], x8 = key[4],
             x9 = key[5], x10 = key[6], x11 = key[7], x12 = counter, x13 = n0, x14 = n1, x15 = n2;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < 10; i++) {
        B2_CHACHA_QR(x0, x4, x8, x12)
        B2_CHACHA_QR(x1, x5, x9, x13)
        B2_CHACHA_QR(x2, x6, x10, x14)
        B2_CHACHA_QR(x3, x7, x11, x15)
        B2_CHACHA_QR(x0, x5, x10, x15)
        B2_CHACHA_QR(x1, x6, x11, x12)
        B2_CHACHA_QR(x2, x7, x8, x13)
        B2_CHACHA_QR(x3, x4, x9, x14)
    }
    out[0] = x0 + s0;       out[1] = x1 + s1;       out[2] = x2 + s2;        out[3] = x3 + s3;
    out[4] = x4 + key[0];   out[5] = x5 + key[1];   out[6] = x6 + key[2];    out[7] = x7 + key[3];
    out[8] = x8 + key[4];   out[9] = x9 + key[5];   out[10] = x10 + key[6];  out[11] = x11 + key[7];
    out[12] = x12 + counter; out[13] = x13 + n0;    out[14] = x14 + n1;      out[15] = x15 + n2;
}

#undef B2_CHACHA_QR

}  // namespace b2
