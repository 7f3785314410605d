// halo2_b200_transcript.hpp -- C++ mirror of the prover's Fiat-Shamir transcript
// (halo2_proofs/src/transcript.rs:14-293: Blake2bWrite + Challenge255), host-only: hashing 65 bytes per commitment is
// host work in the reference too.  BLAKE2b is restated from RFC 7693 (the reference uses the blake2b_simd crate:
// 64-byte digest, personalisation "Halo2-Transcript", no key, no salt).  Points arrive as the engine returns them
// (G1 with z = 1 or G1Affine, Fq Montgomery limbs) and are written as the reference writes them: canonical
// little-endian coordinates into the hash state, the 32-byte compressed form into the proof.  The compressed form
// is GroupEncoding::to_bytes of the pinned pairing crate ([EXT]): x little-endian, parity of y in bit
// `sign_bit` of byte 31.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "halo2_b200.hpp"

namespace halo2_b200 {

// ---- BLAKE2b (RFC 7693), 64-byte digest, with the 16-byte personalisation of the parameter block ----------------
class Blake2b {
  public:
    explicit Blake2b(const char personal[16]) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        uint64_t param[8] = {0x0000000001010040ULL, 0, 0, 0, 0, 0, 0, 0};   // digest 64, key 0, fanout 1, depth 1
        std::memcpy(&param[6], personal, 16);
        for (int i = 0; i < 8; i++) h_[i] = IV[i] ^ param[i];
    }
    void update(const void* data, size_t len) {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        while (len) {
            if (fill_ == 128) {            // a full buffer is only compressed when more input follows (last-block flag)
                t_ += 128;
                compress(false);
                fill_ = 0;
            }
            size_t take = 128 - fill_ < len ? 128 - fill_ : len;
            std::memcpy(buf_ + fill_, p, take);
            fill_ += take; p += take; len -= take;
        }
    }
    void finalize(uint8_t out[64]) const {     // const: the transcript squeezes from a clone of its state
        Blake2b c = *this;
        c.t_ += c.fill_;
        std::memset(c.buf_ + c.fill_, 0, 128 - c.fill_);
        c.compress(true);
        std::memcpy(out, c.h_, 64);
    }

  private:
    uint64_t h_[8];
    uint8_t buf_[128] = {0};
    size_t fill_ = 0;
    uint64_t t_ = 0;                            // bytes compressed so far (inputs stay far below 2^64)
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(bool last) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                       0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        static const uint8_t SIGMA[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        std::memcpy(m, buf_, 128);
        for (int i = 0; i < 8; i++) { v[i] = h_[i]; v[i + 8] = IV[i]; }
        v[12] ^= t_;
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = SIGMA[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[i + 8];
    }
};

// ---- the little Fq arithmetic the transcript needs: Montgomery limbs -> canonical coordinates --------------------
namespace fq {
typedef unsigned __int128 u128;
static const uint64_t Q[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t INV = 0x87d20782e4866389ULL;
inline void from_mont(const uint64_t a[4], uint64_t out[4]) {       // a * 1 * R^-1 mod q
    uint64_t t[5] = {a[0], a[1], a[2], a[3], 0};
    for (int i = 0; i < 4; i++) {
        uint64_t m = t[0] * INV;
        u128 c = (u128)m * Q[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * Q[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = (uint64_t)(c >> 64);
    }
    bool ge = t[4] != 0;
    if (!ge) { ge = true; for (int i = 3; i >= 0; i--) { if (t[i] != Q[i]) { ge = t[i] > Q[i]; break; } } }
    if (ge) { u128 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - Q[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; } }
    std::memcpy(out, t, 32);
}
}  // namespace fq

namespace fr {
// canonical little-endian limbs of a Montgomery-form element (to_repr)
inline void to_repr(const Fr& a, uint64_t out[4]) {
    Fr one_raw = {{1, 0, 0, 0}};
    Fr c = mul(a, one_raw);
    std::memcpy(out, c.l, 32);
}
inline Fr add(const Fr& a, const Fr& b) {
    uint64_t t[5]; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; t[i] = (uint64_t)c; c >>= 64; }
    t[4] = (uint64_t)c;
    if (t[4] || geq_p(t)) sub_p(t);
    Fr r; std::memcpy(r.l, t, 32); return r;
}
// Fr::from_bytes_wide: a 512-bit little-endian integer mod r, in Montgomery form: lo * R + hi * 2^256 * R
inline Fr from_bytes_wide(const uint8_t bytes[64]) {
    static const Fr R3 = {{0x5e94d8e1b4bf0040ULL, 0x2a489cbe1cfbb6b8ULL, 0x893cc664a19fcfedULL, 0x0cf8594b7fcc657cULL}};   // 2^768 mod r
    Fr lo, hi;
    std::memcpy(lo.l, bytes, 32);
    std::memcpy(hi.l, bytes + 32, 32);
    return add(mul(lo, R2), mul(hi, R3));
}
}  // namespace fr

// ---- transcript.rs:150-226 ---------------------------------------------------------------------------------------
class Blake2bWrite {
  public:
    explicit Blake2bWrite(int sign_bit = 7) : state_("Halo2-Transcript"), sign_bit_(sign_bit) {}   // :161-164

    // :197-202 + Challenge255::new (:266-276)
    Fr squeeze_challenge() {
        const uint8_t prefix = 0;                                  // BLAKE2B_PREFIX_CHALLENGE, :14
        state_.update(&prefix, 1);
        uint8_t digest[64];
        state_.finalize(digest);
        return fr::from_bytes_wide(digest);
    }
    // :204-216; affine coordinates in Fq Montgomery limbs, (0, 0) = identity
    void common_point(const G1Affine& p) {
        uint64_t x[4], y[4];
        fq::from_mont(p.x, x);
        fq::from_mont(p.y, y);
        if (!(x[0] | x[1] | x[2] | x[3] | y[0] | y[1] | y[2] | y[3]))
            throw std::runtime_error("cannot write points at infinity to the transcript");
        const uint8_t prefix = 1;                                  // BLAKE2B_PREFIX_POINT, :17
        state_.update(&prefix, 1);
        state_.update(x, 32);
        state_.update(y, 32);
    }
    void common_point(const G1& p) { common_point(affine_of(p)); }
    // :218-223
    void common_scalar(const Fr& s) {
        uint64_t c[4];
        fr::to_repr(s, c);
        const uint8_t prefix = 2;                                  // BLAKE2B_PREFIX_SCALAR, :20
        state_.update(&prefix, 1);
        state_.update(c, 32);
    }
    // :180-184
    void write_point(const G1Affine& p) {
        common_point(p);
        uint64_t x[4], y[4];
        fq::from_mont(p.x, x);
        fq::from_mont(p.y, y);
        uint8_t out[32];
        std::memcpy(out, x, 32);
        out[31] |= (uint8_t)((y[0] & 1) << sign_bit_);
        writer_.insert(writer_.end(), out, out + 32);
    }
    void write_point(const G1& p) { write_point(affine_of(p)); }
    // :185-189
    void write_scalar(const Fr& s) {
        common_scalar(s);
        uint64_t c[4];
        fr::to_repr(s, c);
        const uint8_t* b = reinterpret_cast<const uint8_t*>(c);
        writer_.insert(writer_.end(), b, b + 32);
    }
    const std::vector<uint8_t>& finalize() const { return writer_; }   // :167-171

  private:
    Blake2b state_;
    int sign_bit_;
    std::vector<uint8_t> writer_;
    static G1Affine affine_of(const G1& p) {           // engine results are normalised: z = 1 (Montgomery) or 0
        G1Affine a;
        std::memset(&a, 0, sizeof a);
        if (p.z[0] | p.z[1] | p.z[2] | p.z[3]) { std::memcpy(a.x, p.x, 32); std::memcpy(a.y, p.y, 32); }
        return a;
    }
};

}  // namespace halo2_b200
