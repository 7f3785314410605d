// scan.cuh -- batch inversion and prefix product / prefix sum over Fr: the vector primitives behind the
// grand-product and grand-sum polynomials the prover builds right before commit_lagrange_and_ifft:
//   permutation z   halo2_proofs/src/plonk/permutation/prover.rs:72-165  (batch_invert :104, running product :149-152)
//   logup z         plonk/logup/prover.rs:263-336   (batch_invert of beta + f_i, running SUM :318-336)
//   shuffle z       plonk/shuffle/prover.rs:107-141 (batch_invert :132, running product :137-141)
// The reference runs them on the CPU (rayon `parallelize` + a serial scan); here they are device-resident so
// the z columns never visit the host between the expression kernel and the commitment.
#pragma once
#include "fp.cuh"
#include "fp_shoup.cuh"
#include "chacha.cuh"

namespace b2 {

// ---- batch inversion (ff::BatchInvert semantics: zeros are skipped and stay zero) ---------------------
// Thread t owns elements t, t + T, t + 2T, ... (coalesced), keeps the running product of its non-zero
// elements in `scratch` (same indexing), inverts its total once (Fermat), and walks back.
__global__ void __launch_bounds__(128) batch_invert_kernel(uint4* __restrict__ a, uint4* __restrict__ scratch,
                                                           unsigned long long n, unsigned long long T) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || t >= n) return;
    Fr run = Fr::one();
    unsigned long long i = t;
    for (; i < n; i += T) {
        fp_store<FrParams>(scratch + 2ull * i, run);          // product of the earlier non-zero elements
        const Fr x = fp_load<FrParams>(a + 2ull * i);
        if (!x.is_zero()) run = fp_mul<FrParams>(run, x);
    }
    Fr inv = fp_inv<FrParams>(run);
    // last owned index
    unsigned long long cnt = (n - t + T - 1) / T;
    for (unsigned long long j = cnt; j-- > 0;) {
        const unsigned long long idx = t + j * T;
        const Fr x = fp_load<FrParams>(a + 2ull * idx);
        if (x.is_zero()) continue;
        const Fr pre = fp_load<FrParams>(scratch + 2ull * idx);
        fp_store<FrParams>(a + 2ull * idx, fp_mul<FrParams>(inv, pre));
        inv = fp_mul<FrParams>(inv, x);
    }
}

// ---- prefix scan ------------------------------------------------------------------------------------
// out[0] = init, out[i + 1] = init (op) in[0] (op) ... (op) in[i], for i + 1 < n_out.   OP: 0 product, 1 sum.
constexpr int SCAN_FT = 256;   // threads per block
constexpr int SCAN_FK = 8;     // consecutive elements per thread
constexpr int SCAN_FTILE = SCAN_FT * SCAN_FK;

template <int OP>
__device__ __forceinline__ Fr scan_op(const Fr& x, const Fr& y) {
    return OP == 0 ? fp_mul<FrParams>(x, y) : fp_add<FrParams>(x, y);
}
template <int OP>
__device__ __forceinline__ Fr scan_identity() {
    return OP == 0 ? Fr::one() : Fr::zero();
}

// block-wide exclusive scan of one value per thread through shared memory (Hillis-Steele); *total = all
template <int OP>
__device__ __forceinline__ Fr scan_block_exclusive(const Fr& v, uint4* sm, Fr* total) {
    const uint32_t t = threadIdx.x;
    Fr incl = v;
    sm[2 * t] = make_uint4(incl.v[0], incl.v[1], incl.v[2], incl.v[3]);
    sm[2 * t + 1] = make_uint4(incl.v[4], incl.v[5], incl.v[6], incl.v[7]);
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
        Fr o = scan_identity<OP>();
        const bool has = t >= d;
        if (has) o = fp_load<FrParams>(sm + 2 * (t - d));
        __syncthreads();
        if (has) {
            incl = scan_op<OP>(o, incl);
            fp_store<FrParams>(sm + 2 * t, incl);
        }
        __syncthreads();
    }
    *total = fp_load<FrParams>(sm + 2 * (blockDim.x - 1));
    Fr excl = scan_identity<OP>();
    if (t > 0) excl = fp_load<FrParams>(sm + 2 * (t - 1));
    __syncthreads();
    return excl;
}

// pass 1: tile-local inclusive scan written to out[i + 1]; tile totals to tile_tot
template <int OP>
__global__ void __launch_bounds__(SCAN_FT) scan_tile_kernel(const uint4* __restrict__ in, unsigned long long n_in,
                                                           uint4* __restrict__ out, unsigned long long n_out,
                                                           uint4* __restrict__ tile_tot) {
    __shared__ uint4 sm[2 * SCAN_FT];
    const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_FTILE + (unsigned long long)threadIdx.x * SCAN_FK;
    Fr x[SCAN_FK];
    Fr run = scan_identity<OP>();
#pragma unroll
    for (int i = 0; i < SCAN_FK; i++) {
        const unsigned long long idx = base + i;
        x[i] = idx < n_in ? fp_load<FrParams>(in + 2ull * idx) : scan_identity<OP>();
        run = scan_op<OP>(run, x[i]);
        x[i] = run;
    }
    Fr total;
    const Fr excl = scan_block_exclusive<OP>(run, sm, &total);
#pragma unroll
    for (int i = 0; i < SCAN_FK; i++) {
        const unsigned long long idx = base + i;
        if (idx < n_in && idx + 1 < n_out) fp_store<FrParams>(out + 2ull * (idx + 1), scan_op<OP>(excl, x[i]));
    }
    if (threadIdx.x == 0) fp_store<FrParams>(tile_tot + 2ull * blockIdx.x, total);
}

// pass 2 (one block): tile_tot[j] <- init (op) tot[0] (op) ... (op) tot[j - 1]
template <int OP>
__global__ void __launch_bounds__(SCAN_FT) scan_top_kernel(uint4* __restrict__ tile_tot, uint32_t ntiles, const Fr init,
                                                          const uint4* __restrict__ d_init) {
    __shared__ uint4 sm[2 * SCAN_FT];
    const uint32_t per = (ntiles + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, ntiles);
    Fr run = scan_identity<OP>();
    for (uint32_t j = lo; j < hi; j++) run = scan_op<OP>(run, fp_load<FrParams>(tile_tot + 2ull * j));
    Fr total;
    Fr excl = scan_block_exclusive<OP>(run, sm, &total);
    const Fr start = d_init ? fp_load<FrParams>(d_init) : init;
    excl = scan_op<OP>(start, excl);
    for (uint32_t j = lo; j < hi; j++) {
        const Fr v = fp_load<FrParams>(tile_tot + 2ull * j);
        fp_store<FrParams>(tile_tot + 2ull * j, excl);
        excl = scan_op<OP>(excl, v);
    }
}

// pass 3: out[i + 1] <- prefix[tile(i)] (op) out[i + 1];  out[0] = init
template <int OP>
__global__ void __launch_bounds__(SCAN_FT) scan_apply_kernel(uint4* __restrict__ out, unsigned long long n_in,
                                                            unsigned long long n_out, const uint4* __restrict__ tile_tot) {
    const Fr pre = fp_load<FrParams>(tile_tot + 2ull * blockIdx.x);
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_out > 0) fp_store<FrParams>(out, pre);   // prefix of tile 0 = init
    const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_FTILE;
    for (uint32_t i = threadIdx.x; i < SCAN_FTILE; i += blockDim.x) {
        const unsigned long long idx = base + i;
        if (idx < n_in && idx + 1 < n_out) {
            const Fr v = fp_load<FrParams>(out + 2ull * (idx + 1));
            fp_store<FrParams>(out + 2ull * (idx + 1), scan_op<OP>(pre, v));
        }
    }
}

// out[i] = a[i] (op) b[i] on device pointers.  op: 0 mul, 1 add, 2 sub
__global__ void fr_vec_dev_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o,
                                  unsigned long long n, int op) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const Fr x = fp_load<FrParams>(a + 2ull * i), y = fp_load<FrParams>(b + 2ull * i);
        Fr r;
        if (op == 0) r = fp_mul<FrParams>(x, y);
        else if (op == 1) r = fp_add<FrParams>(x, y);
        else r = fp_sub<FrParams>(x, y);
        fp_store<FrParams>(o + 2ull * i, r);
    }
}

// ---- largest scalar of a column, in bits (find_max_scalar_bits, halo2_proofs/src/plonk/prover.rs:945-962) ------
// The reference folds the column with `max` and takes the bit length of the winner's canonical form: the bound it
// hands to commit_lagrange_with_bound.  Here: de-Montgomery, bit length, warp max, one atomicMax per warp.
__global__ void fr_max_bits_kernel(const uint4* __restrict__ a, unsigned long long n, unsigned* __restrict__ out) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned best = 0;
    for (; i < n; i += stride) {
        const Fr x = fp_from_mont<FrParams>(fp_load<FrParams>(a + 2ull * i));
        unsigned bits = 0;
#pragma unroll
        for (int l = 7; l >= 0; l--)
            if (bits == 0 && x.v[l]) bits = 32u * l + (32u - (unsigned)__clz(x.v[l]));
        best = max(best, bits);
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if ((threadIdx.x & 31u) == 0 && best) atomicMax(out, best);
}

// ---- the vanishing argument's random polynomial (halo2_proofs/src/plonk/vanishing/prover.rs:48-63) ----------
// coeff[i] = (a_i + random[u_i % k]) * (b_i + random[v_i % k]).  The reference draws a_i, u_i, b_i, v_i from
// thread_rng (ChaCha under a 256-bit key) inside a rayon loop; here they come from a counter-based generator of the
// same strength keyed by 256 bits that the caller's RNG supplies, so the polynomial is reproducible under a fixed RNG,
// unpredictable under a real one, and never exists on the host.
// The generator (restated in oracle/prover.py and halo2_gpu_specific_b200/plonk.py vanishing_streams): ChaCha20 key
// stream (RFC 8439 block function, nonce 0) under a 256-bit key from the caller's rng; coefficient i takes blocks
// 3i, 3i + 1, 3i + 2: a_i = (block 3i as a 512-bit little-endian integer) mod r, b_i = (block 3i + 1) mod r -- uniform
// over Fr up to 2^-250, like Fr::random -- and u_i, v_i = the first two little-endian 64-bit words of block 3i + 2.
struct VanishKey {
    uint32_t k[8];
};
// x mod r in Montgomery form for a 512-bit x = lo + hi * 2^256 given as 16 little-endian words:
// lo * R = mont(R^2, lo), hi * 2^256 * R = mont(R^3, hi).  The raw halves (any 256-bit value) are the SECOND operand:
// the interleaved product keeps its accumulator below (first operand) + r whatever the second one is.
__device__ __forceinline__ Fr vanish_wide_fr(const uint32_t (&w)[16]) {
    Fr lo, hi, r3;
#pragma unroll
    for (int l = 0; l < 8; l++) {
        lo.v[l] = w[l];
        hi.v[l] = w[8 + l];
    }
    r3.v[0] = 0xb4bf0040u; r3.v[1] = 0x5e94d8e1u; r3.v[2] = 0x1cfbb6b8u; r3.v[3] = 0x2a489cbeu;      // 2^768 mod r
    r3.v[4] = 0xa19fcfedu; r3.v[5] = 0x893cc664u; r3.v[6] = 0x7fcc657cu; r3.v[7] = 0x0cf8594bu;
    return fp_add<FrParams>(fp_mul<FrParams>(Fr::r2(), lo), fp_mul<FrParams>(r3, hi));
}
__global__ void vanishing_random_poly_kernel(uint4* __restrict__ out, const uint4* __restrict__ random, unsigned k,
                                             unsigned long long n, const VanishKey key) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint32_t w[16];
        chacha20_block(key.k, (uint32_t)(3ull * i), 0u, 0u, 0u, w);
        const Fr a = vanish_wide_fr(w);
        chacha20_block(key.k, (uint32_t)(3ull * i + 1ull), 0u, 0u, 0u, w);
        const Fr b = vanish_wide_fr(w);
        chacha20_block(key.k, (uint32_t)(3ull * i + 2ull), 0u, 0u, 0u, w);
        const unsigned long long u = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
        const unsigned long long v = (unsigned long long)w[2] | ((unsigned long long)w[3] << 32);
        const Fr ra = fp_load<FrParams>(random + 2ull * (u % k)), rb = fp_load<FrParams>(random + 2ull * (v % k));
        fp_store<FrParams>(out + 2ull * i, fp_mul<FrParams>(fp_add<FrParams>(a, ra), fp_add<FrParams>(b, rb)));
    }
}

}  // namespace b2

namespace b2 {

// ---- eval_polynomial (halo2_proofs/src/arithmetic.rs:707-735) ----------------------------------------
// sum_i a[i] * x^i for `columns` polynomials at one point.  Thread t of a column runs Horner over EVAL_C
// consecutive coefficients and scales by x^(t * EVAL_C) (two-level table built by the caller: lo[j] = x^(j * EVAL_C)
// for j < 2^EVAL_LO, hi[j] = x^(j * EVAL_C << EVAL_LO)); a block sums its threads and adds into out[column]
// with one carry-safe atomic-free step: per-block partials are written and a second launch folds them.
constexpr int EVAL_C = 32;
constexpr int EVAL_T = 128;
constexpr int EVAL_LO = 10;

__global__ void __launch_bounds__(EVAL_T) eval_poly_partial_kernel(const uint4* __restrict__ polys, unsigned long long stride,
                                                                  unsigned long long n, const Fr x, const Fr* __restrict__ lo,
                                                                  const Fr* __restrict__ hi, uint4* __restrict__ partials,
                                                                  uint32_t blocks_per_col,
                                                                  const uint4* const* __restrict__ ptrs = nullptr) {
    __shared__ uint4 sm[2 * EVAL_T];
    const uint32_t col = blockIdx.y;
    const unsigned long long t = (unsigned long long)blockIdx.x * EVAL_T + threadIdx.x;
    const unsigned long long i0 = t * EVAL_C;
    // polynomial `col`: column of a strided block, or (ptrs != nullptr) wherever the pointer table says
    const uint4* a = ptrs ? ptrs[col] : polys + 2ull * col * stride;
    Fr acc = Fr::zero();
    if (i0 < n) {
        const unsigned long long i1 = min(n, i0 + EVAL_C);
        // Horner two coefficients at a time: acc * x^2 + a[i + 1] * x under ONE Montgomery reduction (fp_mul2_add: 1.44
        // product-equivalents for two coefficients instead of 2), then + a[i]; the dependent chain is half as long too
        unsigned long long i = i1;
        if ((i1 - i0) & 1ull) {
            i--;
            acc = fp_load<FrParams>(a + 2ull * i);
        }
        const Fr x2 = fp_mul<FrParams>(x, x);
        while (i > i0) {
            i -= 2;
            const Fr hi_c = fp_load<FrParams>(a + 2ull * (i + 1)), lo_c = fp_load<FrParams>(a + 2ull * i);
            acc = fp_add<FrParams>(fp_mul2_add<FrParams>(acc, x2, hi_c, x), lo_c);
        }
        const Fr p = fp_mul<FrParams>(fp_load_nc<FrParams>(lo + (t & ((1u << EVAL_LO) - 1u))),
                                      fp_load_nc<FrParams>(hi + (t >> EVAL_LO)));
        acc = fp_mul<FrParams>(acc, p);
    }
    fp_store<FrParams>(sm + 2 * threadIdx.x, acc);
    __syncthreads();
    for (uint32_t d = EVAL_T / 2; d >= 1; d >>= 1) {
        if (threadIdx.x < d) {
            acc = fp_add<FrParams>(acc, fp_load<FrParams>(sm + 2 * (threadIdx.x + d)));
            fp_store<FrParams>(sm + 2 * threadIdx.x, acc);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) fp_store<FrParams>(partials + 2ull * ((unsigned long long)col * blocks_per_col + blockIdx.x), acc);
}

__global__ void __launch_bounds__(EVAL_T) eval_poly_final_kernel(const uint4* __restrict__ partials, uint32_t blocks_per_col,
                                                                uint4* __restrict__ out) {
    __shared__ uint4 sm[2 * EVAL_T];
    const uint32_t col = blockIdx.x;
    Fr acc = Fr::zero();
    for (uint32_t b = threadIdx.x; b < blocks_per_col; b += EVAL_T)
        acc = fp_add<FrParams>(acc, fp_load<FrParams>(partials + 2ull * ((unsigned long long)col * blocks_per_col + b)));
    fp_store<FrParams>(sm + 2 * threadIdx.x, acc);
    __syncthreads();
    for (uint32_t d = EVAL_T / 2; d >= 1; d >>= 1) {
        if (threadIdx.x < d) {
            acc = fp_add<FrParams>(acc, fp_load<FrParams>(sm + 2 * (threadIdx.x + d)));
            fp_store<FrParams>(sm + 2 * threadIdx.x, acc);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) fp_store<FrParams>(out + 2ull * col, acc);
}

// ---- kate_division (arithmetic.rs:752-773): q = (a(X) - a(b)) / (X - b), q[j] = sum_{i > j} a[i] b^(i-j-1) ------
// With P = prefix sums of a[i] * b^i (P[j+1] = sum_{i <= j}):  q[j] = (P[n] - P[j + 1]) * b^-(j + 1).
// step 1: t[i] = a[i] * bpow[i];   (scan)   step 3 below.
__global__ void kate_scale_kernel(const uint4* __restrict__ a, const Fr* __restrict__ bpow, uint4* __restrict__ t,
                                  unsigned long long n) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        fp_store<FrParams>(t + 2ull * i, fp_mul<FrParams>(fp_load<FrParams>(a + 2ull * i), fp_load_nc<FrParams>(bpow + i)));
}
__global__ void kate_finish_kernel(const uint4* __restrict__ P, const Fr* __restrict__ binvpow, uint4* __restrict__ q,
                                   unsigned long long n) {
    unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const Fr total = fp_load<FrParams>(P + 2ull * n);
    for (; j + 1 < n; j += stride) {
        const Fr s = fp_sub<FrParams>(total, fp_load<FrParams>(P + 2ull * (j + 1)));
        fp_store<FrParams>(q + 2ull * j, fp_mul<FrParams>(s, fp_load_nc<FrParams>(binvpow + j + 1)));
    }
}

// ---- multiopen batching (poly/multiopen/gwc/prover.rs:47-56; shplonk/prover.rs does the same fold) -----------------
// out[i] = sum_j v^(m-1-j) * polys[j][i]: the Horner fold `poly_batch = poly_batch * v + poly` over the m polynomials
// opened at one point, element by element.  v is the same for every element, so the product is the Shoup constant
// multiplication (w = v as a plain integer, wp = floor(v 2^256 / r)); one pass over the m polynomials, nothing
// written but the result.
__global__ void __launch_bounds__(256)
poly_combine_kernel(const uint4* const* __restrict__ polys, uint32_t m, unsigned long long n, const Fr v_plain,
                    const Fr v_comp, uint4* __restrict__ out) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        Fr acc = fp_load_nc<FrParams>(polys[0] + 2ull * i);
        for (uint32_t j = 1; j < m; j++) {
            const Fr pj = fp_load_nc<FrParams>(polys[j] + 2ull * i);
            acc = fp_add<FrParams>(fr_mul_shoup(acc, v_plain, v_comp), pj);
        }
        fp_store<FrParams>(out + 2ull * i, acc);
    }
}

}  // namespace b2
