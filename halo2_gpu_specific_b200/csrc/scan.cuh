// scan.cuh -- batch inversion and prefix product / prefix sum over Fr: the vector primitives behind the
// grand-product and grand-sum polynomials the prover builds right before commit_lagrange_and_ifft:
//   permutation z   halo2_proofs/src/plonk/permutation/prover.rs:72-165  (batch_invert :104, running product :149-152)
//   logup z         plonk/logup/prover.rs:263-336   (batch_invert of beta + f_i, running SUM :318-336)
//   shuffle z       plonk/shuffle/prover.rs:107-141 (batch_invert :132, running product :137-141)
// The reference runs them on the CPU (rayon `parallelize` + a serial scan); here they are device-resident so
// the z columns never visit the host between the expression kernel and the commitment.
#pragma once
#include "fp.cuh"

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

}  // namespace b2
