// lookup.cuh -- logup multiplicities m(X) on device-resident columns.
//
// Replaces the sort + binary-search + count step of logup::Argument::compress
// (halo2_proofs/src/plonk/logup/prover.rs:117-179): the table's usable rows are sorted stably by value
// (par_sort_by_key, :123), every input value is located with <[T]>::binary_search_by_key (:146-148) and the table
// row that search returns takes the count.  When the table repeats a value, WHICH of the equal rows the search
// returns depends on its probe sequence (mid = left + size / 2, the first Equal probe wins: core::slice
// binary_search_by of the pinned nightly-2023-06-01 toolchain), so the kernel below runs exactly that loop on the
// stably sorted table -- the same row gets the count as in the reference.
//
// Pipeline (all on one stream):
//   lk_canon_kernel      table column: Montgomery -> canonical 256-bit keys (the order is the integer order)
//   lk_limb_or_kernel    per 64-bit limb: OR over i of (key_i ^ key_0)  -> which limbs / bytes differ at all
//   lk_hist / scan / lk_scatter   stable LSD radix sort of the row indices, 8 bits per pass, only over the bytes
//                        that differ somewhere (a 16-bit range table needs 2 passes, random field elements 8: the
//                        most significant differing limb decides unless two keys tie on it, in which case the
//                        lower limbs are sorted first)
//   lk_gather_kernel     sorted keys, contiguous (one 32-byte load per probe)
//   lk_search_kernel     one thread per input value: Rust's binary search, atomicAdd on the found row
//   lk_finish_kernel     counts -> Montgomery-form m column (zeros in the blinding rows), largest count
#pragma once
#include "fp.cuh"

namespace b2 {

constexpr int LK_THREADS = 256;
constexpr int LK_ITEMS = 16;                       // keys per thread per radix pass
constexpr int LK_TILE = LK_THREADS * LK_ITEMS;     // keys per block

struct LkKey {
    unsigned long long l[4];                       // canonical value, little-endian 64-bit limbs
};

__device__ __forceinline__ LkKey lk_key_from_fr(const Fr& c) {
    LkKey k;
#pragma unroll
    for (int i = 0; i < 4; i++) k.l[i] = (unsigned long long)c.v[2 * i] | ((unsigned long long)c.v[2 * i + 1] << 32);
    return k;
}
__device__ __forceinline__ LkKey lk_load_key(const uint4* keys, size_t i) {
    const uint4 a = keys[2 * i], b = keys[2 * i + 1];
    LkKey k;
    k.l[0] = (unsigned long long)a.x | ((unsigned long long)a.y << 32);
    k.l[1] = (unsigned long long)a.z | ((unsigned long long)a.w << 32);
    k.l[2] = (unsigned long long)b.x | ((unsigned long long)b.y << 32);
    k.l[3] = (unsigned long long)b.z | ((unsigned long long)b.w << 32);
    return k;
}
// -1 / 0 / +1 for a < b / a == b / a > b as 256-bit integers
__device__ __forceinline__ int lk_cmp(const LkKey& a, const LkKey& b) {
#pragma unroll
    for (int i = 3; i >= 0; i--) {
        if (a.l[i] < b.l[i]) return -1;
        if (a.l[i] > b.l[i]) return 1;
    }
    return 0;
}

// keys[i] = canonical form of col[i], i < count
__global__ void __launch_bounds__(LK_THREADS) lk_canon_kernel(const uint4* __restrict__ col, uint4* __restrict__ keys,
                                                              unsigned long long count) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Fr c = fp_from_mont<FrParams>(fp_load<FrParams>(col + 2 * i));
    fp_store<FrParams>(keys + 2 * i, c);
}

// limb_or[l] |= keys[i].l[l] ^ keys[0].l[l]
__global__ void __launch_bounds__(LK_THREADS) lk_limb_or_kernel(const uint4* __restrict__ keys, unsigned long long count,
                                                                unsigned long long* __restrict__ limb_or) {
    const LkKey first = lk_load_key(keys, 0);
    unsigned long long acc[4] = {0, 0, 0, 0};
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const LkKey k = lk_load_key(keys, i);
#pragma unroll
        for (int l = 0; l < 4; l++) acc[l] |= k.l[l] ^ first.l[l];
    }
#pragma unroll
    for (int l = 0; l < 4; l++) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc[l] |= __shfl_xor_sync(0xffffffffu, acc[l], d);
        if ((threadIdx.x & 31) == 0 && acc[l]) atomicOr(limb_or + l, acc[l]);
    }
}

__global__ void lk_iota_kernel(uint32_t* __restrict__ idx, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) idx[i] = i;
}

// digit of sorted position p: byte `byte` of limb `limb` of the key of row idx[p]
__device__ __forceinline__ uint32_t lk_digit(const uint4* __restrict__ keys, uint32_t row, uint32_t limb, uint32_t byte) {
    const unsigned long long* k = reinterpret_cast<const unsigned long long*>(keys) + 4ull * row + limb;
    return (uint32_t)(__ldg(k) >> (8 * byte)) & 0xffu;
}

// Every warp owns LK_TILE / 8 consecutive positions of the block's tile (so that the order warp 0 .. 7, then position,
// is the input order: the scatter below is stable).  hist[bin * nblocks + block] = count of the bin in the block's tile.
__global__ void __launch_bounds__(LK_THREADS)
lk_hist_kernel(const uint4* __restrict__ keys, const uint32_t* __restrict__ idx_in, uint32_t count, uint32_t limb,
               uint32_t byte, uint32_t* __restrict__ hist, uint32_t nblocks) {
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * LK_TILE;
    for (int it = 0; it < LK_ITEMS; it++) {
        const uint32_t p = base + it * LK_THREADS + threadIdx.x;
        if (p < count) atomicAdd(&cnt[lk_digit(keys, idx_in[p], limb, byte)], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// offsets = exclusive scan of hist (bin-major, block-minor): the first output position of (bin, block)
__global__ void __launch_bounds__(LK_THREADS)
lk_scatter_kernel(const uint4* __restrict__ keys, const uint32_t* __restrict__ idx_in, uint32_t* __restrict__ idx_out,
                  uint32_t count, uint32_t limb, uint32_t byte, const uint32_t* __restrict__ offsets, uint32_t nblocks) {
    constexpr int WARPS = LK_THREADS / 32;
    constexpr int PER_WARP = LK_TILE / WARPS;
    __shared__ uint32_t cur[WARPS][256];          // per warp: next output position of every bin
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b = lane; b < 256; b += 32) cur[w][b] = 0;
    __syncwarp();
    const uint32_t wbase = blockIdx.x * LK_TILE + w * PER_WARP;
    // per-warp histogram of its segment
    for (int it = 0; it < PER_WARP / 32; it++) {
        const uint32_t p = wbase + it * 32 + lane;
        if (p < count) atomicAdd(&cur[w][lk_digit(keys, idx_in[p], limb, byte)], 1u);
    }
    __syncthreads();
    // bin b (thread b): global offset of the block, then the warps in order
    {
        const uint32_t b = threadIdx.x;
        uint32_t run = offsets[(size_t)b * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < WARPS; ww++) {
            const uint32_t c = cur[ww][b];
            cur[ww][b] = run;
            run += c;
        }
    }
    __syncthreads();
    // place: inside a group of 32 the lanes with the same digit keep their lane order
    for (int it = 0; it < PER_WARP / 32; it++) {
        const uint32_t p = wbase + it * 32 + lane;
        const bool live = p < count;
        uint32_t row = 0, d = 0xffffffffu;
        if (live) {
            row = idx_in[p];
            d = lk_digit(keys, row, limb, byte);
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (live) pos = cur[w][d] + rank;
        __syncwarp();
        if (live && rank == 0) cur[w][d] += __popc(peers);
        __syncwarp();
        if (live) idx_out[pos] = row;
    }
}

// skeys[p] = keys[idx[p]]
__global__ void __launch_bounds__(LK_THREADS)
lk_gather_kernel(const uint4* __restrict__ keys, const uint32_t* __restrict__ idx, uint4* __restrict__ skeys,
                 uint32_t count) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    const size_t r = idx[p];
    skeys[2 * (size_t)p] = keys[2 * r];
    skeys[2 * (size_t)p + 1] = keys[2 * r + 1];
}

// adjacent sorted keys that agree on limb `limb` but are not equal: the sort by that limb alone is not enough
__global__ void __launch_bounds__(LK_THREADS)
lk_tie_kernel(const uint4* __restrict__ skeys, uint32_t count, uint32_t limb, int* __restrict__ flag) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0 || p >= count) return;
    const LkKey a = lk_load_key(skeys, p - 1), b = lk_load_key(skeys, p);
    if (a.l[limb] == b.l[limb] && lk_cmp(a, b) != 0) atomicExch(flag, 1);
}

// One thread per input value (columns of n rows, rows < usable count): binary_search_by_key on the sorted table,
// counts[row found] += 1; *miss is raised when a value is not in the table
// ("logup binary_search_by_key should hit", logup/prover.rs:148).
__global__ void __launch_bounds__(LK_THREADS)
lk_search_kernel(const uint4* __restrict__ inputs, unsigned long long n, uint32_t usable, uint32_t n_inputs,
                 const uint4* __restrict__ skeys, const uint32_t* __restrict__ idx, uint32_t* __restrict__ counts,
                 int* __restrict__ miss) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned long long)usable * n_inputs) return;
    const unsigned long long col = t / usable, row = t % usable;
    const LkKey v = lk_key_from_fr(fp_from_mont<FrParams>(fp_load<FrParams>(inputs + 2 * (col * n + row))));
    uint32_t size = usable, left = 0, right = usable;
    while (left < right) {
        const uint32_t mid = left + size / 2;
        const int c = lk_cmp(lk_load_key(skeys, mid), v);        // element.cmp(target)
        if (c == 0) {
            atomicAdd(&counts[idx[mid]], 1u);
            return;
        }
        if (c < 0) left = mid + 1; else right = mid;
        size = right - left;
    }
    atomicExch(miss, 1);
}

// m[i] = counts[i] in Montgomery form for i < usable, 0 above; *largest = max count
__global__ void __launch_bounds__(LK_THREADS)
lk_finish_kernel(const uint32_t* __restrict__ counts, uint32_t usable, unsigned long long n, uint4* __restrict__ m,
                 uint32_t* __restrict__ largest) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (i < n) {
        Fr x = Fr::zero();
        if (i < usable) {
            c = counts[i];
            if (c) {
                x.v[0] = c;
                x = fp_to_mont<FrParams>(x);
            }
        }
        fp_store<FrParams>(m + 2 * i, x);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, d));
    if ((threadIdx.x & 31) == 0 && c) atomicMax(largest, c);
}

}  // namespace b2
