// msm.cuh -- BN254 G1 multi-scalar multiplication (signed-digit windowed Pippenger).
//
// Replaces multiexp_serial / best_multiexp / gpu_multiexp_single_gpu_with_bound
// (halo2_proofs/src/arithmetic.rs:20-108, 334-367, 465-492).  The result is the same
// group element the reference computes; it is returned normalised (Z = 1), and parity is
// defined on the affine point because projective representatives are not unique
// (the reference normalises before the transcript, plonk/prover.rs:130,304,484).
//
// Pipeline (all on one stream, no host round trips):
//   1. msm_digits_kernel   scalar: Montgomery -> canonical, + K (signed-digit bias), W digits
//                          of c bits; writes one 32-bit code per (window, scalar) and
//                          histograms bucket sizes (counting sort, pass 1)
//   2. msm_scan_kernel     exclusive scan of the W * 2^(c-1) bucket sizes
//   3. msm_scatter_kernel  counting sort, pass 2: point index | sign<<31 into bucket order
//   4. msm_accumulate_kernel  fixed-size chunks of the bucket-sorted entry list per thread
//                          (perfect load balance for any scalar distribution: zeros,
//                          16-bit advice values, all-equal scalars); mixed XYZZ adds;
//                          buckets wholly inside a chunk are written directly, the <= 2
//                          buckets cut by a chunk boundary become partial records
//   5. msm_partial_reduce_kernel  sums the partial records of each cut bucket, level by level
//   6. msm_reduce_kernel   per window: sum_j (j+1) * B_j with per-thread running sums,
//                          a block-level suffix scan (no scalar multiplications) and one
//                          small multiply per block
//   7. msm_final_kernel    sums block results per window, Horner over windows
//                          (c doublings each), normalises to Z = 1
#pragma once
#include "curve.cuh"
#include "fp_shoup.cuh"

namespace b2 {

struct MsmGeom {
    uint32_t c;              // window bits
    uint32_t W;              // number of windows
    uint32_t B;              // buckets per window = 2^(c-1)
    uint32_t n;              // number of scalars / points
    uint32_t bucket_stride;  // bucket id = w * bucket_stride + |d| - 1:  B (one bucket set per window)
                             // or 0 (precomputed 2^(c*w) multiples: all windows share one set)
    uint32_t point_stride;   // point index = w * point_stride + point_offset + i  (0 for plain bases)
    uint32_t point_offset;
};

// ---------------------------------------------------------------- 1. digits + histogram
// code[w * n + i] = (|d| << 1) | (d < 0), 0 when the digit is zero.
// Signed digits: add K = sum_{w < W-1} 2^(c*w + c - 1) once, then digit_w = window_w - 2^(c-1)
// for w < W-1 and the top window is taken unsigned (it has at most c-1 significant bits
// because W = floor(max_bits / c) + 1).
// canonical scalar + signed-digit bias K; returns the biased 256-bit value in s
__device__ __forceinline__ Fr msm_biased_scalar(const uint4* __restrict__ scalars, uint32_t i, const MsmGeom& g,
                                                uint32_t max_bits, int* __restrict__ err_flag) {
    Fr s = fp_from_mont<FrParams>(fp_load_nc<FrParams>(scalars + 2ull * i));
    // contract check: scalar < 2^max_bits
    if (err_flag != nullptr && max_bits < 256) {
        uint32_t over = 0;
#pragma unroll
        for (int l = 0; l < 8; l++) {
            const int lo_bit = l * 32;
            if ((int)max_bits <= lo_bit) over |= s.v[l];
            else if ((int)max_bits < lo_bit + 32) over |= s.v[l] >> (max_bits - lo_bit);
        }
        if (over) atomicExch(err_flag, 1);
    }
    uint32_t k[8];
#pragma unroll
    for (int l = 0; l < 8; l++) k[l] = 0;
    for (uint32_t w = 0; w + 1 < g.W; w++) {
        const uint32_t bit = g.c * w + g.c - 1;
        if (bit < 256) k[bit >> 5] |= 1u << (bit & 31);
    }
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(s.v[0]), "+r"(s.v[1]), "+r"(s.v[2]), "+r"(s.v[3]), "+r"(s.v[4]), "+r"(s.v[5]), "+r"(s.v[6]),
          "+r"(s.v[7])
        : "r"(k[0]), "r"(k[1]), "r"(k[2]), "r"(k[3]), "r"(k[4]), "r"(k[5]), "r"(k[6]), "r"(k[7]));
    return s;
}

// digit of window w: code = (|d| << 1) | (d < 0), 0 when the digit is zero
__device__ __forceinline__ uint32_t msm_digit_code(const Fr& s, uint32_t w, const MsmGeom& g) {
    const uint32_t half = 1u << (g.c - 1);
    const uint32_t mask = (1u << g.c) - 1u;
    const uint32_t bit = g.c * w;
    uint32_t raw = 0;
    if (bit < 256) {
        const uint32_t limb = bit >> 5, off = bit & 31;
        uint64_t two = s.v[limb];
        if (limb + 1 < 8) two |= (uint64_t)s.v[limb + 1] << 32;
        raw = (uint32_t)(two >> off) & mask;
    }
    const int32_t d = (w + 1 < g.W) ? (int32_t)raw - (int32_t)half : (int32_t)raw;
    if (d == 0) return 0;
    const uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    // the unsigned top window only exceeds the bucket range when the scalar breaks the max_bits
    // contract; that call fails with B2_ERR_BOUND, but it must not write out of bounds first
    if (mag > half) return 0;
    return (mag << 1) | (d < 0 ? 1u : 0u);
}

__global__ void msm_digits_kernel(const uint4* __restrict__ scalars, uint32_t* __restrict__ codes,
                                  uint32_t* __restrict__ counts, MsmGeom g, uint32_t max_bits,
                                  int* __restrict__ err_flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const Fr s = msm_biased_scalar(scalars, i, g, max_bits, err_flag);
    for (uint32_t w = 0; w + 1 < g.W; w++) {
        const uint32_t code = msm_digit_code(s, w, g);
        if (code) atomicAdd(&counts[(size_t)w * g.bucket_stride + ((code >> 1) - 1)], 1u);
        codes[(size_t)w * g.n + i] = code;
    }
    // The top window often has only a few significant bits (254 mod c, or the carry of a bounded
    // scalar), i.e. a handful of buckets that every scalar hits: aggregate equal buckets inside the
    // warp so that those hot counters see one atomic per warp instead of 32.
    {
        const uint32_t w = g.W - 1;
        const uint32_t code = msm_digit_code(s, w, g);
        codes[(size_t)w * g.n + i] = code;
        const uint32_t key = code ? (w * g.bucket_stride + ((code >> 1) - 1)) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(__activemask(), key);
        if (code && (threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&counts[key], (uint32_t)__popc(peers));
    }
}

// ---------------------------------------------------------------- 2. exclusive scan
// Three small kernels: per-tile scan (SCAN_TILE counts per block), scan of the tile totals,
// and the add-back that writes offsets[] (+ offsets[total] = number of entries) and the
// scatter cursors.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* sm, uint32_t* total) {
    // sm: SCAN_THREADS entries.  Hillis-Steele; returns the exclusive prefix of v.
    const uint32_t t = threadIdx.x;
    sm[t] = v;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
        uint32_t o = (t >= d) ? sm[t - d] : 0;
        __syncthreads();
        sm[t] += o;
        __syncthreads();
    }
    const uint32_t incl = sm[t];
    *total = sm[blockDim.x - 1];
    __syncthreads();
    return incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
msm_scan_tile_kernel(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                     uint32_t* __restrict__ tile_sums, uint32_t total) {
    __shared__ uint32_t sm[SCAN_THREADS];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) {
        v[i] = (base + i < total) ? counts[base + i] : 0;
        s += v[i];
    }
    uint32_t tot;
    uint32_t run = block_exclusive_scan(s, sm, &tot);
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) {
        if (base + i < total) offsets[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of ntiles tile totals (ntiles <= SCAN_THREADS * 64)
__global__ void __launch_bounds__(SCAN_THREADS)
msm_scan_top_kernel(uint32_t* __restrict__ tile_sums, uint32_t ntiles, uint32_t* __restrict__ grand_total) {
    __shared__ uint32_t sm[SCAN_THREADS];
    const uint32_t per = (ntiles + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, ntiles);
    uint32_t s = 0;
    for (uint32_t j = lo; j < hi; j++) s += tile_sums[j];
    uint32_t tot;
    uint32_t run = block_exclusive_scan(s, sm, &tot);
    for (uint32_t j = lo; j < hi; j++) {
        const uint32_t v = tile_sums[j];
        tile_sums[j] = run;
        run += v;
    }
    if (threadIdx.x == 0) *grand_total = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS)
msm_scan_add_kernel(uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor,
                    const uint32_t* __restrict__ tile_sums, uint32_t total) {
    const uint32_t add = tile_sums[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) {
        if (base + i < total) {
            const uint32_t v = offsets[base + i] + add;
            offsets[base + i] = v;
            cursor[base + i] = v;
        }
    }
}

// ---------------------------------------------------------------- 3. scatter
__global__ void msm_scatter_kernel(const uint32_t* __restrict__ codes, uint32_t* __restrict__ cursor,
                                   uint32_t* __restrict__ sorted, MsmGeom g) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = blockIdx.y;
    if (i >= g.n) return;
    const uint32_t code = codes[(size_t)w * g.n + i];
    const uint32_t entry = (w * g.point_stride + g.point_offset + i) | ((code & 1u) << 31);
    if (w + 1 < g.W) {
        if (code == 0) return;
        const uint32_t pos = atomicAdd(&cursor[(size_t)w * g.bucket_stride + ((code >> 1) - 1)], 1u);
        sorted[pos] = entry;
        return;
    }
    // top window: warp-aggregated (see msm_digits_kernel)
    const uint32_t key = code ? (w * g.bucket_stride + ((code >> 1) - 1)) : 0xffffffffu;
    const uint32_t peers = __match_any_sync(__activemask(), key);
    if (code == 0) return;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == (uint32_t)leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    sorted[base + __popc(peers & ((1u << lane) - 1u))] = entry;
}

// ---------------------------------------------------------------- 4. accumulate
__device__ __forceinline__ Affine msm_load_point(const char* __restrict__ bases, uint32_t stride, uint32_t ent) {
    const uint32_t idx = ent & 0x7fffffffu;
    Affine p = affine_load(bases + (size_t)idx * stride);
    if (ent >> 31) p.y = fp_neg<FqParams>(p.y);
    return p;
}

// Thread t owns entries [t*L, min((t+1)*L, E)) of the bucket-sorted list; offsets[] has nb+1
// entries.  The loop is FLAT over entries: every lane of a warp executes the expensive mixed
// add in lock step for exactly L iterations, and bucket boundaries are handled by a short
// divergent flush in between (a per-bucket inner loop would leave lanes idle until the
// largest bucket of the warp is done).
// part_pt[2t], part_pt[2t+1]: head / tail partial sums; part_bucket = bucket id or 0xffffffff.
#ifndef MSM_ACC_MIN_BLOCKS
#define MSM_ACC_MIN_BLOCKS 4
#endif
// The next point is requested two iterations ahead (both 32-byte sectors).  DRAM traffic of this kernel is 2.05x the
// gathered bytes (7.5 GB for 54.5 M points of 64 B) and that factor is the memory system's, not the prefetch's: it is the
// same with one prefetched sector, with none (7.28 GB, kernel 1.4 % slower) and under L2 fetch-granularity hints of
// 32 / 64 / 128 bytes (profiles/r2_ncu_summary.md 4); HBM is 13 % busy, so it costs nothing.
__device__ __forceinline__ void msm_prefetch_point(const char* __restrict__ bases, uint32_t stride, uint32_t ent) {
    const char* p = bases + (size_t)(ent & 0x7fffffffu) * stride;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32));
}

// Partial-record convention (shared with msm_partial_reduce_kernel): every producer thread owns
// two record slots.  slot 0 = sum of a bucket that BEGAN before the thread's range ("head cut"),
// slot 1 = sum of a bucket that CONTINUES after it ("tail cut"); a bucket covering the whole
// range writes its sum to slot 0 and an identity point with the same id to slot 1, so the
// records of one bucket are always adjacent in (thread, slot) order.  Unused slots carry the id
// 0xffffffff.
__device__ __forceinline__ void msm_flush_bucket(const XYZZ& acc, uint32_t b, bool head_cut, bool tail_cut,
                                                 uint32_t t, char* __restrict__ buckets,
                                                 char* __restrict__ part_pt, uint32_t* __restrict__ part_bucket) {
    if (!head_cut && !tail_cut) {
        xyzz_store(buckets + (size_t)b * 128, acc);
        return;
    }
    const uint32_t s0 = head_cut ? 0u : 1u;
    xyzz_store(part_pt + (size_t)(2 * t + s0) * 128, acc);
    part_bucket[2 * t + s0] = b;
    if (head_cut && tail_cut) {
        xyzz_store(part_pt + (size_t)(2 * t + 1) * 128, XYZZ::identity());
        part_bucket[2 * t + 1] = b;
    }
}

__global__ void __launch_bounds__(128, MSM_ACC_MIN_BLOCKS)
msm_accumulate_kernel(const char* __restrict__ bases, uint32_t base_stride, const uint32_t* __restrict__ sorted,
                      const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t chunk,
                      uint32_t nthreads_total, char* __restrict__ buckets, char* __restrict__ part_pt,
                      uint32_t* __restrict__ part_bucket) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads_total) return;
    const uint32_t E = offsets[nb];
    // equal shares of the REAL entry count (zero digits produce no entries), at least `chunk`
    chunk = max(chunk, (uint32_t)(((uint64_t)E + nthreads_total - 1) / nthreads_total));
    const uint64_t lo64 = (uint64_t)t * chunk;
    part_bucket[2 * t] = 0xffffffffu;
    part_bucket[2 * t + 1] = 0xffffffffu;
    if (lo64 >= E) return;
    const uint32_t lo = (uint32_t)lo64;
    const uint32_t hi = (uint32_t)min((uint64_t)E, lo64 + chunk);

    // bucket containing entry lo: offsets[b] <= lo < offsets[b + 1]
    uint32_t bl = 0, bh = nb;
    while (bh - bl > 1) {
        const uint32_t mid = (bl + bh) >> 1;
        if (offsets[mid] <= lo) bl = mid; else bh = mid;
    }
    uint32_t b = bl;
    uint32_t b_begin = offsets[b], b_end = offsets[b + 1];
    XYZZ acc = XYZZ::identity();
    uint32_t ent = sorted[lo];
    uint32_t ent_next = (lo + 1 < hi) ? sorted[lo + 1] : ent;
    msm_prefetch_point(bases, base_stride, ent_next);
#pragma unroll 1
    for (uint32_t e = lo; e < hi; e++) {
        if (e == b_end) {   // bucket b is complete: flush it and move to the next non-empty one
            msm_flush_bucket(acc, b, b_begin < lo, false, t, buckets, part_pt, part_bucket);
            acc = XYZZ::identity();
            do {
                b++;
                b_begin = b_end;
                b_end = offsets[b + 1];
            } while (b_end <= e);
        }
        Affine p = msm_load_point(bases, base_stride, ent);
        ent = ent_next;
        if (e + 2 < hi) {
            ent_next = sorted[e + 2];
            msm_prefetch_point(bases, base_stride, ent_next);
        }
        xyzz_madd(acc, p);
    }
    msm_flush_bucket(acc, b, b_begin < lo, b_end > hi, t, buckets, part_pt, part_bucket);
}

// ---------------------------------------------------------------- 5. fix-up of cut buckets
// The partial records form a list sorted by bucket id in which equal ids are adjacent.  Each
// thread sums runs of equal ids over PR_L consecutive records; complete runs go to buckets[],
// runs cut by the thread's range become the next level's records (same convention).  The host
// launches this with shrinking record counts until one thread sees everything, so a bucket cut
// into thousands of pieces (all-equal scalars, the 1-bit top window) costs O(log) launches,
// not one serial chain.
// Fast path first: one thread per record; a run of <= FIX_G records (the normal case is 2-4: the
// tail of one chunk and the head of the next, or a bucket spanning a few short chunks) is summed by its first record's thread.  Longer
// runs keep their ids in ids_out and raise *need_levels for the level kernels.
constexpr int FIX_G = 16;
__global__ void __launch_bounds__(128)
msm_fixup_small_kernel(const char* __restrict__ part_pt, const uint32_t* __restrict__ part_bucket, uint32_t nrec,
                       uint32_t* __restrict__ ids_out, int* __restrict__ need_levels, char* __restrict__ buckets) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    const uint32_t b = part_bucket[r];
    if (b == 0xffffffffu) { ids_out[r] = b; return; }
    // run boundaries within a +-FIX_G window
    uint32_t start = r, end = r + 1;
    while (start > 0 && r - start < FIX_G && part_bucket[start - 1] == b) start--;
    while (end < nrec && end - r < FIX_G && part_bucket[end] == b) end++;
    const bool open_left = start > 0 && part_bucket[start - 1] == b;
    const bool open_right = end < nrec && part_bucket[end] == b;
    if (open_left || open_right || end - start > FIX_G) {
        ids_out[r] = b;              // long run: leave it to the level kernels
        if (r == start || open_left) atomicExch(need_levels, 1);
        return;
    }
    ids_out[r] = 0xffffffffu;
    if (r != start) return;
    XYZZ acc = xyzz_load(part_pt + (size_t)r * 128);
    for (uint32_t q = r + 1; q < end; q++) {
        XYZZ o = xyzz_load(part_pt + (size_t)q * 128);
        xyzz_add_ni(acc, o);
    }
    xyzz_store(buckets + (size_t)b * 128, acc);
}

// Level kernel: one WARP per 32 consecutive records; a segmented shuffle reduction (5 rounds, every
// lane adding in parallel) sums each run of equal ids, so a level costs ~5 point additions of
// latency and shrinks the list 16x.  Run leaders flush with the same head/tail convention.
constexpr int PR_L = 32;
__device__ __forceinline__ Fq shfl_down_fq(const Fq& a, int d) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], d);
    return r;
}
__global__ void __launch_bounds__(128)
msm_partial_reduce_kernel(const char* __restrict__ in_pt, const uint32_t* __restrict__ in_bucket, uint32_t nrec,
                          char* __restrict__ out_pt, uint32_t* __restrict__ out_bucket, uint32_t nwarps,
                          const int* __restrict__ need_levels, char* __restrict__ buckets) {
    if (*need_levels == 0) return;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one output pair per warp
    const uint32_t lane = threadIdx.x & 31;
    if (wid >= nwarps) return;
    if (lane < 2) out_bucket[2 * wid + lane] = 0xffffffffu;
    const uint32_t r = wid * PR_L + lane;
    uint32_t id = (r < nrec) ? in_bucket[r] : 0xffffffffu;
    XYZZ acc = XYZZ::identity();
    if (id != 0xffffffffu) acc = xyzz_load(in_pt + (size_t)r * 128);
    // neighbours outside the warp decide whether the first / last run is cut
    const uint32_t lo = wid * PR_L, hi = min(nrec, lo + PR_L);
    const uint32_t prev_out = (lo > 0) ? in_bucket[lo - 1] : 0xffffffffu;
    const uint32_t next_out = (hi < nrec) ? in_bucket[hi] : 0xffffffffu;
    const uint32_t id_prev = __shfl_up_sync(0xffffffffu, id, 1);
    const bool leader = (id != 0xffffffffu) && (lane == 0 || id_prev != id);
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        XYZZ o;
        o.x = shfl_down_fq(acc.x, d);
        o.y = shfl_down_fq(acc.y, d);
        o.zz = shfl_down_fq(acc.zz, d);
        o.zzz = shfl_down_fq(acc.zzz, d);
        const uint32_t oid = __shfl_down_sync(0xffffffffu, id, d);
        if (lane + d < 32 && oid == id && id != 0xffffffffu) xyzz_add(acc, o);
    }
    // run extent: the run of a leader ends at the last lane with the same id
    const uint32_t same_as_last = __shfl_sync(0xffffffffu, id, 31);
    if (leader) {
        const bool head_cut = (lane == 0) && prev_out == id;
        const bool tail_cut = (same_as_last == id) && (hi - lo == 32) && next_out == id;
        msm_flush_bucket(acc, id, head_cut, tail_cut, wid, buckets, out_pt, out_bucket);
    }
}

// ---------------------------------------------------------------- 6. bucket reduction
// Block of RT threads handles RT*RM consecutive buckets of one window; bucket j (0-based
// inside the window) has weight j+1.  Output per block: sum_j (j+1) B_j over its range.
constexpr int MSM_RT = 128;      // threads per reduce block; each thread takes `rm` buckets (power of two)

__global__ void __launch_bounds__(MSM_RT)
msm_reduce_kernel(const char* __restrict__ buckets, const uint32_t* __restrict__ offsets, MsmGeom g,
                  uint32_t blocks_per_window, uint32_t rm, char* __restrict__ block_out) {
    extern __shared__ uint4 red_smem[];   // MSM_RT XYZZ points (128 B each)
    char* sm = reinterpret_cast<char*>(red_smem);
    const uint32_t w = blockIdx.x / blocks_per_window;
    const uint32_t blk = blockIdx.x % blocks_per_window;
    const uint32_t t = threadIdx.x;
    const uint32_t j0 = (blk * MSM_RT + t) * rm;   // first bucket of this thread (in-window)
    XYZZ run = XYZZ::identity(), sum = XYZZ::identity();
    for (int i = (int)rm - 1; i >= 0; i--) {
        const uint32_t j = j0 + (uint32_t)i;
        if (j < g.B) {
            const size_t gb = (size_t)w * g.B + j;
            if (offsets[gb + 1] > offsets[gb]) {
                XYZZ bk = xyzz_load(buckets + gb * 128);
                xyzz_add_ni(run, bk);
            }
        }
        xyzz_add_ni(sum, run);
    }
    // thread value: sum_i (i+1) B_{j0+i} = sum;  run = sum_i B_{j0+i}
    // block total = sum_t [ sum_t + (t*rm) * run_t ]  (+ blk offset handled below)
    // S_t = suffix sum of run over threads >= t  ->  sum_t t*run_t = sum_{t>=1} S_t
    xyzz_store(sm + (size_t)t * 128, run);
    __syncthreads();
    for (uint32_t d = 1; d < MSM_RT; d <<= 1) {
        XYZZ o = XYZZ::identity();
        const bool has = (t + d < MSM_RT);
        if (has) o = xyzz_load(sm + (size_t)(t + d) * 128);
        __syncthreads();
        if (has) {
            xyzz_add_ni(run, o);
            xyzz_store(sm + (size_t)t * 128, run);
        }
        __syncthreads();
    }
    // run == S_t now.  total_run = S_0.
    XYZZ total_run = xyzz_load(sm);
    __syncthreads();
    // v_t = sum_t + RM * S_t (t >= 1)
    XYZZ v = sum;
    if (t >= 1) {
        XYZZ s = run;
#pragma unroll 1
        for (uint32_t i = 1; i < rm; i <<= 1) xyzz_dbl_ni(s);
        xyzz_add_ni(v, s);
    }
    xyzz_store(sm + (size_t)t * 128, v);
    __syncthreads();
    for (uint32_t d = MSM_RT / 2; d >= 1; d >>= 1) {
        if (t < d) {
            XYZZ o = xyzz_load(sm + (size_t)(t + d) * 128);
            xyzz_add_ni(v, o);
            xyzz_store(sm + (size_t)t * 128, v);
        }
        __syncthreads();
    }
    if (t == 0) {
        // + (blk * RT * RM) * total_run
        XYZZ off = total_run;
        xyzz_mul_small(off, blk * MSM_RT * rm);
        xyzz_add_ni(v, off);
        xyzz_store(block_out + (size_t)blockIdx.x * 128, v);
    }
}

// ---------------------------------------------------------------- 7. window combine
// One block.  For every window: tree-sum of its block results; then (only when there is more
// than one bucket set) Horner over windows with c doublings each.  Output: 96 B Jacobian
// (X*ZZ, Y*ZZZ, ZZ) -- NOT normalised; the host entry points normalise after the 96-byte
// read-back (one field inversion is ~100x cheaper on a CPU core than on one GPU thread).
constexpr int MSM_FT = 128;

__device__ __forceinline__ void xyzz_store_jacobian(char* out, const XYZZ& p) {
    if (p.is_identity()) {
        fp_store<FqParams>(out, Fq::zero());
        fp_store<FqParams>(out + 32, Fq::one());
        fp_store<FqParams>(out + 64, Fq::zero());
    } else {
        fp_store<FqParams>(out, FQ_MUL(p.x, p.zz));
        fp_store<FqParams>(out + 32, FQ_MUL(p.y, p.zzz));
        fp_store<FqParams>(out + 64, p.zz);
    }
}

__global__ void __launch_bounds__(MSM_FT)
msm_final_kernel(const char* __restrict__ block_out, uint32_t c, uint32_t nsets, uint32_t blocks_per_window,
                 char* __restrict__ window_sums, char* __restrict__ out) {
    __shared__ uint4 fin_smem[MSM_FT * 8];
    char* sm = reinterpret_cast<char*>(fin_smem);
    const uint32_t t = threadIdx.x;
    for (uint32_t w = 0; w < nsets; w++) {
        XYZZ acc = XYZZ::identity();
        for (uint32_t b = t; b < blocks_per_window; b += MSM_FT) {
            XYZZ o = xyzz_load(block_out + ((size_t)w * blocks_per_window + b) * 128);
            xyzz_add_ni(acc, o);
        }
        xyzz_store(sm + (size_t)t * 128, acc);
        __syncthreads();
        for (uint32_t d = MSM_FT / 2; d >= 1; d >>= 1) {
            if (t < d) {
                XYZZ o = xyzz_load(sm + (size_t)(t + d) * 128);
                xyzz_add_ni(acc, o);
                xyzz_store(sm + (size_t)t * 128, acc);
            }
            __syncthreads();
        }
        if (t == 0) xyzz_store(window_sums + (size_t)w * 128, acc);
        __syncthreads();
    }
    if (t != 0) return;
    XYZZ acc = xyzz_load(window_sums + (size_t)(nsets - 1) * 128);
    for (int w2 = (int)nsets - 2; w2 >= 0; w2--) {
        for (uint32_t i = 0; i < c; i++) xyzz_dbl_ni(acc);
        XYZZ o = xyzz_load(window_sums + (size_t)w2 * 128);
        xyzz_add_ni(acc, o);
    }
    xyzz_store_jacobian(out, acc);
}

// Sum of `count` Jacobian points (96 B each) -> Jacobian (un-normalised).  Single thread.
// (combine of per-GPU / per-chunk partials: arithmetic.rs:428-435 does this on the host)
__global__ void g1_sum_kernel(const char* __restrict__ pts, uint32_t count, char* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ acc = XYZZ::identity();
    for (uint32_t i = 0; i < count; i++) {
        const char* p = pts + (size_t)i * 96;
        XYZZ o = xyzz_from_jacobian(fp_load<FqParams>(p), fp_load<FqParams>(p + 32), fp_load<FqParams>(p + 64));
        xyzz_add_ni(acc, o);
    }
    xyzz_store_jacobian(out, acc);
}

// `groups` independent sums at once: out[g] = sum_r pts[r * groups + g], r < count (the layout an all-gather of every
// rank's `groups` partials produces: rank-major).  One thread per group.
__global__ void g1_sum_groups_kernel(const char* __restrict__ pts, uint32_t count, uint32_t groups, char* __restrict__ out) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    XYZZ acc = XYZZ::identity();
    for (uint32_t r = 0; r < count; r++) {
        const char* p = pts + ((size_t)r * groups + g) * 96;
        XYZZ o = xyzz_from_jacobian(fp_load<FqParams>(p), fp_load<FqParams>(p + 32), fp_load<FqParams>(p + 64));
        xyzz_add_ni(acc, o);
    }
    xyzz_store_jacobian(out + (size_t)g * 96, acc);
}

// ---------------------------------------------------------------- SRS window tables
// table[w * n + i] = 2^(c*w) * P_i (affine).  One launch per window: c doublings in XYZZ, then
// one shared field inversion per thread for its PRE_K points (Montgomery's trick).
constexpr int PRE_K = 8;
__global__ void __launch_bounds__(128)
srs_precompute_kernel(char* __restrict__ table, unsigned long long n, uint32_t w, uint32_t c) {
    const unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * PRE_K;
    if (i0 >= n) return;
    XYZZ pts[PRE_K];
    Fq prefix[PRE_K];
    Fq run = Fq::one();
    const char* src = table + (size_t)(w - 1) * n * 64;
    char* dst = table + (size_t)w * n * 64;
    for (int j = 0; j < PRE_K; j++) {
        const unsigned long long i = i0 + j;
        pts[j] = XYZZ::identity();
        prefix[j] = run;
        if (i >= n) continue;
        Affine a = affine_load(src + i * 64);
        if (a.is_identity()) continue;
        XYZZ p = xyzz_dbl_affine(a);
        for (uint32_t d = 1; d < c; d++) xyzz_dbl_ni(p);
        pts[j] = p;
        run = FQ_MUL(run, FQ_MUL(p.zz, p.zzz));
    }
    Fq inv = fp_inv<FqParams>(run);
    for (int j = PRE_K - 1; j >= 0; j--) {
        const unsigned long long i = i0 + j;
        if (i >= n) continue;
        if (pts[j].is_identity()) {
            fp_store<FqParams>(dst + i * 64, Fq::zero());
            fp_store<FqParams>(dst + i * 64 + 32, Fq::zero());
            continue;
        }
        Fq den_inv = FQ_MUL(inv, prefix[j]);                      // 1 / (zz * zzz)
        inv = FQ_MUL(inv, FQ_MUL(pts[j].zz, pts[j].zzz));
        fp_store<FqParams>(dst + i * 64, FQ_MUL(pts[j].x, FQ_MUL(den_inv, pts[j].zzz)));
        fp_store<FqParams>(dst + i * 64 + 32, FQ_MUL(pts[j].y, FQ_MUL(den_inv, pts[j].zz)));
    }
}

// ---------------------------------------------------------------- synthetic SRS
// bases[i] = [h(seed, i)] G with h a 64-bit splitmix64 hash (never 0), affine, Montgomery.
// Used by the benchmark and the full-size property tests: MSM(s, bases) must equal
// [sum_i s_i * h_i mod r] G, which the host can check with O(n) field work.
__device__ __forceinline__ unsigned long long msm_splitmix(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(128)
srs_synth_kernel(char* __restrict__ bases, unsigned long long n, unsigned long long first, unsigned long long seed) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long h = msm_splitmix(seed ^ msm_splitmix(first + i));
    if (h == 0) h = 1;
    Affine gen;
    gen.x = Fq::one();
    gen.y = fp_dbl<FqParams>(Fq::one());
    XYZZ acc = XYZZ::identity();
    for (int b = 63; b >= 0; b--) {
        xyzz_dbl_ni(acc);
        if ((h >> b) & 1ull) xyzz_madd(acc, gen);
    }
    Affine a = xyzz_to_affine(acc);
    fp_store<FqParams>(bases + i * 64, a.x);
    fp_store<FqParams>(bases + i * 64 + 32, a.y);
}

// bases[i] = [k_i] G for Montgomery-form Fr scalars resident on the device: the point side of
// Params::unsafe_setup (poly/commitment.rs:63-112: g[i] = [s^i] G, g_lagrange[i] = [l_i(s)] G), one thread per
// point, plain double-and-add from the top bit (setup-time work, 254 doublings + ~127 mixed adds per point)
__global__ void __launch_bounds__(128)
srs_from_scalars_kernel(char* __restrict__ bases, const uint4* __restrict__ scalars, unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr k = fp_from_mont<FrParams>(fp_load<FrParams>(scalars + 2 * i));
    Affine gen;
    gen.x = Fq::one();
    gen.y = fp_dbl<FqParams>(Fq::one());
    XYZZ acc = XYZZ::identity();
    for (int b = 253; b >= 0; b--) {
        xyzz_dbl_ni(acc);
        if ((k.v[b >> 5] >> (b & 31)) & 1u) xyzz_madd(acc, gen);
    }
    Affine a = xyzz_to_affine(acc);
    fp_store<FqParams>(bases + i * 64, a.x);
    fp_store<FqParams>(bases + i * 64 + 32, a.y);
}

// ---------------------------------------------------------------- element-wise test kernels
// op: 0 mul, 1 add, 2 sub, 3 sqr, 4 Shoup constant multiplication (Fr), 5 / 6 fused x*y +- y*y
template <class P>
__device__ __forceinline__ Fp<P> field_vec_shoup(const Fp<P>& x, const Fp<P>& y) { return fp_mul<P>(x, y); }
template <>
__device__ __forceinline__ Fr field_vec_shoup<FrParams>(const Fr& x, const Fr& y) {
    return fr_mul_shoup(x, fp_from_mont<FrParams>(y), fr_shoup_companion(y));
}
template <class P>
__global__ void field_vec_kernel(const uint4* a, const uint4* b, uint4* o, unsigned long long n, int op) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<P> x = fp_load<P>(a + 2 * i), y = fp_load<P>(b + 2 * i), r;
    switch (op) {
    case 0: r = fp_mul<P>(x, y); break;
    case 1: r = fp_add<P>(x, y); break;
    case 2: r = fp_sub<P>(x, y); break;
    case 3: r = fp_sqr<P>(x); break;
    case 4:   // (Fr only): x * y through the Shoup path, y treated as a Montgomery-form constant
        r = field_vec_shoup(x, y);
        break;
    case 5: r = fp_mul2_add<P>(x, y, y, y); break;   // x*y + y*y with one reduction
    default: r = fp_mul2_sub<P>(x, y, y, y); break;  // 6: x*y - y*y
    }
    fp_store<P>(o + 2 * i, r);
}

// Wide-MAC throughput probe: every thread runs `iters` dependent Montgomery products on
// ILP independent chains (the MSM / NTT instruction mix: IMAD.WIDE with carry).
template <int ILP>
__global__ void __launch_bounds__(256) imad_probe_kernel(uint4* sink, int iters) {
    Fq x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
        x[j] = Fq::one();
        x[j].v[0] ^= threadIdx.x + j;
    }
    Fq y = Fq::r2();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = fp_mul_cios<FqParams>(x[j], y);   // the schoolbook mix: the roofline's denominator
    }
    Fq acc = x[0];
#pragma unroll
    for (int j = 1; j < ILP; j++) acc = fp_add<FqParams>(acc, x[j]);
    if (acc.v[0] == 0x12345678u && acc.v[7] == 0x9abcdef0u) fp_store<FqParams>(sink, acc);
}

// Same probe for the generated variants: KIND 2 = fp_mul_kara, 3 = fp_sqr_sos, 4 = fp_mul2_add (two products, one reduction)
template <int ILP, int KIND>
__global__ void __launch_bounds__(256) mul_probe_kernel(uint4* sink, int iters) {
    Fq x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
        x[j] = Fq::one();
        x[j].v[0] ^= threadIdx.x + j;
    }
    Fq y = Fq::r2();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            x[j] = fp_mul2_add<FqParams>(x[j], y, y, x[j]);
        }
    }
    Fq acc = x[0];
#pragma unroll
    for (int j = 1; j < ILP; j++) acc = fp_add<FqParams>(acc, x[j]);
    if (acc.v[0] == 0x12345678u && acc.v[7] == 0x9abcdef0u) fp_store<FqParams>(sink, acc);
}

// Same probe for the Shoup constant multiplication (NTT butterflies).
template <int ILP>
__global__ void __launch_bounds__(256) shoup_probe_kernel(uint4* sink, int iters) {
    Fr x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
        x[j] = Fr::one();
        x[j].v[0] ^= threadIdx.x + j;
    }
    const Fr ym = Fr::r2();
    const Fr y = fp_from_mont<FrParams>(ym), yp = fr_shoup_companion(ym);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = fr_mul_shoup(x[j], y, yp);
    }
    Fr acc = x[0];
#pragma unroll
    for (int j = 1; j < ILP; j++) acc = fp_add<FrParams>(acc, x[j]);
    if (acc.v[0] == 0x12345678u && acc.v[7] == 0x9abcdef0u) fp_store<FrParams>(sink, acc);
}

// Raw multiplier-pipe probes, independent of the field code (the roofline's denominator must not be "how fast my own
// fp_mul runs"): ILP independent accumulator chains of ONE instruction kind, no carries between instructions, no loads.
//   KIND 0: IMAD.WIDE.U32 Rd, Ra, b, Rd  (32 x 32 + 64 -> 64: every limb product of the bignum kernels; written as the
//           mad.lo.cc / madc.hi pair the field code uses, which ptxas fuses into one IMAD.WIDE -- check with
//           tools/sass_stats.py pipe_probe)
//   KIND 1: IMAD Rd, Rd, b, Rc           (32 x 32 + 32 -> 32, the "64 results per clock per SM" instruction of the CUDA
//           programming guide's throughput table)
// One multiplicand is the chain's own running value so that ptxas cannot fold the products into additions.
template <int ILP, int KIND>
__global__ void __launch_bounds__(256) pipe_probe_kernel(unsigned long long* sink, int iters, uint32_t b) {
    uint32_t lo[ILP], hi[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) { lo[j] = threadIdx.x * 0x9E3779B9u + j; hi[j] = threadIdx.x + 17 * j + 1; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int j = 0; j < ILP; j++) {
                if (KIND == 0) {
                    asm volatile("{\n\t.reg .u32 x;\n\tmov.u32 x, %1;\n\tmad.lo.cc.u32 %0, x, %2, %0;\n\t"
                                 "madc.hi.u32 %1, x, %2, %1;\n\t}"
                                 : "+r"(lo[j]), "+r"(hi[j]) : "r"(b));
                } else {
                    asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(lo[j]) : "r"(hi[j]), "r"(b));
                }
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) x += lo[j] ^ hi[j];
    if (x == 0x9abcdef0u) *sink = x;
}

// FP64 FMA throughput probe (is the fp64 pipe a usable second multiplier on this part?)
template <int ILP>
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* sink, int iters) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = 1.0 + 1e-9 * (threadIdx.x + j);
    const double a = 1.0000001, b = 1e-7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = fma(x[j], a, b);
    }
    double acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) acc += x[j];
    if (acc == 123.456) *sink = acc;
}

}  // namespace b2
