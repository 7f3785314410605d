// ntt.cuh -- radix-2 NTT over BN254 Fr, natural order in -> natural order out.
//
// Replaces best_fft / gpu_fft / gpu_ifft (halo2_proofs/src/arithmetic.rs:495-645) and
// the transforms of EvaluationDomain (poly/domain.rs:233-414).  Field elements are
// canonical, so any exact DFT is bit-identical to the reference's DIT loop.
//
// Decomposition: N = 2^k is split into P digits of m_1..m_P bits (m_i <= 12).  Pass p
// runs, for every setting of the other index bits, one 2^m_p-point sub-NTT that lives
// entirely in shared memory (bit-reversed on load, DIT stages grouped three at a time so
// each thread keeps 8 elements in registers between shared-memory exchanges), then
// multiplies by the inter-pass twiddle w_N^(i_{p+1} * o_partial) taken from a two-level
// table (w^lo * w^hi).  Passes 1..P-1 work in place; the last pass writes each result to
// its digit-reversed (= natural) position, so it needs out != in when P > 1.
//
// Fused into the passes (reference does each as a separate sweep):
//   first pass : zero padding n_in -> 2^k (domain.rs:280) and the period-3 coset scaling
//                {1, zeta, zeta^2}[i % 3] (distribute_powers_zeta, domain.rs:382-398)
//   1st twiddle: the iNTT divisor 2^-k (domain.rs:404-409) is folded into the hi table
//   last pass  : the inverse period-3 scaling and truncation of extended_to_coeff
//                (domain.rs:341-347)
#pragma once
#include <cooperative_groups.h>

#include "fp.cuh"
#include "fp_shoup.cuh"

namespace b2 {

constexpr int NTT_MAX_PASSES = 4;
#ifndef NTT_RMAX
#define NTT_RMAX 3      // DIT stages per shared-memory round trip (2^NTT_RMAX elements per thread)
#endif
#ifndef NTT_CL_MINB
#define NTT_CL_MINB 2      // resident cluster-kernel CTAs per SM the register allocation must allow
#endif

struct NttPassArgs {
    const uint4* in;
    uint4* out;
    unsigned long long in_col_stride;   // elements between batch columns
    unsigned long long out_col_stride;
    unsigned long long n_in;            // first pass: elements present in `in` (rest read as zero)
    unsigned long long n_out;           // last pass: only outputs with index < n_out are stored
    uint32_t log_n;
    uint32_t m;                         // digit size of this pass
    uint32_t s_lo;                      // number of index bits below this digit
    uint32_t pass, npass;
    uint32_t mm[NTT_MAX_PASSES];        // digit sizes m_1..m_P
    uint32_t tw_h;                      // split of the two-level table: e = hi * 2^tw_h + lo
    const Fr* tw_sub;                   // w_{2^m}^j, j < 2^(m-1) in Montgomery form, or the stage-major four-plane
                                        // (w, floor(w 2^256 / r)) table of ntt_shoup_table_kernel (Shoup kernels)
    const Fr* tw_lo;                    // w_N^j, j < 2^tw_h
    const Fr* tw_hi;                    // w_N^(j * 2^tw_h) (times the divisor for pass 0 of an iNTT)
    const Fr* tw_full;                  // optional: w_N^e for e < N/2 (times the divisor), used by pass 0;
                                        // w^(e + N/2) = -w^e.  One product per element instead of two.
    int coset_in;                       // first pass: multiply x[i] by zin[i % 3 - 1]
    int coset_out;                      // last pass : multiply X[o] by zout[o % 3 - 1]
    int scale_out;                      // last pass : multiply everything by `scale` (P == 1 iNTT)
    uint32_t cl_log;                    // log2 of the cluster size that owns one tile (0 = single CTA)
    const Fr* in_scale;                 // first pass: optional table, x[i] *= in_scale[i] (g^i: evaluation on the coset g*H)
    Fr zin1, zin2, zout1, zout2, scale;
};

__device__ __forceinline__ uint32_t ntt_swz(uint32_t i, uint32_t hs) {
    return i ^ ((i >> 3) & 7u) ^ ((i >> hs) & 7u);
}

__device__ __forceinline__ Fr sm_ld(const uint4* lo, const uint4* hi, uint32_t idx) {
    uint4 a = lo[idx], b = hi[idx];
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sm_st(uint4* lo, uint4* hi, uint32_t idx, const Fr& x) {
    lo[idx] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    hi[idx] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

// Shoup twiddle table of one digit size m, built from the Montgomery-form table tw_mont[j] = w_{2^m}^j:
// stage-major (stage s = 1..m owns the 2^(s-1) entries w_{2^s}^jj at positions 2^(s-1) - 1 + jj) so that the
// lanes of a warp, which hold consecutive jj, read consecutive entries; and split into four 16-byte planes
// (w low / high half, floor(w 2^256 / r) low / high half), each 2^m entries long, so that every LDG.128 of a
// warp is one contiguous 512-byte run.  (The stride-2^(m-s) gathers from a single table were the NTT's second
// bottleneck: 5.8x the shared-memory wavefronts on the L1 data pipe.)
__global__ void ntt_shoup_table_kernel(const Fr* __restrict__ tw_mont, uint4* __restrict__ planes, uint32_t m) {
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = (1u << m) - 1u;
    if (pos >= total) return;
    const uint32_t s = 32u - __clz(pos + 1u);              // stage: 2^(s-1) <= pos + 1 < 2^s
    const uint32_t jj = pos + 1u - (1u << (s - 1));
    const Fr wm = fp_load<FrParams>(tw_mont + ((size_t)jj << (m - s)));
    const Fr w = fp_from_mont<FrParams>(wm), wp = fr_shoup_companion(wm);
    const size_t plane = (size_t)1 << m;
    planes[pos] = make_uint4(w.v[0], w.v[1], w.v[2], w.v[3]);
    planes[plane + pos] = make_uint4(w.v[4], w.v[5], w.v[6], w.v[7]);
    planes[2 * plane + pos] = make_uint4(wp.v[0], wp.v[1], wp.v[2], wp.v[3]);
    planes[3 * plane + pos] = make_uint4(wp.v[4], wp.v[5], wp.v[6], wp.v[7]);
}

// ---- lazy (Harvey) butterflies: values live in [0, 4r) between stages (4r < 2^256 for BN254's r) -------------
// if x >= 2r: x -= 2r
__device__ __forceinline__ void fr_csub_2r(Fr& x) {
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
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(borrow)
        : "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]), "r"(x.v[6]), "r"(x.v[7]),
          "n"(0xe0000002u), "n"(0x87c3eb27u), "n"(0xf372e122u), "n"(0x5067d090u), "n"(0x0302b0bau), "n"(0x70a08b6du),
          "n"(0xc2634053u), "n"(0x60c89ce5u));
#pragma unroll
    for (int i = 0; i < 8; i++) x.v[i] = borrow ? x.v[i] : t[i];
}
// [0, 4r) -> [0, r)
__device__ __forceinline__ void fr_reduce_4r(Fr& x) {
    fr_csub_2r(x);
    fp_reduce_once<FrParams>(x.v);
}
// a, t < 2r:  b = a - t + 2r  (in (0, 4r)),  a = a + t  (< 4r); no reductions
__device__ __forceinline__ void fr_lazy_addsub(Fr& a, Fr& b, const Fr& t) {
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=&r"(b.v[0]), "=&r"(b.v[1]), "=&r"(b.v[2]), "=&r"(b.v[3]), "=&r"(b.v[4]), "=&r"(b.v[5]), "=&r"(b.v[6]),
          "=&r"(b.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "n"(0xe0000002u), "n"(0x87c3eb27u), "n"(0xf372e122u), "n"(0x5067d090u), "n"(0x0302b0bau), "n"(0x70a08b6du),
          "n"(0xc2634053u), "n"(0x60c89ce5u));
    asm("sub.cc.u32 %0, %0, %8;\n\t"
        "subc.cc.u32 %1, %1, %9;\n\t"
        "subc.cc.u32 %2, %2, %10;\n\t"
        "subc.cc.u32 %3, %3, %11;\n\t"
        "subc.cc.u32 %4, %4, %12;\n\t"
        "subc.cc.u32 %5, %5, %13;\n\t"
        "subc.cc.u32 %6, %6, %14;\n\t"
        "subc.u32 %7, %7, %15;"
        : "+r"(b.v[0]), "+r"(b.v[1]), "+r"(b.v[2]), "+r"(b.v[3]), "+r"(b.v[4]), "+r"(b.v[5]), "+r"(b.v[6]), "+r"(b.v[7])
        : "r"(t.v[0]), "r"(t.v[1]), "r"(t.v[2]), "r"(t.v[3]), "r"(t.v[4]), "r"(t.v[5]), "r"(t.v[6]), "r"(t.v[7]));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(a.v[0]), "+r"(a.v[1]), "+r"(a.v[2]), "+r"(a.v[3]), "+r"(a.v[4]), "+r"(a.v[5]), "+r"(a.v[6]), "+r"(a.v[7])
        : "r"(t.v[0]), "r"(t.v[1]), "r"(t.v[2]), "r"(t.v[3]), "r"(t.v[4]), "r"(t.v[5]), "r"(t.v[6]), "r"(t.v[7]));
}

// Where a sub-NTT's butterfly twiddles come from: the stage-major four-plane Shoup table in global memory, and for
// the first sm_stages stages (positions 0 .. 2^sm_stages - 2 of every plane) a copy staged into shared memory by one
// cp.async.bulk per plane at kernel start (TMA bulk copy, completion on an mbarrier).
struct NttTw {
    const Fr* g;            // global table (Shoup planes or Montgomery table)
    const uint4* sm;        // shared-memory copy of the first sm_n entries of each plane (nullptr: none)
    uint32_t sm_stages;     // stages 1 .. sm_stages are served from shared memory
    uint32_t sm_n;          // entries per shared plane
    uint32_t m;             // log size of the whole sub-NTT (plane stride / Montgomery table stride)
};

// butterfly of stage s with twiddle w_{2^s}^jj; LAZY: inputs and outputs in [0, 4r)
template <bool SHOUP, bool LAZY>
__device__ __forceinline__ void ntt_bfly_tw(Fr& a, Fr& b, const NttTw& tw, uint32_t s, uint32_t jj) {
    Fr t;
    if (SHOUP) {
        const uint32_t pos = (1u << (s - 1)) - 1u + jj;
        uint4 w0, w1, p0, p1;
        if (tw.sm != nullptr && s <= tw.sm_stages) {
            const uint4* pl = tw.sm + pos;
            w0 = pl[0]; w1 = pl[tw.sm_n]; p0 = pl[2 * tw.sm_n]; p1 = pl[3 * tw.sm_n];
        } else {
            const uint4* pl = reinterpret_cast<const uint4*>(tw.g) + pos;
            const size_t plane = (size_t)1 << tw.m;
            w0 = __ldg(pl); w1 = __ldg(pl + plane); p0 = __ldg(pl + 2 * plane); p1 = __ldg(pl + 3 * plane);
        }
        Fr w, wp;
        w.v[0] = w0.x; w.v[1] = w0.y; w.v[2] = w0.z; w.v[3] = w0.w;
        w.v[4] = w1.x; w.v[5] = w1.y; w.v[6] = w1.z; w.v[7] = w1.w;
        wp.v[0] = p0.x; wp.v[1] = p0.y; wp.v[2] = p0.z; wp.v[3] = p0.w;
        wp.v[4] = p1.x; wp.v[5] = p1.y; wp.v[6] = p1.z; wp.v[7] = p1.w;
        t = LAZY ? fr_mul_shoup_lazy(b, w, wp) : fr_mul_shoup(b, w, wp);
    } else {
        t = fp_mul<FrParams>(b, fp_load_nc<FrParams>(tw.g + ((size_t)jj << (tw.m - s))));
    }
    if (LAZY) {
        fr_csub_2r(t);
        fr_csub_2r(a);
        fr_lazy_addsub(a, b, t);
    } else {
        b = fp_sub<FrParams>(a, t);
        a = fp_add<FrParams>(a, t);
    }
}
template <bool LAZY>
__device__ __forceinline__ void ntt_bfly_one(Fr& a, Fr& b) {  // twiddle == 1
    Fr t = b;
    if (LAZY) {
        fr_csub_2r(t);
        fr_csub_2r(a);
        fr_lazy_addsub(a, b, t);
    } else {
        b = fp_sub<FrParams>(a, t);
        a = fp_add<FrParams>(a, t);
    }
}

// One group of R consecutive DIT stages (s0+1 .. s0+R) on the shared-memory tile.
template <int R, bool FIRST, bool SHOUP, bool LAZY>
__device__ __forceinline__ void ntt_step(uint4* s_lo4, uint4* s_hi4, const NttTw& tw,
                                         uint32_t mloc, uint32_t s0, uint32_t hs) {
    // tw.m: log size of the whole sub-NTT (twiddle stride); mloc: log size of the part in this CTA
    constexpr int E = 1 << R;
    const uint32_t items = 1u << (mloc - R);
    for (uint32_t w = threadIdx.x; w < items; w += blockDim.x) {
        const uint32_t low = w & ((1u << s0) - 1u);
        const uint32_t base = ((w >> s0) << (s0 + R)) | low;
        Fr x[E];
#pragma unroll
        for (int q = 0; q < E; q++) x[q] = sm_ld(s_lo4, s_hi4, ntt_swz(base + ((uint32_t)q << s0), hs));
#pragma unroll
        for (int st = 0; st < R; st++) {
            // stage s = s0 + st + 1: partner differs in bit `st` of q; twiddle exponent
            // jj = low + (q & (2^st - 1)) * 2^s0, table index jj << (m - s)
#pragma unroll
            for (int q = 0; q < E; q++) {
                if (q & (1 << st)) continue;
                const int lowq = q & ((1 << st) - 1);
                if (FIRST && lowq == 0) {
                    ntt_bfly_one<LAZY>(x[q], x[q | (1 << st)]);
                } else {
                    const uint32_t jj = low + ((uint32_t)lowq << s0);
                    ntt_bfly_tw<SHOUP, LAZY>(x[q], x[q | (1 << st)], tw, s0 + st + 1, jj);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < E; q++) sm_st(s_lo4, s_hi4, ntt_swz(base + ((uint32_t)q << s0), hs), x[q]);
    }
}

// digit-reverse: position bits (high -> low) hold o_1 .. o_np; the natural output index
// has o_1 as its LEAST significant digit.
__device__ __forceinline__ uint64_t ntt_digit_reverse(uint64_t v, const uint32_t* mm, int np) {
    uint64_t r = 0;
    uint32_t placed = 0;
    // peel digits from the low end of v: o_np first
    uint32_t total = 0;
    for (int i = 0; i < np; i++) total += mm[i];
    uint32_t below = total;
    (void)placed;
    for (int i = np - 1; i >= 0; i--) {
        uint64_t d = v & ((1ull << mm[i]) - 1ull);
        v >>= mm[i];
        below -= mm[i];          // bits occupied by digits 0..i-1 in the reversed number
        r |= d << below;
    }
    return r;
}

// CL_LOG = 0: one CTA owns the whole 2^m tile.  CL_LOG = 1, 2: a thread-block cluster of 2 / 4 CTAs
// owns it, 2^(m - CL_LOG) elements (64 KB) per CTA: the first m - CL_LOG DIT stages are local to each
// CTA, the last CL_LOG stages exchange through distributed shared memory.  This keeps 2-3 CTAs
// resident per SM for 2^12 / 2^13-point digits (a single CTA with a 128 KB tile runs alone on its
// SM and exposes its load / store latency), and lets k = 25, 26 run in two passes.
// TWSM > 0: the twiddles of the first min(TWSM, m) stages are staged into shared memory with cp.async.bulk (one bulk
// copy per plane, issued by one thread before the tile is loaded, completion awaited on an mbarrier before the first
// butterfly that needs them).  Dynamic shared memory: tile (32 B << mloc) | 4 planes x 2^TWSM x 16 B | mbarrier.
__device__ __forceinline__ uint32_t ntt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CL_LOG, bool SHOUP, bool LAZY = false, int TWSM = 0>
__device__ __forceinline__ void ntt_pass_impl(const NttPassArgs& a) {
    namespace cg = cooperative_groups;
    extern __shared__ uint4 ntt_smem[];
    const uint32_t m = a.m, mloc = m - CL_LOG, Nloc = 1u << mloc;
    uint4* s_lo4 = ntt_smem;
    uint4* s_hi4 = ntt_smem + Nloc;
    const uint32_t hs = (mloc >= 9) ? (mloc - 3) : 31u;
    uint32_t rank = 0;
    if (CL_LOG > 0) rank = cg::this_cluster().block_rank();

    NttTw tw;
    tw.g = a.tw_sub;
    tw.m = m;
    tw.sm = nullptr;
    tw.sm_stages = 0;
    tw.sm_n = 0;
    if (TWSM > 0 && SHOUP) {
        const uint32_t st = min(m, (uint32_t)TWSM);
        uint4* sm_tw = ntt_smem + 2 * Nloc;
        unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sm_tw + (4u << TWSM));
        tw.sm = sm_tw;
        tw.sm_stages = st;
        tw.sm_n = 1u << TWSM;
        if (threadIdx.x == 0) {
            const uint32_t mb = ntt_smem_u32(mbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t bytes = 16u << st;                       // per plane: entries 0 .. 2^st - 1
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(4u * bytes) : "memory");
            const uint4* g = reinterpret_cast<const uint4*>(a.tw_sub);
#pragma unroll
            for (int pl = 0; pl < 4; pl++) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ntt_smem_u32(sm_tw + (size_t)pl * tw.sm_n)), "l"(g + ((size_t)pl << m)), "r"(bytes),
                               "r"(mb)
                             : "memory");
            }
        }
    }

    const uint64_t line = blockIdx.x >> CL_LOG;
    const uint64_t col = blockIdx.y;
    const uint64_t L = line & ((1ull << a.s_lo) - 1ull);
    const uint64_t H = line >> a.s_lo;
    const uint64_t base_pos = (H << (m + a.s_lo)) | L;
    const uint4* in = a.in + 2ull * col * a.in_col_stride;
    uint4* out = a.out + 2ull * col * a.out_col_stride;
    const bool first = (a.pass == 0), last = (a.pass + 1 == a.npass);

    // ---- load (bit-reversed into the tile): tile position p = bitrev_m(j); this CTA holds the
    // positions whose top CL_LOG bits equal its rank, i.e. the j with low bits bitrev(rank) ----
    const uint32_t jr = CL_LOG ? (__brev(rank) >> (32 - CL_LOG)) : 0u;
    for (uint32_t jl = threadIdx.x; jl < Nloc; jl += blockDim.x) {
        const uint32_t j = (jl << CL_LOG) | jr;
        const uint64_t pos = base_pos + ((uint64_t)j << a.s_lo);
        Fr x;
        if (!first || pos < a.n_in) {
            x = fp_load<FrParams>(in + 2ull * pos);
            if (first && a.coset_in) {
                const uint32_t r3 = (uint32_t)(pos % 3ull);
                if (r3 == 1) x = fp_mul<FrParams>(x, a.zin1);
                else if (r3 == 2) x = fp_mul<FrParams>(x, a.zin2);
            }
            if (first && a.in_scale != nullptr) x = fp_mul<FrParams>(x, fp_load_nc<FrParams>(a.in_scale + pos));
        } else {
            x = Fr::zero();
        }
        const uint32_t pl = mloc ? (__brev(jl) >> (32 - mloc)) : 0u;
        sm_st(s_lo4, s_hi4, ntt_swz(pl, hs), x);
    }
    __syncthreads();   // also publishes the mbarrier initialisation to the waiting threads
    if (TWSM > 0 && SHOUP) {
        // twiddle planes have landed (phase 0 of the mbarrier completes when all four bulk copies are in)
        const uint32_t mb = ntt_smem_u32(ntt_smem + 2 * Nloc + (4u << TWSM));
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "NTT_TW_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@p bra NTT_TW_DONE;\n\t"
            "bra NTT_TW_WAIT;\n\t"
            "NTT_TW_DONE:\n\t"
            "}" ::"r"(mb)
            : "memory");
    }

    // ---- local DIT stages, three per shared-memory round trip ----
    uint32_t s0 = 0;
    {
        const uint32_t r = mloc < NTT_RMAX ? mloc : NTT_RMAX;
        if (r == 3) ntt_step<3, true, SHOUP, LAZY>(s_lo4, s_hi4, tw, mloc, 0, hs);
        else if (r == 2) ntt_step<2, true, SHOUP, LAZY>(s_lo4, s_hi4, tw, mloc, 0, hs);
        else if (r == 1) ntt_step<1, true, SHOUP, LAZY>(s_lo4, s_hi4, tw, mloc, 0, hs);
        s0 = r;
        __syncthreads();
    }
    while (s0 < mloc) {
        const uint32_t r = (mloc - s0) < NTT_RMAX ? (mloc - s0) : NTT_RMAX;
        if (r == 3) ntt_step<3, false, SHOUP, LAZY>(s_lo4, s_hi4, tw, mloc, s0, hs);
        else if (r == 2) ntt_step<2, false, SHOUP, LAZY>(s_lo4, s_hi4, tw, mloc, s0, hs);
        else ntt_step<1, false, SHOUP, LAZY>(s_lo4, s_hi4, tw, mloc, s0, hs);
        s0 += r;
        __syncthreads();
    }

    // ---- cross-CTA stages through distributed shared memory ----
    if (CL_LOG > 0) {
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int q = 0; q < CL_LOG; q++) {
            cluster.sync();   // partner's previous stage is complete and visible
            // stage s = mloc + q + 1: position p pairs with p + 2^(mloc + q): rank bit q selects a / b
            const uint32_t partner = rank ^ (1u << q);
            const bool is_b = (rank >> q) & 1u;
            const uint32_t ra = is_b ? partner : rank;           // rank holding the "a" element
            uint4* r_lo4 = cluster.map_shared_rank(s_lo4, partner);
            uint4* r_hi4 = cluster.map_shared_rank(s_hi4, partner);
            uint4* a_lo = is_b ? r_lo4 : s_lo4;
            uint4* a_hi = is_b ? r_hi4 : s_hi4;
            uint4* b_lo = is_b ? s_lo4 : r_lo4;
            uint4* b_hi = is_b ? s_hi4 : r_hi4;
            // each CTA of the pair takes half of the butterflies
            const uint32_t half = Nloc >> 1;
            const uint32_t i0 = is_b ? half : 0u;
            const uint32_t jj_hi = (ra & ((1u << q) - 1u)) << mloc;   // p_a mod 2^(mloc+q), high part
            for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) {
                const uint32_t i = i0 + t;
                const uint32_t sw = ntt_swz(i, hs);
                Fr xa = sm_ld(a_lo, a_hi, sw);
                Fr xb = sm_ld(b_lo, b_hi, sw);
                ntt_bfly_tw<SHOUP, LAZY>(xa, xb, tw, mloc + q + 1, jj_hi | i);
                sm_st(a_lo, a_hi, sw, xa);
                sm_st(b_lo, b_hi, sw, xb);
            }
        }
        cluster.sync();
    }

    // ---- store: this CTA holds tile positions j = rank * Nloc + il (natural order) ----
    if (!last) {
        // twiddle exponent = i_next * digit_reverse(H, j) * 2^(k - (m_1+..+m_{p+1}))
        const uint32_t m_next = a.mm[a.pass + 1];
        const uint64_t i_next = L >> (a.s_lo - m_next);
        const uint32_t shift = a.s_lo - m_next;
        const uint64_t lo_mask = (1ull << a.tw_h) - 1ull;
        const bool full = first && a.tw_full != nullptr;
        const uint64_t half_mask = (1ull << (a.log_n - 1)) - 1ull;
        for (uint32_t il = threadIdx.x; il < Nloc; il += blockDim.x) {
            const uint32_t j = (rank << mloc) | il;
            Fr x = sm_ld(s_lo4, s_hi4, ntt_swz(il, hs));
            if (LAZY) fr_reduce_4r(x);
            const uint64_t opart = ntt_digit_reverse((H << m) | j, a.mm, (int)a.pass + 1);
            const uint64_t e = (i_next * opart) << shift;
            if (full) {
                // entry 0 carries the iNTT divisor (or 1): skip the product only when it is 1
                if (e != 0 || a.scale_out) {
                    x = fp_mul<FrParams>(x, fp_load_nc<FrParams>(a.tw_full + (e & half_mask)));
                    if (e >> (a.log_n - 1)) x = fp_neg<FrParams>(x);
                }
            } else if (e != 0) {
                Fr t = fp_mul<FrParams>(fp_load_nc<FrParams>(a.tw_lo + (e & lo_mask)),
                                        fp_load_nc<FrParams>(a.tw_hi + (e >> a.tw_h)));
                x = fp_mul<FrParams>(x, t);
            } else if (first) {
                x = fp_mul<FrParams>(x, fp_load_nc<FrParams>(a.tw_hi));  // carries the iNTT divisor (or 1)
            }
            const uint64_t pos = base_pos + ((uint64_t)j << a.s_lo);
            fp_store<FrParams>(out + 2ull * pos, x);
        }
    } else {
        for (uint32_t il = threadIdx.x; il < Nloc; il += blockDim.x) {
            const uint32_t j = (rank << mloc) | il;
            Fr x = sm_ld(s_lo4, s_hi4, ntt_swz(il, hs));
            const uint64_t o = ntt_digit_reverse((H << m) | j, a.mm, (int)a.npass);
            if (o >= a.n_out) continue;
            if (LAZY) fr_reduce_4r(x);
            if (a.scale_out) x = fp_mul<FrParams>(x, a.scale);
            if (a.coset_out) {
                const uint32_t r3 = (uint32_t)(o % 3ull);
                if (r3 == 1) x = fp_mul<FrParams>(x, a.zout1);
                else if (r3 == 2) x = fp_mul<FrParams>(x, a.zout2);
            }
            fp_store<FrParams>(out + 2ull * o, x);
        }
    }
}

__global__ void __launch_bounds__(512) ntt_pass_kernel(const NttPassArgs a) { ntt_pass_impl<0, true>(a); }
__global__ void __launch_bounds__(256, NTT_CL_MINB) ntt_pass_cluster2_kernel(const NttPassArgs a) { ntt_pass_impl<1, true>(a); }
__global__ void __launch_bounds__(256, NTT_CL_MINB) ntt_pass_cluster4_kernel(const NttPassArgs a) { ntt_pass_impl<2, true>(a); }
// Variants selected by B2_NTT_VARIANT (A/B measurements; the default is chosen in ntt_run_dev):
//   1: lazy butterflies   2: lazy + twiddles of stages 1..8 staged into shared memory by TMA bulk copies
//   3: shared-memory twiddles only
// (register-capped builds of variant 1 -- 5 / 6 CTAs per SM at <= 102 / 85 registers, 3 / 2 CTAs at 156 / 180 -- were
// measured in round 2 and removed: throughput follows the warp count up to 4 CTAs and spills beyond, profiles/r2_ncu_summary.md)
constexpr int NTT_TWSM = 8;
#ifndef NTT_DEFAULT_VARIANT
#define NTT_DEFAULT_VARIANT 1
#endif
__global__ void __launch_bounds__(512) ntt_pass_v1_kernel(const NttPassArgs a) { ntt_pass_impl<0, true, true, 0>(a); }
__global__ void __launch_bounds__(256, NTT_CL_MINB) ntt_pass_cluster2_v1_kernel(const NttPassArgs a) { ntt_pass_impl<1, true, true, 0>(a); }
__global__ void __launch_bounds__(256, NTT_CL_MINB) ntt_pass_cluster4_v1_kernel(const NttPassArgs a) { ntt_pass_impl<2, true, true, 0>(a); }
__global__ void __launch_bounds__(512) ntt_pass_v2_kernel(const NttPassArgs a) { ntt_pass_impl<0, true, true, NTT_TWSM>(a); }
__global__ void __launch_bounds__(256, NTT_CL_MINB) ntt_pass_cluster2_v2_kernel(const NttPassArgs a) { ntt_pass_impl<1, true, true, NTT_TWSM>(a); }
__global__ void __launch_bounds__(256, NTT_CL_MINB) ntt_pass_cluster4_v2_kernel(const NttPassArgs a) { ntt_pass_impl<2, true, true, NTT_TWSM>(a); }
__global__ void __launch_bounds__(512) ntt_pass_v3_kernel(const NttPassArgs a) { ntt_pass_impl<0, true, false, NTT_TWSM>(a); }
__global__ void __launch_bounds__(256, NTT_CL_MINB) ntt_pass_cluster2_v3_kernel(const NttPassArgs a) { ntt_pass_impl<1, true, false, NTT_TWSM>(a); }
__global__ void __launch_bounds__(256, NTT_CL_MINB) ntt_pass_cluster4_v3_kernel(const NttPassArgs a) { ntt_pass_impl<2, true, false, NTT_TWSM>(a); }
// Montgomery-twiddle variants (B2_NTT_SHOUP=0: A/B measurements)
__global__ void __launch_bounds__(512) ntt_pass_mont_kernel(const NttPassArgs a) { ntt_pass_impl<0, false>(a); }
__global__ void __launch_bounds__(256) ntt_pass_mont_cluster2_kernel(const NttPassArgs a) { ntt_pass_impl<1, false>(a); }
__global__ void __launch_bounds__(256) ntt_pass_mont_cluster4_kernel(const NttPassArgs a) { ntt_pass_impl<2, false>(a); }

// out[j] = base^(j * mult) * (scale if has_scale), j < count
__global__ void ntt_pow_table_kernel(Fr* out, const Fr base, unsigned long long mult, uint32_t count,
                                     int has_scale, const Fr scale) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    unsigned long long e = (unsigned long long)j * mult;
    Fr acc = Fr::one();
    Fr b = base;
    while (e) {
        if (e & 1ull) acc = fp_mul<FrParams>(acc, b);
        b = fp_sqr<FrParams>(b);
        e >>= 1;
    }
    if (has_scale) acc = fp_mul<FrParams>(acc, scale);
    fp_store<FrParams>(out + j, acc);
}

// out[j] = base^j, j < count: every thread raises base to its first index, then multiplies along
constexpr int POW_SEQ = 64;
__global__ void __launch_bounds__(128) ntt_pow_seq_kernel(Fr* out, const Fr base, unsigned long long count) {
    const unsigned long long j0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * POW_SEQ;
    if (j0 >= count) return;
    Fr acc = Fr::one(), b = base;
    for (unsigned long long e = j0; e; e >>= 1) {
        if (e & 1ull) acc = fp_mul<FrParams>(acc, b);
        b = fp_sqr<FrParams>(b);
    }
    for (int i = 0; i < POW_SEQ && j0 + i < count; i++) {
        fp_store<FrParams>(out + j0 + i, acc);
        acc = fp_mul<FrParams>(acc, base);
    }
}

// a[i] *= t[i % period]  (divide_by_vanishing_poly, poly/domain.rs:354-373)
__global__ void fr_scale_periodic_kernel(uint4* a, const Fr* __restrict__ t, unsigned long long n, uint32_t period_mask) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        Fr x = fp_load<FrParams>(a + 2ull * i);
        x = fp_mul<FrParams>(x, fp_load_nc<FrParams>(t + (i & period_mask)));
        fp_store<FrParams>(a + 2ull * i, x);
    }
}

}  // namespace b2
