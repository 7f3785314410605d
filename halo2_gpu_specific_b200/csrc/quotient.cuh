// quotient.cuh -- quotient-polynomial evaluation (evaluate_h) as ONE fused per-row kernel.
//
// Replaces the row loops of Evaluator::evaluate_h (halo2_proofs/src/plonk/evaluation.rs:778-1226;
// the `cuda` variant :1229-1987 runs one OpenCL-style kernel per operator over whole columns and
// juggles a buffer cache).  Here every extended-domain row is one thread that interprets a small
// straight-line program: the circuit's Calculations (evaluation.rs:95-112), the Horner fold
// `value = value * y + part` and the permutation / lookup / shuffle terms are all lowered by the
// host (api_quotient.inl) to four primitive field instructions whose operands are
//   constants, per-proof challenges, columns (base pointer table + rotation), the coset point
//   x0 * x_step^row, or intermediates.
// Intermediates live in shared memory (one 32-byte slot per live value per thread, two 16-byte
// planes so that consecutive threads hit consecutive banks); the host allocates slots from live
// ranges, so the footprint is the program's maximum live width, not its length.  Columns are read
// once per use through the read-only path (consecutive threads = consecutive rows: every warp load
// is 1 KB contiguous), nothing is materialised between terms, and the division by the vanishing
// polynomial (poly/domain.rs:354-373) is folded into the store.
#pragma once
#include "fp.cuh"

namespace b2 {

// Q_MUL2ADD / Q_MUL2SUB: a * b +- c * d under ONE Montgomery reduction (fp_mul2_add, 1.44 product-equivalents instead of
// 2 + an addition).  The host fuses an ADD / SUB whose operands are two products nobody else reads -- gates are sums of
// products, so this is a quarter of all products of a zkWasm-sized program.  Such an instruction takes two QInstr
// entries: {op | dst, a, b, c} and {Q_EXT, d}.
enum QOp : uint32_t { Q_ADD = 0, Q_SUB = 1, Q_MUL = 2, Q_NEG = 3, Q_COPY = 4, Q_MUL2ADD = 5, Q_MUL2SUB = 6, Q_EXT = 7 };
enum QKind : uint32_t {
    QK_CONST = 0,   // constants[index]
    QK_SLOT = 1,    // intermediate in slot `index`
    QK_COLUMN = 2,  // columns[index][(row + rot_off[rot]) & mask]   (all column tables concatenated)
    QK_CHAL = 3,    // challenges[index] (per-proof values and their powers)
    QK_COSET_X = 4, // x0 * x_step^row
};

// operand word: kind << 28 | rot << 20 | index
__host__ __device__ __forceinline__ uint32_t q_operand(uint32_t kind, uint32_t index, uint32_t rot) {
    return (kind << 28) | ((rot & 0xffu) << 20) | (index & 0xfffffu);
}

struct QInstr {          // 16 bytes, read with one uniform 128-bit load
    uint32_t op_dst;     // op | dst_slot << 8
    uint32_t a, b;       // operand words
    uint32_t pad;        // third operand of a fused instruction (the fourth sits in the following Q_EXT entry's `a`)
};

constexpr int Q_THREADS = 128;
constexpr int Q_XLO_BITS = 10;

struct QArgs {
    const uint4* prog;            // QInstr[n_instr]
    uint32_t n_instr;
    uint32_t result;              // operand word of the value to store
    const Fr* constants;
    const Fr* challenges;
    const uint4* const* columns;  // device array of column base pointers
    const uint32_t* rot_off;      // (rotation * rot_scale) mod rows, per rotation index
    const Fr* x_lo;               // x_step^j, j < 2^Q_XLO_BITS
    const Fr* x_hi;               // x0 * x_step^(j << Q_XLO_BITS)
    const Fr* scale;              // optional periodic multiplier of the result (t_evaluations), or nullptr
    uint32_t scale_mask;
    uint32_t n_slots;
    unsigned long long rows;      // power of two
    uint4* out;
    unsigned long long out_stride;   // result of row i goes to out[out_offset + i * out_stride]
    unsigned long long out_offset;
    uint4* slot_spill;            // global slots: all of them when they do not fit in shared memory, or the global class
    uint32_t n_slots_shared;      // hybrid programs: slots below this index are shared memory, the others global
    unsigned long long row_begin; // rows [row_begin, row_begin + row_count) are evaluated (a rank's share);
    unsigned long long row_count; // result of row i still goes to out[out_offset + (i - row_begin) * out_stride]
};

// HYB (with SMEM): two slot classes -- slots below n_sh in shared memory, slots from n_sh on in a per-CTA global scratch
// (the values the host found live across most of the program).  The class test is uniform over the CTA (every thread
// runs the same instruction), so it costs a compare and no divergence.
template <bool SMEM, bool HYB = false>
struct QSlots {
    uint4* lo;
    uint4* hi;
    uint4* glo;
    uint4* ghi;
    uint32_t n_sh;
    __device__ __forceinline__ Fr load(uint32_t s) const {
        uint4 a, b;
        if (HYB && s >= n_sh) {
            a = glo[(size_t)(s - n_sh) * Q_THREADS];
            b = ghi[(size_t)(s - n_sh) * Q_THREADS];
        } else {
            a = lo[(size_t)s * Q_THREADS];
            b = hi[(size_t)s * Q_THREADS];
        }
        Fr r;
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void store(uint32_t s, const Fr& x) const {
        const uint4 a = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]), b = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
        if (HYB && s >= n_sh) {
            glo[(size_t)(s - n_sh) * Q_THREADS] = a;
            ghi[(size_t)(s - n_sh) * Q_THREADS] = b;
        } else {
            lo[(size_t)s * Q_THREADS] = a;
            hi[(size_t)s * Q_THREADS] = b;
        }
    }
};

template <bool SMEM, bool HYB>
__device__ __forceinline__ Fr q_fetch(uint32_t w, const QArgs& a, const QSlots<SMEM, HYB>& slots, unsigned long long row,
                                      const Fr& x_here) {
    const uint32_t kind = w >> 28, index = w & 0xfffffu;
    switch (kind) {
    case QK_CONST: return fp_load_nc<FrParams>(a.constants + index);
    case QK_SLOT: return slots.load(index);
    case QK_COLUMN: {
        const uint32_t rot = (w >> 20) & 0xffu;
        const unsigned long long r = (row + a.rot_off[rot]) & (a.rows - 1ull);
        return fp_load_nc<FrParams>(a.columns[index] + 2ull * r);
    }
    case QK_CHAL: return fp_load_nc<FrParams>(a.challenges + index);
    default: return x_here;
    }
}

template <bool SMEM, bool HYB = false>
__global__ void __launch_bounds__(Q_THREADS) quotient_eval_kernel(const QArgs a) {
    extern __shared__ uint4 q_smem[];
    QSlots<SMEM, HYB> slots;
    slots.glo = slots.ghi = nullptr;
    slots.n_sh = a.n_slots_shared;
    if (SMEM && HYB) {
        const uint32_t n_gl = a.n_slots - a.n_slots_shared;
        slots.lo = q_smem + threadIdx.x;
        slots.hi = q_smem + (size_t)a.n_slots_shared * Q_THREADS + threadIdx.x;
        uint4* base = a.slot_spill + (size_t)blockIdx.x * n_gl * Q_THREADS * 2;
        slots.glo = base + threadIdx.x;
        slots.ghi = base + (size_t)n_gl * Q_THREADS + threadIdx.x;
    } else if (SMEM) {
        slots.lo = q_smem + threadIdx.x;
        slots.hi = q_smem + (size_t)a.n_slots * Q_THREADS + threadIdx.x;
    } else {
        uint4* base = a.slot_spill + (size_t)blockIdx.x * a.n_slots * Q_THREADS * 2;
        slots.lo = base + threadIdx.x;
        slots.hi = base + (size_t)a.n_slots * Q_THREADS + threadIdx.x;
    }
    for (unsigned long long rel = (unsigned long long)blockIdx.x * Q_THREADS + threadIdx.x; rel < a.row_count;
         rel += (unsigned long long)gridDim.x * Q_THREADS) {
        const unsigned long long row = a.row_begin + rel;
        // coset point of this row (beta_term of evaluation.rs:1018-1019): two-level table, one product
        Fr x_here = Fr::zero();
        if (a.x_lo != nullptr)
            x_here = fp_mul<FrParams>(fp_load_nc<FrParams>(a.x_lo + (row & ((1u << Q_XLO_BITS) - 1u))),
                                      fp_load_nc<FrParams>(a.x_hi + (row >> Q_XLO_BITS)));
#pragma unroll 1
        for (uint32_t pc = 0; pc < a.n_instr; pc++) {
            const uint4 ins = __ldg(a.prog + pc);
            const uint32_t op = ins.x & 0xffu, dst = ins.x >> 8;
            Fr x = q_fetch<SMEM, HYB>(ins.y, a, slots, row, x_here);
            Fr r;
            if (op == Q_NEG) {
                r = fp_neg<FrParams>(x);
            } else if (op == Q_COPY) {
                r = x;
            } else if (op >= Q_MUL2ADD) {
                const uint4 ext = __ldg(a.prog + pc + 1);
                pc++;
                const Fr y = q_fetch<SMEM, HYB>(ins.z, a, slots, row, x_here);
                const Fr c = q_fetch<SMEM, HYB>(ins.w, a, slots, row, x_here);
                const Fr d = q_fetch<SMEM, HYB>(ext.y, a, slots, row, x_here);
                r = (op == Q_MUL2ADD) ? fp_mul2_add<FrParams>(x, y, c, d) : fp_mul2_sub<FrParams>(x, y, c, d);
            } else {
                Fr y = q_fetch<SMEM, HYB>(ins.z, a, slots, row, x_here);
                if (op == Q_MUL) r = fp_mul<FrParams>(x, y);
                else if (op == Q_ADD) r = fp_add<FrParams>(x, y);
                else r = fp_sub<FrParams>(x, y);
            }
            slots.store(dst, r);
        }
        Fr res = q_fetch<SMEM, HYB>(a.result, a, slots, row, x_here);
        if (a.scale != nullptr) res = fp_mul<FrParams>(res, fp_load_nc<FrParams>(a.scale + (row & a.scale_mask)));
        fp_store<FrParams>(a.out + 2ull * (a.out_offset + rel * a.out_stride), res);
    }
}

}  // namespace b2
