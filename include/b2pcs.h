/*
 * b2pcs.h -- C ABI of the B200-native polynomial-commitment engine (BN254 / KZG).
 *
 * This is the drop-in boundary for the one hot path of halo2_proofs as used by
 * DelphinusLab/halo2-gpu-specific: the G1 MSM behind best_multiexp / Params::commit*,
 * and the radix-2 NTT over Fr behind best_fft / EvaluationDomain::{lagrange_to_coeff,
 * coeff_to_extended, extended_to_coeff}.  The reference has no C ABI: under its `cuda`
 * feature the Rust functions cited below transmute their slices to BN254 types and call
 * the ec-gpu-gen Rust API.  Each entry point here names the reference function whose body
 * it replaces (paths relative to halo2_proofs/src); INTEGRATION.md shows the Rust
 * `extern "C"` block and the shim bodies.
 *
 * Data layout (all little-endian, identical to the reference's in-memory types):
 *   Fr, Fq      4 x u64 limbs, Montgomery form a * 2^256 mod p          (32 bytes)
 *   G1Affine    x || y in Fq; the identity is encoded as (0, 0)         (64 bytes;
 *               b2_srs_register takes a stride so a 72-byte struct with a trailing
 *               flag also works)
 *   G1          X || Y || Z Jacobian, x = X/Z^2, y = Y/Z^3; identity Z=0 (96 bytes).
 *               Results are returned NORMALISED (Z = 1 in Montgomery form, or
 *               (0, 1, 0) for the identity): projective representatives are not
 *               unique and the reference always normalises before the transcript.
 *
 * Conventions: every function returns B2_OK (0) or a negative error code and never
 * panics/aborts; b2_last_error() returns a thread-local message.  The library never
 * keeps a caller pointer past the call.  All functions are thread-safe; calls on one
 * device are serialised internally (the reference serialises per device with
 * acquire_gpu/release_gpu, arithmetic.rs:313-331).  There is NO CPU fallback: without a
 * usable CUDA device every compute entry point returns B2_ERR_CUDA.
 */
#ifndef B2PCS_H
#define B2PCS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_ERR_ARG (-1)      /* bad argument / length mismatch (reference: assert_eq!, arithmetic.rs:466,569) */
#define B2_ERR_CUDA (-2)     /* CUDA runtime error (reference: .unwrap()/.expect panics, arithmetic.rs:356-360,509) */
#define B2_ERR_OOM (-3)      /* device or pinned-host allocation failed */
#define B2_ERR_BOUND (-4)    /* a scalar exceeded the max_bits contract of *_with_bound */
#define B2_ERR_HANDLE (-5)   /* unknown / freed handle */

typedef uint64_t b2_handle_t;

/* ---- library / device ------------------------------------------------------------- */
int b2_version(void);
const char* b2_last_error(void);
int b2_device_count(void);
/* Select the device used by subsequent calls from this host thread (default 0).
 * Reference: the device index popped from GPU_LOCK (plonk/prover.rs:56-74). */
int b2_set_device(int device);
int b2_get_device(void);
int b2_synchronize(void);
/* A stream of the current device for the `stream` arguments of the _dev entry points, for callers that do not bring
 * their own CUDA runtime (the Python mirror; a Rust host would pass its own cudaStream_t).  Work a _dev call enqueues
 * on it runs asynchronously to the library's own lanes: the prover uses one to transform the advice columns already
 * uploaded while the next ones still cross PCIe. */
int b2_stream_create(void** stream);
int b2_stream_synchronize(void* stream);
int b2_stream_destroy(void* stream);
/* number of kernels this library launched on the current device since the last reset */
uint64_t b2_launch_count(int reset);

/* ---- SRS (bases resident in HBM) -------------------------------------------------- */
/* Upload n affine points once; they stay resident until b2_srs_free.  Replaces the
 * per-call re-upload of `bases` in gpu_multiexp_single_gpu_with_bound
 * (arithmetic.rs:349-360) for Params::g / Params::g_lagrange (poly/commitment.rs:23-29). */
int b2_srs_register(const void* bases, size_t n, size_t stride_bytes, b2_handle_t* out);
/* Synthetic bases for benchmarks / property tests: bases[i] = [h(seed, first+i)] G with
 * h = 64-bit splitmix64 hash, generated on the device. */
int b2_srs_synthetic(size_t n, uint64_t first_index, uint64_t seed, b2_handle_t* out);
/* bases[i] = [k_i] G for n Montgomery-form Fr scalars resident on the device: the point side of
 * Params::unsafe_setup (poly/commitment.rs:63-112: g[i] = [s^i] G and g_lagrange[i] = [l_i(s)] G); the scalar side
 * (powers, batch inversion) is composed from b2_prefix_scan_dev / b2_batch_invert_dev / b2_fr_vec_dev by the host
 * mirror (commitment.py Params.unsafe_setup).  The new SRS is resident like a registered one. */
int b2_srs_from_scalars_dev(const void* d_scalars, size_t n, b2_handle_t* out);
/* Build the window table of a resident SRS: table[w][i] = 2^(window_bits * w) * P_i for every
 * window w (affine, in HBM: 254/window_bits + 1 copies of the SRS).  MSMs against this SRS then
 * drop every scalar digit into ONE shared bucket set: no per-window bucket reduction and no
 * doublings at the end, and the window can be wider (fewer point additions).  window_bits = 0
 * picks it from the SRS length (16 at 2^16, 17 at 2^18, 20 for 2^20..2^24, 22 from 2^25).  Optional; results are identical.
 * Memory: (254/window_bits + 1) * n * 64 bytes on top of the SRS -- 3.25 GiB per basis at n = 2^22 (13 windows),
 * 48 GiB at n = 2^26 (12 windows).  When that allocation fails the call returns B2_ERR_OOM and changes nothing: the SRS
 * stays registered and MSMs against it run from the plain bases (one bucket set per window of at most 16 bits + Horner:
 * more point additions per scalar and a bucket reduction per window), so a caller may treat the error as "no table" and go on.  B2_ERR_ARG when
 * windows * n would not fit the 31-bit point index of the sort entries (n > 2^27 / windows). */
int b2_srs_precompute(b2_handle_t srs, uint32_t window_bits);
int b2_srs_len(b2_handle_t srs, size_t* n);
int b2_srs_read(b2_handle_t srs, size_t offset, size_t count, void* out_affine64);
int b2_srs_free(b2_handle_t srs);

/* ---- Params::read / Params::write (poly/commitment.rs:241-294) ---------------------- */
/* The params file holds g and g_lagrange as 32-byte compressed points (C::to_bytes / C::from_bytes of the pinned
 * pairing crate: x as 32 little-endian bytes of the canonical integer, the parity of the canonical y in bit
 * `sign_bit` of byte 31 -- 7 in the pasta / pairing_bn256 convention -- and the identity as 32 zero bytes).  The
 * reference decompresses on the CPU (:262-273: one Fq square root per point); these take the file bytes as they are.
 * An invalid encoding (x >= q, or x^3 + 3 not a square) fails with B2_ERR_ARG naming the point, where the reference
 * panics on `Option::from(C::from_bytes(..)).unwrap()` (:270). */
int b2_g1_decompress(const void* bytes32, size_t n, uint32_t sign_bit, void* out_affine64);
int b2_g1_compress(const void* affine64, size_t n, uint32_t sign_bit, void* out_bytes32);
/* file bytes -> resident SRS (the affine form never exists on the host) and back */
int b2_srs_register_compressed(const void* bytes32, size_t n, uint32_t sign_bit, b2_handle_t* out);
int b2_srs_read_compressed(b2_handle_t srs, size_t offset, size_t count, uint32_t sign_bit, void* out_bytes32);

/* ---- MSM -------------------------------------------------------------------------- */
/* sum_i scalars[i] * srs[offset + i], i < n.  max_bits bounds every scalar
 * (scalar < 2^max_bits); pass 254 (Fr::NUM_BITS) when unknown.  n == 0 or max_bits == 0
 * gives the identity.
 * Replaces gpu_multiexp_single_gpu_with_bound (arithmetic.rs:334-367) and, through
 * Params::commit / commit_lagrange / commit_lagrange_with_bound
 * (poly/commitment.rs:129-142,199-222), best_multiexp_gpu_cond (arithmetic.rs:442-458).
 * Zero scalars contribute nothing, so the CPU-side zero filtering of
 * commit_lagrange_with_bound (commitment.rs:204-212) is not needed. */
int b2_msm(b2_handle_t srs, size_t offset, const void* scalars, size_t n, uint32_t max_bits, void* out_jac96);
/* Asynchronous form of b2_msm for ONE caller thread that wants several columns in flight (the reference gets the same
 * overlap from several rayon workers, plonk/prover.rs:293-299): enqueues the H2D copy of the scalars, the MSM and the
 * read-back of the point on a free lane and returns a ticket at once; the copy of the next call then runs under the
 * kernels of this one.  `scalars` must stay valid and unchanged until b2_msm_wait(ticket) returns (pinned host memory
 * keeps the call non-blocking); b2_msm_wait blocks until the normalised point is in out_jac96 and returns the
 * call's status (B2_ERR_BOUND if a scalar exceeded max_bits).  Every ticket must be waited on exactly once; a call
 * blocks while all lanes (B2_LANES, default 3) are held by unfinished tickets. */
int b2_msm_async(b2_handle_t srs, size_t offset, const void* scalars, size_t n, uint32_t max_bits, void* out_jac96,
                 uint64_t* ticket);
int b2_msm_wait(uint64_t ticket);

/* Same with `scalars` and `out_jac96` in device memory, asynchronous on `stream`
 * (a cudaStream_t, NULL = the library's per-device stream).  The device result is a valid
 * Jacobian point but NOT normalised (the host entry points normalise after the read-back;
 * use b2_g1_normalize on the 96 bytes once they are on the host). */
int b2_msm_dev(b2_handle_t srs, size_t offset, const void* d_scalars, size_t n, uint32_t max_bits,
               void* d_out_jac96, void* stream);
/* best_multiexp (arithmetic.rs:465-492) with both slices on the host: uploads the bases
 * for this call only.  Prefer b2_srs_register + b2_msm. */
int b2_best_multiexp(const void* coeffs, const void* bases, size_t n, void* out_jac96);
/* Sum of `count` Jacobian points (96 B each): the combine of per-GPU partials that
 * gpu_multiexp_bound does on the host (arithmetic.rs:428-435). */
int b2_g1_sum(const void* jac96, size_t count, void* out_jac96);
/* Normalise `count` Jacobian points in host memory in place: Z = 1, or (0, 1, 0). */
int b2_g1_normalize(void* jac96, size_t count);
/* Same as b2_g1_sum with device pointers, asynchronous on `stream` (NULL = the library stream),
 * result not normalised: used after the NCCL all-gather of one partial per rank. */
int b2_g1_sum_dev(const void* d_jac96, size_t count, void* d_out_jac96, void* stream);
/* `groups` such sums in one launch: out[g] = sum over r < count of jac96[r * groups + g] -- the rank-major layout that ONE
 * all-gather of every rank's `groups` partials (a block of columns committed by point range) leaves behind. */
int b2_g1_sum_groups_dev(const void* d_jac96, size_t count, size_t groups, void* d_out_jac96, void* stream);

/* ---- NTT -------------------------------------------------------------------------- */
typedef struct b2_ntt_desc {
    uint32_t log_n;        /* transform length 2^log_n, 1 <= log_n <= 28 (Fr::S) */
    uint32_t location;     /* 0: host -> host; 1: device -> device; 2: host in, device out; 3: device in, host out */
    const void* omega;     /* 32 B, primitive 2^log_n-th root of unity */
    const void* divisor;   /* NULL, or 32 B multiplied into every output (iNTT 2^-k) */
    const void* coset_in;  /* NULL, or 64 B {z1, z2}: x[i] *= z_(i%3) for i%3 != 0 before the transform */
    const void* coset_out; /* NULL, or 64 B {z1, z2}: X[o] *= z_(o%3) for o%3 != 0 after it */
    uint64_t n_in;         /* elements present per input column (<= 2^log_n); the rest is zero */
    uint64_t n_out;        /* outputs kept per column (<= 2^log_n) */
    uint64_t columns;      /* independent transforms in this batch */
    const void* in;
    uint64_t in_stride;    /* elements between input columns */
    void* out;             /* may equal `in` */
    uint64_t out_stride;
    void* stream;          /* location 1 only: cudaStream_t (call is asynchronous on it) or NULL */
    const void* coset_gen; /* NULL, or 32 B g: x[i] *= g^i before the transform, i.e. the outputs are the evaluations
                            * of the input polynomial on the coset g * <omega>.  With g = zeta * extended_omega^c and
                            * omega of order 2^k this is the c-th of the 2^(extended_k - k) interleaved cosets that make
                            * up coeff_to_extended's output (row 2^(extended_k-k) * i + c), computed without zero padding */
} b2_ntt_desc;
/* General entry point; the functions below are thin wrappers over it. */
int b2_ntt_exec(const b2_ntt_desc* desc);

/* best_fft / gpu_fft (arithmetic.rs:495-512, 546-554): in-place natural-order NTT */
int b2_best_fft(void* a, const void* omega, uint32_t log_n);
/* gpu_ifft (arithmetic.rs:515-534) = EvaluationDomain::ifft (poly/domain.rs:400-414):
 * NTT with omega_inv, then every element times `divisor`.  lagrange_to_coeff[_st]
 * (poly/domain.rs:233-266) is this with (omega_inv, ifft_divisor, k). */
int b2_gpu_ifft(void* a, const void* omega_inv, uint32_t log_n, const void* divisor);
/* EvaluationDomain::coeff_to_extended (poly/domain.rs:270-287) for `columns` polynomials
 * of 2^k coefficients each (contiguous): coset scaling by {1, zeta, zeta^2}[i%3],
 * zero-extension to 2^ext_k, forward NTT with extended_omega.  `out` receives
 * columns * 2^ext_k elements. */
int b2_coeff_to_extended(const void* a, void* out, uint64_t columns, uint32_t k, uint32_t ext_k,
                         const void* zeta, const void* zeta_sq, const void* ext_omega);
/* EvaluationDomain::extended_to_coeff (poly/domain.rs:328-350): iNTT of size 2^ext_k,
 * scaling by {1, zeta^2, zeta}[i%3], truncation to n_out = n * (j - 1) elements. */
int b2_extended_to_coeff(const void* a, void* out, uint64_t n_out, uint32_t ext_k, const void* zeta,
                         const void* zeta_sq, const void* ext_omega_inv, const void* ext_divisor);
/* EvaluationDomain::divide_by_vanishing_poly (poly/domain.rs:354-373):
 * a[i] *= t_evaluations[i % t_len], t_len a power of two, in place (host pointers). */
int b2_divide_by_vanishing_poly(void* a, uint32_t ext_k, const void* t_evaluations, uint32_t t_len);

/* ---- fused commit + iNTT ---------------------------------------------------------- */
/* gpu_multiexp_bound_and_fft (arithmetic.rs:375-410) = Params::commit_lagrange_and_ifft
 * (poly/commitment.rs:144-170): one upload of the 2^log_n Lagrange values, MSM against
 * srs[0 .. 2^log_n) and in-place iNTT (omega = omega_inv, divisor = 2^-k) of the same
 * device-resident vector.  `coeffs` is overwritten with the coefficient form. */
int b2_msm_and_ifft(b2_handle_t srs, void* coeffs, uint32_t max_bits, const void* omega_inv,
                    const void* divisor, uint32_t log_n, void* out_jac96);
/* The prover's per-column batching (plonk/prover.rs:293-299, 470-501, 561-593): `columns`
 * contiguous columns of n scalars each, one commitment per column against the same SRS
 * (out: columns * 96 B); with do_ifft != 0 each column is also replaced by its iNTT. */
int b2_commit_batch(b2_handle_t srs, void* columns_data, uint64_t columns, size_t n, uint32_t max_bits,
                    int do_ifft, const void* omega_inv, const void* divisor, uint32_t log_n, void* out_jac96);

/* Same commitments, but the columns stay (or already are) in HBM: with columns_on_device == 0 each host column is
 * copied straight into d_columns[c * n] (pipelined over the lanes), committed there and, with do_ifft, transformed
 * in place there -- nothing is copied back; with columns_on_device != 0 the columns are read from d_columns and
 * columns_data is ignored.  This is what lets the advice / z columns be uploaded once per proof: their coefficient
 * forms are then consumed on the device by the coset transforms of evaluate_h (plonk/prover.rs:639-661). */
/* max_bits of b2_commit_batch[_resident]: B2_MAX_BITS_AUTO makes the bound of each column the bit length of its largest
 * scalar, found on the device right after the column has landed (find_max_scalar_bits + commit_lagrange_with_bound,
 * plonk/prover.rs:945-962, 293-299), still pipelined: the scan of column c waits for its copy only, its MSM runs under
 * the copy of column c + 1. */
#define B2_MAX_BITS_AUTO 0xFFFFFFFFu
int b2_commit_batch_resident(b2_handle_t srs, const void* columns_data, int columns_on_device, void* d_columns,
                             uint64_t columns, size_t n, uint32_t max_bits, int do_ifft, const void* omega_inv,
                             const void* divisor, uint32_t log_n, void* out_jac96);

/* ---- quotient evaluation (evaluate_h) ---------------------------------------------- */
/* Evaluator::evaluate_h (plonk/evaluation.rs:778-1226) computes, for every row of the extended
 * domain, the y-fold of all gate polynomials and of the permutation / lookup / shuffle terms.
 * The circuit-specific part is the reference's own data: `rotations`, `constants` and the list
 * of `Calculation`s over `ValueSource`s (evaluation.rs:46-112, built by Evaluator::new,
 * :309-620).  A program is that list in flat form; the engine lowers it once (dead-code
 * elimination, Store inlining, slot allocation from live ranges) and then evaluates it with ONE
 * kernel launch per call, every row a thread, columns device-resident. */
#define B2_Q_CONSTANT 0      /* ValueSource::Constant(index)                       (evaluation.rs:48) */
#define B2_Q_INTERMEDIATE 1  /* ValueSource::Intermediate(index): result of calcs[index]        (:50) */
#define B2_Q_FIXED 2         /* ValueSource::Fixed(index, rotation)                            (:52) */
#define B2_Q_ADVICE 3        /* ValueSource::Advice(index, rotation)                           (:54) */
#define B2_Q_INSTANCE 4      /* ValueSource::Instance(index, rotation)                         (:56) */
#define B2_Q_AUX 5           /* engine-side columns: z / sigma / m cosets, l0, l_last, l_active_row */
#define B2_Q_CHALLENGE 6     /* challenges[index] (beta, gamma, theta, y, beta*zeta*delta^j, ...) */
#define B2_Q_COSET_X 7       /* x0 * x_step^row: `beta_term` of evaluation.rs:1018-1019 */

#define B2_QOP_ADD 0            /* Calculation::Add(a, b)                                      (:97) */
#define B2_QOP_SUB 1            /* Calculation::Sub(a, b)                                      (:99) */
#define B2_QOP_MUL 2            /* Calculation::Mul(a, b)                                     (:101) */
#define B2_QOP_NEGATE 3         /* Calculation::Negate(a)                                     (:103) */
#define B2_QOP_LC_CHALLENGE 4   /* Calculation::LcChallenge(a, b, ch, p) = (a + ch^p) * b, ch^1 when p <= 1 (:105,208-211) */
#define B2_QOP_MUL_CH_ADD 5     /* a * ch + b: Calculation::LcTheta (ch = theta, :107) and the fold value * y + part (:897) */
#define B2_QOP_ADD_CHALLENGE 6  /* Calculation::AddChallenge(a, ch)                           (:109) */
#define B2_QOP_STORE 7          /* Calculation::Store(a)                                      (:111) */

typedef struct b2_qsrc {
    uint32_t kind;      /* B2_Q_* */
    uint32_t index;
    uint32_t rotation;  /* column kinds: index into `rotations` */
} b2_qsrc;
typedef struct b2_qcalc {
    uint32_t op;        /* B2_QOP_* */
    b2_qsrc a, b;
    uint32_t challenge; /* LC_CHALLENGE / MUL_CH_ADD / ADD_CHALLENGE: index into challenges */
    uint32_t power;     /* LC_CHALLENGE */
} b2_qcalc;
typedef struct b2_quotient_program_desc {
    const int32_t* rotations;   /* Evaluator::rotations (:274) */
    uint32_t n_rotations;
    const void* constants;      /* Evaluator::constants (:272), n_constants * 32 B, Montgomery */
    uint32_t n_constants;
    const b2_qcalc* calcs;      /* Evaluator::calculations (:276) followed by whatever the caller appends */
    uint32_t n_calcs;
    b2_qsrc result;             /* the value stored per row */
    uint32_t n_fixed, n_advice, n_instance, n_aux, n_challenges;   /* table sizes (validated) */
} b2_quotient_program_desc;
/* Lowers and validates on the host (works without a GPU); device upload happens on first use. */
int b2_quotient_program_create(const b2_quotient_program_desc* desc, b2_handle_t* out);
int b2_quotient_program_free(b2_handle_t program);
/* lowered instruction count, shared-memory slots per row, field multiplications / additions per row */
int b2_quotient_program_info(b2_handle_t program, uint32_t* n_instr, uint32_t* n_slots, uint32_t* n_mul,
                             uint32_t* n_addsub);
/* The two slot classes of the lowered program: n_slots = n_shared + n_global.  A program of 10 or more live values
 * (shared-memory slots of 4 KB per CTA: from there on the lost resident CTAs cost more than global round trips, measured)
 * keeps 7 slots in shared memory and its longest-lived values -- sub-expressions the circuit shares between gates far
 * apart in its gate list (evaluation.rs:877-907 keeps every such value for the whole row) -- in a per-CTA global
 * scratch; n_global = 0 for every other program.  B2_Q_HYBRID=0 in the environment
 * puts every slot into shared memory (A/B runs). */
int b2_quotient_program_slot_classes(b2_handle_t program, uint32_t* n_shared, uint32_t* n_global);

/* Diagnostic: the lowered program (4 words per instruction: op | dst_slot << 8, operand a, operand b, operand c;
 * operand word = kind << 28 | rotation << 20 | index with kind 0 constant, 1 slot, 2 column (fixed, advice,
 * instance, aux concatenated), 3 challenge (derived powers appended after the caller's), 4 coset x; ops
 * 0 add, 1 sub, 2 mul, 3 neg, 4 copy, 5 / 6 the fused a * b +- c * d whose fourth operand d is word a of the
 * following entry (op 7, never executed by itself)) and the (challenge, power) pairs of the derived challenge entries.
 * Lets the host-side lowering be checked without a GPU. */
int b2_quotient_program_dump(b2_handle_t program, uint32_t* instr_words, size_t instr_capacity, uint32_t* result_word,
                             uint32_t* derived_pairs, size_t derived_capacity, uint32_t* n_derived);

typedef struct b2_quotient_args {
    uint32_t log_rows;            /* rows = 2^log_rows: extended_k (evaluation.rs:792), or k for one coset */
    uint32_t rot_scale;           /* 2^(extended_k - k) (:793), or 1 */
    const void* const* fixed;     /* host arrays of DEVICE pointers, one per column, 2^log_rows Fr each */
    const void* const* advice;
    const void* const* instance;
    const void* const* aux;
    const void* challenges;       /* host, n_challenges * 32 B */
    const void* x0;               /* host 32 B each; NULL when the program has no B2_Q_COSET_X operand */
    const void* x_step;
    const void* scale;            /* host, NULL or scale_len * 32 B: result *= scale[row % scale_len] */
    uint32_t scale_len;           /*   (divide_by_vanishing_poly's t_evaluations, poly/domain.rs:354-373) */
    void* out;                    /* device: result of row i is stored at out[out_offset + i * out_stride] */
    uint64_t out_stride;
    uint64_t out_offset;
    void* stream;                 /* cudaStream_t or NULL; the call returns after the launch is enqueued */
    uint64_t row_begin;           /* row_count != 0: only rows [row_begin, row_begin + row_count) are evaluated and   */
    uint64_t row_count;           /*   row i is stored at out[out_offset + (i - row_begin) * out_stride] (a rank's share) */
} b2_quotient_args;
int b2_quotient_eval(b2_handle_t program, const b2_quotient_args* args);

/* ---- grand products / grand sums ----------------------------------------------------- */
/* The vector steps between the expression kernel and commit_lagrange_and_ifft when the prover builds
 * the permutation, logup and shuffle z columns (plonk/permutation/prover.rs:72-165,
 * plonk/logup/prover.rs:263-336, plonk/shuffle/prover.rs:107-141).  *_dev variants take device pointers
 * and are asynchronous on `stream` when it is not NULL. */
/* batch_invert (arithmetic.rs:840-844; ff::BatchInvert semantics: zeros are left as zero), in place */
int b2_batch_invert(void* a, size_t n);
int b2_batch_invert_dev(void* d_a, size_t n, void* stream);
/* out[0] = init, out[i + 1] = init (op) in[0] (op) ... (op) in[i] for i + 1 < n_out <= n_in + 1.
 * op 0: product (mul_acc, arithmetic.rs:806-836; permutation/prover.rs:149-152; shuffle/prover.rs:137-141),
 * op 1: sum (logup/prover.rs:318-336).  init: 32 B on the host, NULL = the operator's identity; d_init (device,
 * 32 B) overrides it -- the last_z hand-over between column sets (permutation/prover.rs:146,160) without a
 * host round trip.  out must not alias in. */
int b2_prefix_scan(int op, const void* in, size_t n_in, const void* init, void* out, size_t n_out);
int b2_prefix_scan_dev(int op, const void* d_in, size_t n_in, const void* init, const void* d_init, void* d_out,
                       size_t n_out, void* stream);
/* d_out[i] = d_a[i] op d_b[i] over Fr on device pointers; op: 0 mul, 1 add, 2 sub.  d_out may alias an input. */
int b2_fr_vec_dev(int op, const void* d_a, const void* d_b, size_t n, void* d_out, void* stream);
/* Bit length of the largest canonical value among n resident Montgomery-form scalars: find_max_scalar_bits
 * (plonk/prover.rs:945-962), the bound the reference passes to commit_lagrange_with_bound for every advice column. */
int b2_fr_max_bits_dev(const void* d_a, size_t n, uint32_t* bits);
/* The vanishing argument's random polynomial (plonk/vanishing/prover.rs:48-63): d_out[i] = (a_i + random[u_i % k]) *
 * (b_i + random[v_i % k]) for i < n, with a_i, u_i, b_i, v_i from a counter-based generator keyed by the 32 bytes at
 * `key32` (the reference draws them from thread_rng, i.e. ChaCha under a 256-bit key; here: the ChaCha20 key stream
 * of RFC 8439 under that key, nonce 0 -- blocks 3i and 3i + 1 reduced mod r as 512-bit little-endian integers are a_i
 * and b_i, the first two 64-bit words of block 3i + 2 are u_i and v_i; csrc/scan.cuh + csrc/chacha.cuh state the
 * generator, oracle/prover.py restates it).  random: k Montgomery field elements on the HOST (the reference's `random`
 * vector, k = domain.k() <= 64).  n <= 2^30.  Synchronous. */
int b2_vanishing_random_poly_dev(const void* key32, const void* random, uint32_t k, size_t n, void* d_out, void* stream);

/* ---- polynomial evaluation / division on resident coefficient forms --------------------- */
/* eval_polynomial (arithmetic.rs:707-735): out[c] = sum_i poly_c[i] * point^i for `columns` polynomials of n
 * coefficients (device, `stride` elements apart) at one point; the evaluations land on the host (columns * 32 B).
 * This is what lets the coefficient forms stay in HBM after evaluate_h: the prover's evaluation phase
 * (plonk/prover.rs:693-760) needs numbers, not polynomials. */
int b2_eval_polynomial_dev(const void* d_polys, uint64_t columns, uint64_t stride, uint64_t n, const void* point,
                           void* out_host);
/* The same for `count` polynomials that do not sit in one strided block: d_poly_ptrs is a HOST array of device
 * pointers (n coefficients each); one launch evaluates them all at `point` -- the evaluation phase of create_proof
 * (plonk/prover.rs:693-790) makes one call per distinct point instead of one per query. */
int b2_eval_polynomials_dev(const void* const* d_poly_ptrs, uint64_t count, uint64_t n, const void* point, void* out_host);

int b2_eval_polynomial(const void* poly, uint64_t n, const void* point, void* out);
/* kate_division (arithmetic.rs:752-773): q = (a(X) - a(b)) / (X - b), n - 1 coefficients; q must not alias a.
 * Used by the multiopen provers (poly/multiopen/gwc/prover.rs:158, shplonk/prover.rs:23). */
int b2_kate_division_dev(const void* d_a, uint64_t n, const void* b, void* d_q, void* stream);
int b2_kate_division(const void* a, uint64_t n, const void* b, void* q);

/* ---- witness file ------------------------------------------------------------------------ */
/* The on-disk witness of halo2_proofs/src/helpers.rs:919-1015 (store_witness / fetch_witness): a u32 LE column count,
 * then advice column i at byte offset 4 + i * 2^(k+5), 2^k elements as they sit in memory.  b2_commit_witness_file
 * replaces fetch_witness + the per-column commit_lagrange_with_bound loop (plonk/prover.rs:293-299) for columns
 * [first, first + count): the file is read through two pinned staging buffers while the previous group is being
 * copied and committed; with d_keep != NULL the columns also stay resident there (count * 2^k * 32 B, Lagrange form).
 * out_jac96: count points, normalised. */
int b2_witness_file_columns(const char* path, uint32_t* n_columns);
int b2_commit_witness_file(b2_handle_t srs, const char* path, uint32_t k, uint64_t first, uint64_t count, uint32_t max_bits,
                           void* d_keep, void* out_jac96);

/* Multiopen batching (poly/multiopen/gwc/prover.rs:47-56 and its cuda variant :58-140, which runs the same fold on
 * the GPU through the Bn256_Fr_eval_mul_c / _eval_sum kernels; shplonk/prover.rs folds the same way):
 * out[i] = sum_j v^(m-1-j) * polys[j][i], i.e. poly_batch = poly_batch * v + poly over the m polynomials opened at
 * one point.  `polys` is a HOST array of m pointers (device pointers for _dev, host pointers otherwise), n
 * coefficients each; v: 32 B Montgomery on the host.  d_out may alias polys[0] only. */
int b2_poly_combine_dev(const void* const* d_polys, uint32_t m, uint64_t n, const void* v, void* d_out, void* stream);
int b2_poly_combine(const void* const* polys, uint32_t m, uint64_t n, const void* v, void* out);

/* ---- memory helpers --------------------------------------------------------------- */
int b2_host_alloc(size_t bytes, void** out);   /* page-locked host memory */
int b2_host_free(void* p);
/* Page-lock an existing host allocation in place (e.g. the Vec<Fr> of a long-lived advice column)
 * so that its copies run at full PCIe rate and overlap with kernels; undo with b2_host_unregister. */
int b2_host_register(void* p, size_t bytes);
int b2_host_unregister(void* p);
int b2_dev_alloc(size_t bytes, void** out);
int b2_dev_free(void* p);
int b2_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes);
int b2_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes);
int b2_memcpy_d2d(void* dst_dev, const void* src_dev, size_t bytes);

/* ---- logup multiplicities (plonk/logup/prover.rs:117-179) on device-resident columns ------------------------------
 * d_inputs: n_inputs compressed input columns of n Fr each (Montgomery form, column stride n); d_table: the compressed
 * table column (n Fr).  Only rows < usable take part.  The table's usable rows are sorted stably by value on the
 * device (radix sort of the canonical 256-bit values), every input value is located with the probe sequence of the
 * pinned toolchain's binary_search_by_key (so that a repeated table value credits the same row as the reference) and
 * d_m receives the multiplicities as n Montgomery-form Fr (zeros in rows >= usable).  *largest_count = the largest
 * multiplicity (the bound for commit_lagrange_with_bound, :206-212).  B2_ERR_ARG when an input value is not in the
 * table ("logup binary_search_by_key should hit", :148).  Synchronous; n < 2^31. */
int b2_logup_multiplicity_dev(const void* d_inputs, uint32_t n_inputs, const void* d_table, uint64_t usable, uint64_t n,
                              void* d_m, uint64_t* largest_count);

/* ---- diagnostics used by tests / bench -------------------------------------------- */
/* out[i] = a[i] op b[i] on the device.  field: 0 Fr, 1 Fq.  op: 0 mul, 1 add, 2 sub, 3 sqr(a),
 * 4 (Fr only) the same product as op 0 computed through the NTT's Shoup constant-multiplication path,
 * 5 a*b + b*b and 6 a*b - b*b through the fused two-product reduction the MSM point additions use. */
int b2_field_vec(int field, int op, const void* a, const void* b, size_t n, void* out);
/* Measures the device's sustained 32x32->64 multiply-accumulate rate with the kernels'
 * own instruction mix (carry-chained IMAD.WIDE Montgomery products); returns modular
 * multiplications per second and wide MACs/s counted as 128 per product (64 product + 64
 * reduction terms, the accounting of SURVEY.md section 8d). */
int b2_imad_probe(double* wide_macs_per_s, double* modmuls_per_s);
/* Same probe for the constant multiplication the NTT butterflies use (Shoup: 92 wide MACs + 23 narrow
 * products per multiplication by a precomputed twiddle instead of 128 + 8). */
int b2_shoup_probe(double* muls_per_s);
/* Rates of the other field products, same harness: kind 0 = interleaved Montgomery product (= b2_imad_probe),
 * 1 = Shoup constant product, 4 = two products under one reduction (fp_mul2_add, 192 MACs; counted as ONE call);
 * 2 = Karatsuba Montgomery product (112 MACs) and 3 = Montgomery squaring (100 MACs) exist only in builds with
 * -DB2_FP_GEN (generated variants that measured slower, see csrc/fp.cuh) and return B2_ERR_ARG otherwise. */
int b2_mul_probe(int kind, double* per_s);
/* Raw multiplier-pipe rates, independent of the field code (8 independent accumulator chains of one instruction kind
 * per thread, no carries between instructions, no loads): per-thread multiply-accumulates per second of
 *   kind 0: IMAD.WIDE.U32 Rd, Ra, b, Rd (32 x 32 + 64 -> 64; every limb product of the bignum kernels),
 *   kind 1: IMAD (32 x 32 + 32 -> 32; the "64 results per clock per SM" instruction of the CUDA programming guide).
 * The roofline of the integer kernels is reported against kind 0 as well as against b2_imad_probe. */
int b2_pipe_probe(int kind, double* macs_per_s);
/* FP64 FMA rate of the device (diagnostic: documents why the fp64 pipe is / is not a usable
 * second multiplier for the bignum kernels on this part). */
int b2_dfma_probe(double* dfma_per_s);
/* Two-pipe probe (diagnostic): blocks of 8 warps, 2 blocks per SM; warps whose bit is set in imad_mask run
 * `iters_int` x 2 Montgomery products (int_kind 0) or the same number of raw IMAD.WIDE (int_kind 1), warps in dfma_mask run
 * `iters_f64` x 64 DFMA, the rest exit.  Returns the best kernel time of three launches in ms.  Comparing (mask_i, 0),
 * (0, mask_d) and (mask_i, mask_d) shows whether the fp64 pipe runs next to the wide-integer multiplier. */
int b2_mixed_probe(uint32_t imad_mask, uint32_t dfma_mask, int int_kind, int iters_int, int iters_f64, double* ms_out);
/* Batched-affine addition probe (diagnostic): n_pairs additions P[i] + Q[i] of synthetic affine points, B
 * consecutive pairs per thread with one shared inversion (Montgomery's trick, prefix products parked in global memory),
 * against the same points through the XYZZ mixed add the MSM uses (two mixed adds per pair into a running bucket).
 * Times are kernel ms (best of three).  The first n_copy points P, Q and sums (64-byte affine) are copied to the host
 * buffers when those are not NULL, so that a test can check the sums. */
int b2_affine_batch_probe(size_t n_pairs, uint32_t B, double* batch_ms, double* xyzz_ms, void* host_p, void* host_q,
                          void* host_out, size_t n_copy);
/* Timing of the last b2_msm / b2_ntt_exec / b2_commit_batch on this device, CUDA events
 * on the launching stream: kernel-only ms and (host variants) total ms incl. copies. */
int b2_last_timing(double* kernel_ms, double* total_ms);
/* Per-phase kernel times of the last MSM (ms): digits, scan, scatter, accumulate, fixup,
 * reduce, final.  `phases` must hold 8 doubles. */
int b2_last_msm_phases(double* phases);
/* Window configuration an MSM of n scalars against `srs` (0 = plain bases) would use:
 * window bits c, number of windows, number of bucket sets (1 with a window table). */
int b2_msm_config(b2_handle_t srs, size_t n, uint32_t max_bits, uint32_t* c, uint32_t* windows,
                  uint32_t* bucket_sets);

#ifdef __cplusplus
}
#endif
#endif /* B2PCS_H */
