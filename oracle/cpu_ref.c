/*
 * oracle/cpu_ref.c -- C restatement of the reference's CPU hot path.
 *
 * TEST INFRASTRUCTURE ONLY (checker + timed CPU baseline).  Product code never
 * links or loads this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * PARITY STATUS: byte-level parity unpinned (see oracle/bn254.py header): the
 * reference is Rust with un-vendored arithmetic (pairing_bn256 @30b052f2) and no
 * golden vectors; this file restates the reference's algorithms on the published
 * BN254 parameters and is validated bit-for-bit against oracle/bn254.py, which is
 * itself pinned on independent known-answer vectors (tests/test_oracle.py).
 *
 * What is restated (reference file:line under halo2_proofs/src):
 *   multiexp_serial        arithmetic.rs:20-108   (Pippenger, unsigned c-bit digits,
 *                                                  None/Affine/Projective buckets)
 *   best_multiexp          arithmetic.rs:465-492  (chunk = n / T, ordered fold)
 *   best_fft_cpu           arithmetic.rs:556-645  (bit-reverse, n/2 twiddles, DIT)
 *   EvaluationDomain::{ifft, distribute_powers_zeta, coeff_to_extended,
 *                      extended_to_coeff}         poly/domain.rs:270-414
 *
 * Element layout = the C ABI's (include/b2pcs.h): Fr/Fq 4 x u64 LE Montgomery;
 * affine 64 B (identity = (0,0)); Jacobian 96 B (identity Z = 0).
 *
 * Build: make -C oracle   (gcc -O3 -march=native -pthread -shared)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

typedef struct { u64 l[4]; } fe;              /* field element, Montgomery */
typedef struct { fe x, y; } aff;              /* affine point, (0,0) = identity */
typedef struct { fe x, y, z; } jac;           /* Jacobian point, z = 0 identity */

typedef struct {
    u64 p[4];
    u64 inv;     /* -p^{-1} mod 2^64 */
    fe one;      /* R mod p */
    fe r2;       /* R^2 mod p */
} field_t;

static const field_t FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
    {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}},
};
static const field_t FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
    {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}},
};

/* ------------------------------------------------------------------ field */
static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int geq_p(const u64 a[4], const u64 p[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > p[i]) return 1;
        if (a[i] < p[i]) return 0;
    }
    return 1;
}
static inline void sub_p(u64 a[4], const u64 p[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - p[i] - (u64)br;
        a[i] = (u64)d;
        br = (d >> 64) & 1;
    }
}
static inline void fe_add(const field_t *F, fe *r, const fe *a, const fe *b) {
    u128 c = 0;
    u64 t[4];
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; t[i] = (u64)c; c >>= 64; }
    if (geq_p(t, F->p)) sub_p(t, F->p);          /* p < 2^254: no carry out */
    memcpy(r->l, t, 32);
}
static inline void fe_sub(const field_t *F, fe *r, const fe *a, const fe *b) {
    u64 t[4];
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - (u64)br;
        t[i] = (u64)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) { c += (u128)t[i] + F->p[i]; t[i] = (u64)c; c >>= 64; }
    }
    memcpy(r->l, t, 32);
}
static inline void fe_neg(const field_t *F, fe *r, const fe *a) {
    fe z = {{0, 0, 0, 0}};
    fe_sub(F, r, &z, a);
}
static inline void fe_dbl(const field_t *F, fe *r, const fe *a) { fe_add(F, r, a, a); }

/* Montgomery multiplication (CIOS, 4 x 64-bit limbs) */
static inline void fe_mul(const field_t *F, fe *r, const fe *a, const fe *b) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (u64)c;
        t[5] = (u64)(c >> 64);
        u64 m = t[0] * F->inv;
        c = (u128)m * F->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (u64)c;
        t[4] = t[5] + (u64)(c >> 64);
    }
    if (t[4] || geq_p(t, F->p)) sub_p(t, F->p);
    memcpy(r->l, t, 32);
}
static inline void fe_sqr(const field_t *F, fe *r, const fe *a) { fe_mul(F, r, a, a); }
static void fe_from_mont(const field_t *F, u64 out[4], const fe *a) {
    fe one = {{1, 0, 0, 0}}, r;
    fe_mul(F, &r, a, &one);
    memcpy(out, r.l, 32);
}
static void fe_pow(const field_t *F, fe *r, const fe *a, const u64 e[4]) {
    fe acc = F->one;
    for (int i = 255; i >= 0; i--) {
        fe_sqr(F, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) fe_mul(F, &acc, &acc, a);
    }
    *r = acc;
}
static void fe_inv(const field_t *F, fe *r, const fe *a) {
    u64 e[4];
    memcpy(e, F->p, 32);
    e[0] -= 2; /* p is odd and p[0] >= 2 */
    fe_pow(F, r, a, e);
}

/* ------------------------------------------------------------------ curve */
#define Q (&FQ)
static inline int jac_is_id(const jac *p) { return fe_is_zero(&p->z); }
static inline int aff_is_id(const aff *p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void jac_set_id(jac *p) { memset(p, 0, sizeof *p); p->y = FQ.one; }

static void jac_double(jac *r, const jac *p) {           /* dbl-2009-l, a = 0 */
    if (jac_is_id(p)) { *r = *p; return; }
    fe a, b, c, d, e, f, t;
    fe_sqr(Q, &a, &p->x);
    fe_sqr(Q, &b, &p->y);
    fe_sqr(Q, &c, &b);
    fe_add(Q, &t, &p->x, &b); fe_sqr(Q, &t, &t); fe_sub(Q, &t, &t, &a); fe_sub(Q, &t, &t, &c);
    fe_dbl(Q, &d, &t);
    fe_dbl(Q, &e, &a); fe_add(Q, &e, &e, &a);
    fe_sqr(Q, &f, &e);
    fe z3; fe_mul(Q, &z3, &p->y, &p->z); fe_dbl(Q, &z3, &z3);
    fe x3; fe_dbl(Q, &t, &d); fe_sub(Q, &x3, &f, &t);
    fe y3; fe_sub(Q, &t, &d, &x3); fe_mul(Q, &y3, &e, &t);
    fe c8; fe_dbl(Q, &c8, &c); fe_dbl(Q, &c8, &c8); fe_dbl(Q, &c8, &c8);
    fe_sub(Q, &y3, &y3, &c8);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add_mixed(jac *r, const jac *p, const aff *q) { /* complete madd */
    if (aff_is_id(q)) { *r = *p; return; }
    if (jac_is_id(p)) { r->x = q->x; r->y = q->y; r->z = FQ.one; return; }
    fe z1z1, u2, s2, h, rr, t;
    fe_sqr(Q, &z1z1, &p->z);
    fe_mul(Q, &u2, &q->x, &z1z1);
    fe_mul(Q, &s2, &q->y, &p->z); fe_mul(Q, &s2, &s2, &z1z1);
    fe_sub(Q, &h, &u2, &p->x);
    fe_sub(Q, &rr, &s2, &p->y);
    if (fe_is_zero(&h)) {
        if (fe_is_zero(&rr)) { jac_double(r, p); return; }
        jac_set_id(r); return;
    }
    fe hh, hhh, v;
    fe_sqr(Q, &hh, &h);
    fe_mul(Q, &hhh, &hh, &h);
    fe_mul(Q, &v, &p->x, &hh);
    fe x3, y3, z3;
    fe_sqr(Q, &x3, &rr); fe_sub(Q, &x3, &x3, &hhh); fe_dbl(Q, &t, &v); fe_sub(Q, &x3, &x3, &t);
    fe_sub(Q, &t, &v, &x3); fe_mul(Q, &y3, &rr, &t); fe_mul(Q, &t, &p->y, &hhh); fe_sub(Q, &y3, &y3, &t);
    fe_mul(Q, &z3, &p->z, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(jac *r, const jac *p, const jac *q) {      /* complete add */
    if (jac_is_id(q)) { *r = *p; return; }
    if (jac_is_id(p)) { *r = *q; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, rr, t;
    fe_sqr(Q, &z1z1, &p->z);
    fe_sqr(Q, &z2z2, &q->z);
    fe_mul(Q, &u1, &p->x, &z2z2);
    fe_mul(Q, &u2, &q->x, &z1z1);
    fe_mul(Q, &s1, &p->y, &q->z); fe_mul(Q, &s1, &s1, &z2z2);
    fe_mul(Q, &s2, &q->y, &p->z); fe_mul(Q, &s2, &s2, &z1z1);
    fe_sub(Q, &h, &u2, &u1);
    fe_sub(Q, &rr, &s2, &s1);
    if (fe_is_zero(&h)) {
        if (fe_is_zero(&rr)) { jac_double(r, p); return; }
        jac_set_id(r); return;
    }
    fe hh, hhh, v;
    fe_sqr(Q, &hh, &h);
    fe_mul(Q, &hhh, &hh, &h);
    fe_mul(Q, &v, &u1, &hh);
    fe x3, y3, z3;
    fe_sqr(Q, &x3, &rr); fe_sub(Q, &x3, &x3, &hhh); fe_dbl(Q, &t, &v); fe_sub(Q, &x3, &x3, &t);
    fe_sub(Q, &t, &v, &x3); fe_mul(Q, &y3, &rr, &t); fe_mul(Q, &t, &s1, &hhh); fe_sub(Q, &y3, &y3, &t);
    fe_mul(Q, &z3, &p->z, &q->z); fe_mul(Q, &z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_to_affine(aff *r, const jac *p) {
    if (jac_is_id(p)) { memset(r, 0, sizeof *r); return; }
    fe zi, zi2, zi3;
    fe_inv(Q, &zi, &p->z);
    fe_sqr(Q, &zi2, &zi);
    fe_mul(Q, &zi3, &zi2, &zi);
    fe_mul(Q, &r->x, &p->x, &zi2);
    fe_mul(Q, &r->y, &p->y, &zi3);
}
static void jac_mul_u256(jac *r, const aff *p, const u64 k[4]) {
    jac acc; jac_set_id(&acc);
    for (int i = 255; i >= 0; i--) {
        jac_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) jac_add_mixed(&acc, &acc, p);
    }
    *r = acc;
}

/* ------------------------------------------------- multiexp_serial :20-108 */
typedef struct { int kind; /* 0 None, 1 Affine, 2 Projective */ aff a; jac j; } bucket_t;

static inline size_t get_at(size_t segment, size_t c, const uint8_t bytes[32]) { /* :31-49 */
    size_t skip_bits = segment * c;
    size_t skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t avail = 32 - skip_bytes;
    memcpy(v, bytes + skip_bytes, avail < 8 ? avail : 8);
    u64 tmp;
    memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    tmp = tmp % ((u64)1 << c);
    return (size_t)tmp;
}

static void multiexp_serial(const fe *coeffs, const aff *bases, size_t m, jac *acc) {
    uint8_t (*reprs)[32] = malloc(m ? m * 32 : 32);
    for (size_t i = 0; i < m; i++) fe_from_mont(&FR, (u64 *)reprs[i], &coeffs[i]); /* to_repr :21 */
    size_t c;
    if (m < 4) c = 1; else if (m < 32) c = 3; else c = (size_t)ceil(log((double)(uint32_t)m)); /* :23-29 */
    size_t segments = 256 / c + 1;                                                  /* :51 */
    size_t nb = ((size_t)1 << c) - 1;
    bucket_t *buckets = malloc(nb * sizeof(bucket_t));
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t i = 0; i < c; i++) jac_double(acc, acc);                         /* :54-56 */
        for (size_t i = 0; i < nb; i++) buckets[i].kind = 0;                         /* :89 */
        for (size_t i = 0; i < m; i++) {                                             /* :91-96 */
            size_t d = get_at(seg, c, reprs[i]);
            if (d != 0) {
                bucket_t *b = &buckets[d - 1];
                if (b->kind == 0) { b->kind = 1; b->a = bases[i]; }
                else if (b->kind == 1) {
                    jac t; t.x = b->a.x; t.y = b->a.y; t.z = FQ.one;
                    if (aff_is_id(&b->a)) jac_set_id(&t);
                    jac_add_mixed(&b->j, &t, &bases[i]);
                    b->kind = 2;
                } else jac_add_mixed(&b->j, &b->j, &bases[i]);
            }
        }
        jac running; jac_set_id(&running);                                           /* :102-106 */
        for (size_t i = nb; i-- > 0;) {
            bucket_t *b = &buckets[i];
            if (b->kind == 1) jac_add_mixed(&running, &running, &b->a);
            else if (b->kind == 2) jac_add(&running, &running, &b->j);
            jac_add(acc, acc, &running);
        }
    }
    free(buckets);
    free(reprs);
}

typedef struct { const fe *c; const aff *b; size_t m; jac acc; } msm_job;
static void *msm_worker(void *arg) {
    msm_job *j = arg;
    jac_set_id(&j->acc);
    multiexp_serial(j->c, j->b, j->m, &j->acc);
    return NULL;
}

/* best_multiexp :465-492.  T = "rayon threads".  Returns 0, or -1 on bad args. */
int ref_best_multiexp(const u64 *coeffs, const u64 *bases, size_t n, int T, u64 *out_jac) {
    const fe *cs = (const fe *)coeffs;
    const aff *bs = (const aff *)bases;
    jac acc; jac_set_id(&acc);
    if (T < 1) T = 1;
    if (n > (size_t)T) {
        size_t chunk = n / (size_t)T;
        size_t nchunks = (n + chunk - 1) / chunk;
        msm_job *jobs = malloc(nchunks * sizeof *jobs);
        pthread_t *th = malloc(nchunks * sizeof *th);
        for (size_t i = 0; i < nchunks; i++) {
            size_t lo = i * chunk, hi = lo + chunk > n ? n : lo + chunk;
            jobs[i].c = cs + lo; jobs[i].b = bs + lo; jobs[i].m = hi - lo;
            pthread_create(&th[i], NULL, msm_worker, &jobs[i]);
        }
        for (size_t i = 0; i < nchunks; i++) pthread_join(th[i], NULL);
        for (size_t i = 0; i < nchunks; i++) jac_add(&acc, &acc, &jobs[i].acc);      /* :486 */
        free(jobs); free(th);
    } else {
        multiexp_serial(cs, bs, n, &acc);
    }
    memcpy(out_jac, &acc, sizeof acc);
    return 0;
}

/* ------------------------------------------------------ best_fft_cpu :556 */
typedef struct { fe *a; const fe *tw; size_t n, lo, hi, half, tw_chunk; } fft_job;

static inline void butterfly(fe *a, fe *b, const fe *tw) {   /* :632-636 */
    fe t;
    fe_mul(&FR, &t, b, tw);
    fe u = *a;
    fe_add(&FR, a, &u, &t);
    fe_sub(&FR, b, &u, &t);
}
/* all butterflies with global butterfly index in [lo, hi) of the stage with
 * half-size `half`: butterfly g -> block g / half, offset g % half */
static void *fft_stage_worker(void *arg) {
    fft_job *j = arg;
    for (size_t g = j->lo; g < j->hi; g++) {
        size_t blk = g / j->half, i = g % j->half;
        fe *lo = j->a + blk * 2 * j->half + i;
        butterfly(lo, lo + j->half, &j->tw[i * j->tw_chunk]);
    }
    return NULL;
}
typedef struct { fe *tw; const fe *prev; fe base; size_t lo, hi; } tw_job;
static void *tw_worker(void *arg) {
    tw_job *j = arg;
    for (size_t i = j->lo; i < j->hi; i++) fe_mul(&FR, &j->tw[i], &j->base, &j->prev[i]);
    return NULL;
}
static size_t bitrev(size_t n, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}

int ref_best_fft(u64 *a_, const u64 *omega_, uint32_t log_n, int T) {
    fe *a = (fe *)a_;
    fe omega; memcpy(&omega, omega_, 32);
    size_t n = (size_t)1 << log_n;
    if (T < 1) T = 1;
    for (size_t k = 0; k < n; k++) {                                   /* :571-576 */
        size_t rk = bitrev(k, log_n);
        if (k < rk) { fe t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    size_t half_n = n / 2 ? n / 2 : 1;
    fe *tw = malloc(half_n * sizeof(fe));                              /* :580-611 */
    tw[0] = FR.one;
    size_t chunk_size = (size_t)1 << 14;
    pthread_t *th = malloc((size_t)T * sizeof *th);
    if (n / 2 < chunk_size) {
        for (size_t i = 1; i < n / 2; i++) fe_mul(&FR, &tw[i], &tw[i - 1], &omega);
    } else {
        for (size_t i = 1; i < chunk_size; i++) fe_mul(&FR, &tw[i], &tw[i - 1], &omega);
        fe base; fe_mul(&FR, &base, &tw[chunk_size - 1], &omega);
        tw_job *jobs = malloc((size_t)T * sizeof *jobs);
        for (size_t c0 = chunk_size; c0 < n / 2; c0 += chunk_size) {
            size_t per = (chunk_size + (size_t)T - 1) / (size_t)T;
            int used = 0;
            for (int t = 0; t < T; t++) {
                size_t lo = (size_t)t * per, hi = lo + per > chunk_size ? chunk_size : lo + per;
                if (lo >= hi) break;
                jobs[t] = (tw_job){tw + c0, tw + c0 - chunk_size, base, lo, hi};
                pthread_create(&th[t], NULL, tw_worker, &jobs[t]);
                used++;
            }
            for (int t = 0; t < used; t++) pthread_join(th[t], NULL);
        }
        free(jobs);
    }
    /* butterflies :613-641 (the recursive rayon variant :647-705 computes the same
     * values; field elements are canonical so the bits are identical) */
    fft_job *jobs = malloc((size_t)T * sizeof *jobs);
    size_t half = 1, tw_chunk = n / 2;
    for (uint32_t s = 0; s < log_n; s++) {
        size_t total = n / 2;
        size_t per = (total + (size_t)T - 1) / (size_t)T;
        if (total < 4096 || T == 1) {
            fft_job j = {a, tw, n, 0, total, half, tw_chunk};
            fft_stage_worker(&j);
        } else {
            int used = 0;
            for (int t = 0; t < T; t++) {
                size_t lo = (size_t)t * per, hi = lo + per > total ? total : lo + per;
                if (lo >= hi) break;
                jobs[t] = (fft_job){a, tw, n, lo, hi, half, tw_chunk};
                pthread_create(&th[t], NULL, fft_stage_worker, &jobs[t]);
                used++;
            }
            for (int t = 0; t < used; t++) pthread_join(th[t], NULL);
        }
        half *= 2;
        tw_chunk /= 2;
    }
    free(jobs); free(th); free(tw);
    return 0;
}

/* ----------------------------------------- EvaluationDomain transforms */
typedef struct { fe *a; size_t lo, hi; fe f0, f1, f2; int period3; } scale_job;
static void *scale_worker(void *arg) {
    scale_job *j = arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        if (!j->period3) { fe_mul(&FR, &j->a[i], &j->a[i], &j->f0); continue; }
        size_t r = i % 3;
        if (r == 1) fe_mul(&FR, &j->a[i], &j->a[i], &j->f1);
        else if (r == 2) fe_mul(&FR, &j->a[i], &j->a[i], &j->f2);
    }
    return NULL;
}
static void par_scale(fe *a, size_t n, int T, int period3, const fe *f0, const fe *f1, const fe *f2) {
    if (T < 1) T = 1;
    pthread_t *th = malloc((size_t)T * sizeof *th);
    scale_job *jobs = malloc((size_t)T * sizeof *jobs);
    size_t per = (n + (size_t)T - 1) / (size_t)T;
    int used = 0;
    for (int t = 0; t < T; t++) {
        size_t lo = (size_t)t * per, hi = lo + per > n ? n : lo + per;
        if (lo >= hi) break;
        jobs[t].a = a; jobs[t].lo = lo; jobs[t].hi = hi; jobs[t].period3 = period3;
        if (f0) jobs[t].f0 = *f0;
        if (f1) jobs[t].f1 = *f1;
        if (f2) jobs[t].f2 = *f2;
        pthread_create(&th[t], NULL, scale_worker, &jobs[t]);
        used++;
    }
    for (int t = 0; t < used; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* EvaluationDomain::ifft  poly/domain.rs:400-414 */
int ref_ifft(u64 *a, const u64 *omega_inv, const u64 *divisor, uint32_t log_n, int T) {
    ref_best_fft(a, omega_inv, log_n, T);
    par_scale((fe *)a, (size_t)1 << log_n, T, 0, (const fe *)divisor, NULL, NULL);
    return 0;
}
/* coeff_to_extended  poly/domain.rs:270-287.  `a` holds 2^ext_k elements of which the
 * first 2^k are the coefficients (the resize-with-zeros of :280 is done here). */
int ref_coeff_to_extended(u64 *a, uint32_t k, uint32_t ext_k, const u64 *zeta, const u64 *zeta_sq,
                          const u64 *ext_omega, int T) {
    size_t n = (size_t)1 << k, en = (size_t)1 << ext_k;
    par_scale((fe *)a, n, T, 1, NULL, (const fe *)zeta, (const fe *)zeta_sq);  /* :277, :382-398 */
    memset((fe *)a + n, 0, (en - n) * sizeof(fe));                             /* :280 */
    return ref_best_fft(a, ext_omega, ext_k, T);                               /* :281 */
}
/* extended_to_coeff  poly/domain.rs:328-350; caller truncates to n*(j-1). */
int ref_extended_to_coeff(u64 *a, uint32_t ext_k, const u64 *zeta, const u64 *zeta_sq,
                          const u64 *ext_omega_inv, const u64 *ext_divisor, int T) {
    ref_ifft(a, ext_omega_inv, ext_divisor, ext_k, T);                         /* :332-337 */
    par_scale((fe *)a, (size_t)1 << ext_k, T, 1, NULL, (const fe *)zeta_sq, (const fe *)zeta); /* :341 */
    return 0;
}

/* --------------------------------------------------------- test helpers */
/* out[i] = a[i] op b[i];  op: 0 mul, 1 add, 2 sub, 3 sqr(a);  field: 0 Fr, 1 Fq */
int ref_field_vec(int field, int op, const u64 *a, const u64 *b, size_t n, u64 *out) {
    const field_t *F = field ? &FQ : &FR;
    for (size_t i = 0; i < n; i++) {
        const fe *x = (const fe *)a + i, *y = (const fe *)b + i;
        fe *r = (fe *)out + i;
        switch (op) {
        case 0: fe_mul(F, r, x, y); break;
        case 1: fe_add(F, r, x, y); break;
        case 2: fe_sub(F, r, x, y); break;
        case 3: fe_sqr(F, r, x); break;
        default: return -1;
        }
    }
    return 0;
}
/* canonical (non-Montgomery) <-> Montgomery, vectorised */
int ref_to_mont(int field, const u64 *a, size_t n, u64 *out) {
    const field_t *F = field ? &FQ : &FR;
    for (size_t i = 0; i < n; i++) fe_mul(F, (fe *)out + i, (const fe *)a + i, &F->r2);
    return 0;
}
int ref_from_mont(int field, const u64 *a, size_t n, u64 *out) {
    const field_t *F = field ? &FQ : &FR;
    for (size_t i = 0; i < n; i++) fe_from_mont(F, out + 4 * i, (const fe *)a + i);
    return 0;
}
/* out_aff[i] = [k_i] G, k_i canonical 256-bit LE (not Montgomery); threaded */
typedef struct { const u64 *k; aff *out; size_t lo, hi; } gen_job;
static void *gen_worker(void *arg) {
    gen_job *j = arg;
    aff g; memset(&g, 0, sizeof g);
    fe two = {{2, 0, 0, 0}}, one = {{1, 0, 0, 0}};
    fe_mul(Q, &g.x, &one, &FQ.r2);
    fe_mul(Q, &g.y, &two, &FQ.r2);
    for (size_t i = j->lo; i < j->hi; i++) {
        jac p;
        jac_mul_u256(&p, &g, j->k + 4 * i);
        jac_to_affine(&j->out[i], &p);
    }
    return NULL;
}
int ref_g1_mul_gen(const u64 *k, size_t n, int T, u64 *out_aff) {
    if (T < 1) T = 1;
    pthread_t *th = malloc((size_t)T * sizeof *th);
    gen_job *jobs = malloc((size_t)T * sizeof *jobs);
    size_t per = (n + (size_t)T - 1) / (size_t)T;
    int used = 0;
    for (int t = 0; t < T; t++) {
        size_t lo = (size_t)t * per, hi = lo + per > n ? n : lo + per;
        if (lo >= hi) break;
        jobs[t] = (gen_job){k, (aff *)out_aff, lo, hi};
        pthread_create(&th[t], NULL, gen_worker, &jobs[t]);
        used++;
    }
    for (int t = 0; t < used; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
int ref_jac_to_affine(const u64 *in_jac, size_t n, u64 *out_aff) {
    for (size_t i = 0; i < n; i++) jac_to_affine((aff *)out_aff + i, (const jac *)in_jac + i);
    return 0;
}
/* naive sum_i [k_i] P_i with k_i Montgomery Fr (double-and-add), for cross-checks */
int ref_msm_naive(const u64 *coeffs, const u64 *bases, size_t n, u64 *out_jac) {
    jac acc; jac_set_id(&acc);
    for (size_t i = 0; i < n; i++) {
        u64 k[4];
        fe_from_mont(&FR, k, (const fe *)coeffs + i);
        jac t;
        jac_mul_u256(&t, (const aff *)bases + i, k);
        jac_add(&acc, &acc, &t);
    }
    memcpy(out_jac, &acc, sizeof acc);
    return 0;
}
/* sum of n Jacobian points (the reference's host-side combine of per-GPU partials,
 * arithmetic.rs:428-435) */
int ref_jac_sum(const u64 *in_jac, size_t n, u64 *out_jac) {
    jac acc; jac_set_id(&acc);
    for (size_t i = 0; i < n; i++) jac_add(&acc, &acc, (const jac *)in_jac + i);
    memcpy(out_jac, &acc, sizeof acc);
    return 0;
}
/* ------------------------------------------------------------------------------------------------
 * evaluate_h row loop: Calculation::evaluate / ValueSource::get over all rows
 * (plonk/evaluation.rs:60-268 and the "expressions" loop :846-1001): for every row, every Calculation
 * in order into `intermediates`, then the designated result.  Rows are split into one contiguous
 * chunk per thread (:849-851).  The calculation records have the layout of b2_qcalc (include/b2pcs.h),
 * i.e. the reference's enums flattened: kinds 0 Constant, 1 Intermediate, 2 Fixed, 3 Advice, 4 Instance,
 * 5 Aux (z / sigma / l0 ... cosets, which the reference reads in its hand-written permutation / lookup /
 * shuffle loops, :1005-1220), 6 Challenge, 7 coset point x0 * step^row (beta_term, :1018-1019);
 * ops 0 Add, 1 Sub, 2 Mul, 3 Negate, 4 LcChallenge, 5 a * ch + b (LcTheta / the y fold), 6 AddChallenge, 7 Store. */
typedef struct { uint32_t kind, index, rot; } qsrc_t;
typedef struct { uint32_t op; qsrc_t a, b; uint32_t challenge, power; } qcalc_t;
typedef struct {
    const int32_t *rotations; uint32_t n_rot;
    const fe *constants;
    const qcalc_t *calcs; uint32_t n_calcs;
    qsrc_t result;
    const fe *const *cols[4];          /* fixed, advice, instance, aux */
    const fe *challenges;
    fe x0, step;
    uint32_t log_rows, rot_scale;
    fe *out;
    size_t lo, hi;
} qjob;

static inline void q_get(const qjob *j, const qsrc_t *s, const size_t *rot_idx, const fe *inter, const fe *x, fe *r) {
    switch (s->kind) {
    case 0: *r = j->constants[s->index]; break;
    case 1: *r = inter[s->index]; break;
    case 2: case 3: case 4: case 5: *r = j->cols[s->kind - 2][s->index][rot_idx[s->rot]]; break;
    case 6: *r = j->challenges[s->index]; break;
    default: *r = *x; break;
    }
}

static void *q_worker(void *arg) {
    qjob *j = (qjob *)arg;
    const size_t size = (size_t)1 << j->log_rows;
    fe *inter = (fe *)malloc(sizeof(fe) * (j->n_calcs ? j->n_calcs : 1));
    fe *chp = (fe *)malloc(sizeof(fe) * (j->n_calcs ? j->n_calcs : 1));   /* challenge powers of the LcChallenge calcs */
    for (uint32_t c = 0; c < j->n_calcs; c++) {
        if (j->calcs[c].op != 4) continue;
        chp[c] = j->challenges[j->calcs[c].challenge];
        if (j->calcs[c].power > 1) {
            u64 e[4] = {j->calcs[c].power, 0, 0, 0};
            fe_pow(&FR, &chp[c], &j->challenges[j->calcs[c].challenge], e);
        }
    }
    size_t *rot_idx = (size_t *)malloc(sizeof(size_t) * (j->n_rot ? j->n_rot : 1));
    fe x;
    {   /* x0 * step^lo */
        u64 e[4] = {j->lo, 0, 0, 0};
        fe p;
        fe_pow(&FR, &p, &j->step, e);
        fe_mul(&FR, &x, &j->x0, &p);
    }
    for (size_t idx = j->lo; idx < j->hi; idx++) {
        for (uint32_t r = 0; r < j->n_rot; r++) {   /* get_rotation_idx, :40-42 */
            long long v = ((long long)idx + (long long)j->rotations[r] * (long long)j->rot_scale) % (long long)size;
            if (v < 0) v += (long long)size;
            rot_idx[r] = (size_t)v;
        }
        for (uint32_t c = 0; c < j->n_calcs; c++) {
            const qcalc_t *q = &j->calcs[c];
            fe a, b, t;
            q_get(j, &q->a, rot_idx, inter, &x, &a);
            switch (q->op) {
            case 0: q_get(j, &q->b, rot_idx, inter, &x, &b); fe_add(&FR, &inter[c], &a, &b); break;
            case 1: q_get(j, &q->b, rot_idx, inter, &x, &b); fe_sub(&FR, &inter[c], &a, &b); break;
            case 2: q_get(j, &q->b, rot_idx, inter, &x, &b); fe_mul(&FR, &inter[c], &a, &b); break;
            case 3: fe_neg(&FR, &inter[c], &a); break;
            case 4: {   /* (a + ch^p) * b, ch^1 when p <= 1 (:205-211) */
                q_get(j, &q->b, rot_idx, inter, &x, &b);
                fe_add(&FR, &t, &a, &chp[c]);
                fe_mul(&FR, &inter[c], &t, &b);
                break;
            }
            case 5: q_get(j, &q->b, rot_idx, inter, &x, &b); fe_mul(&FR, &t, &a, &j->challenges[q->challenge]);
                    fe_add(&FR, &inter[c], &t, &b); break;
            case 6: fe_add(&FR, &inter[c], &a, &j->challenges[q->challenge]); break;
            default: inter[c] = a; break;
            }
        }
        q_get(j, &j->result, rot_idx, inter, &x, &j->out[idx]);
        fe_mul(&FR, &x, &x, &j->step);
    }
    free(inter);
    free(chp);
    free(rot_idx);
    return NULL;
}

int ref_quotient_eval(const int32_t *rotations, uint32_t n_rot, const u64 *constants, const void *calcs, uint32_t n_calcs,
                      const uint32_t result[3], const u64 *const *fixed, const u64 *const *advice,
                      const u64 *const *instance, const u64 *const *aux, const u64 *challenges, const u64 *x0,
                      const u64 *step, uint32_t log_rows, uint32_t rot_scale, u64 *out, int T) {
    const size_t size = (size_t)1 << log_rows;
    if (T < 1) T = 1;
    if ((size_t)T > size) T = (int)size;
    qjob *jobs = (qjob *)calloc((size_t)T, sizeof(qjob));
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
    const size_t chunk = (size + (size_t)T - 1) / (size_t)T;
    for (int t = 0; t < T; t++) {
        qjob *j = &jobs[t];
        j->rotations = rotations; j->n_rot = n_rot;
        j->constants = (const fe *)constants;
        j->calcs = (const qcalc_t *)calcs; j->n_calcs = n_calcs;
        j->result.kind = result[0]; j->result.index = result[1]; j->result.rot = result[2];
        j->cols[0] = (const fe *const *)fixed; j->cols[1] = (const fe *const *)advice;
        j->cols[2] = (const fe *const *)instance; j->cols[3] = (const fe *const *)aux;
        j->challenges = (const fe *)challenges;
        if (x0) memcpy(&j->x0, x0, 32); else j->x0 = FR.one;
        if (step) memcpy(&j->step, step, 32); else j->step = FR.one;
        j->log_rows = log_rows; j->rot_scale = rot_scale;
        j->out = (fe *)out;
        j->lo = (size_t)t * chunk; j->hi = j->lo + chunk > size ? size : j->lo + chunk;
        if (j->lo > size) j->lo = size;
        pthread_create(&th[t], NULL, q_worker, j);
    }
    for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
    return 0;
}

int ref_version(void) { return 1; }
