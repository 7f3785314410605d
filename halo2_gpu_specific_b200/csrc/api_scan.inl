// api_scan.inl -- batch inversion, prefix product / sum and element-wise ops on device-resident Fr
// vectors (included by api.cu).  Reference: batch_invert (halo2_proofs/src/arithmetic.rs:840-844),
// mul_acc (:806-836) and the running products / sums of the z polynomials
// (plonk/permutation/prover.rs:149-152, plonk/logup/prover.rs:318-336, plonk/shuffle/prover.rs:137-141).

namespace {

int scan_run(Lane& ctx, int op, const void* d_in, size_t n_in, const Fr& init, const void* d_init, void* d_out,
             size_t n_out, cudaStream_t st) {
    if (n_out == 0) return B2_OK;
    if (n_in + 1 < n_out) return fail(B2_ERR_ARG, "prefix_scan: n_out may be at most n_in + 1");
    const uint32_t ntiles = (uint32_t)((n_in + SCAN_FTILE - 1) / SCAN_FTILE);
    int rc;
    if ((rc = ctx.scan_tot.reserve((size_t)std::max(1u, ntiles) * 32))) return rc;
    uint4* tot = ctx.scan_tot.as<uint4>();
    if (ntiles == 0) {   // only out[0] = init
        if (d_init) CK(cudaMemcpyAsync(d_out, d_init, 32, cudaMemcpyDeviceToDevice, st));
        else CK(cudaMemcpyAsync(d_out, init.v, 32, cudaMemcpyHostToDevice, st));
        return B2_OK;
    }
    if (op == 0) {
        LAUNCH(ctx, scan_tile_kernel<0>, ntiles, SCAN_FT, 0, st, (const uint4*)d_in, (unsigned long long)n_in,
               (uint4*)d_out, (unsigned long long)n_out, tot);
        LAUNCH(ctx, scan_top_kernel<0>, 1, SCAN_FT, 0, st, tot, ntiles, init, (const uint4*)d_init);
        LAUNCH(ctx, scan_apply_kernel<0>, ntiles, SCAN_FT, 0, st, (uint4*)d_out, (unsigned long long)n_in,
               (unsigned long long)n_out, tot);
    } else {
        LAUNCH(ctx, scan_tile_kernel<1>, ntiles, SCAN_FT, 0, st, (const uint4*)d_in, (unsigned long long)n_in,
               (uint4*)d_out, (unsigned long long)n_out, tot);
        LAUNCH(ctx, scan_top_kernel<1>, 1, SCAN_FT, 0, st, tot, ntiles, init, (const uint4*)d_init);
        LAUNCH(ctx, scan_apply_kernel<1>, ntiles, SCAN_FT, 0, st, (uint4*)d_out, (unsigned long long)n_in,
               (unsigned long long)n_out, tot);
    }
    return B2_OK;
}

int invert_run(Lane& ctx, void* d_a, size_t n, cudaStream_t st) {
    if (n == 0) return B2_OK;
    int rc;
    if ((rc = ctx.scan_tmp.reserve(n * 32))) return rc;
    // enough threads to fill the machine, at least ~32 elements each so that the one Fermat inversion
    // per thread (~380 products) is amortised
    unsigned long long T = (unsigned long long)ctx.sms * 4 * 128;
    if (T * 32 > n) T = std::max<unsigned long long>(1, n / 32);
    T = (T + 127) / 128 * 128;
    LAUNCH(ctx, batch_invert_kernel, (unsigned)(T / 128), 128, 0, st, (uint4*)d_a, ctx.scan_tmp.as<uint4>(),
           (unsigned long long)n, T);
    return B2_OK;
}

}  // namespace

extern "C" {

int b2_batch_invert_dev(void* d_a, size_t n, void* stream) {
    if (!d_a && n) return fail(B2_ERR_ARG, "batch_invert: null pointer");
    LaneLock ll;
    int rc = ll.acquire((cudaStream_t)stream);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : ll.lane->stream;
    if ((rc = ll.order_after_busy(st))) return rc;
    if ((rc = invert_run(*ll.lane, d_a, n, st))) return rc;
    if (stream) return ll.mark_busy(st);
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

int b2_batch_invert(void* a, size_t n) {
    if (!a && n) return fail(B2_ERR_ARG, "batch_invert: null pointer");
    if (n == 0) return B2_OK;
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->ntt_in.reserve(n * 32))) return rc;
    CK(cudaMemcpyAsync(ctx->ntt_in.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = invert_run(*ctx, ctx->ntt_in.p, n, ctx->stream))) return rc;
    CK(cudaMemcpyAsync(a, ctx->ntt_in.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B2_OK;
}

int b2_prefix_scan_dev(int op, const void* d_in, size_t n_in, const void* init, const void* d_init, void* d_out,
                       size_t n_out, void* stream) {
    if (op < 0 || op > 1 || !d_out || (n_in && !d_in)) return fail(B2_ERR_ARG, "prefix_scan: bad arguments");
    LaneLock ll;
    int rc = ll.acquire((cudaStream_t)stream);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : ll.lane->stream;
    if ((rc = ll.order_after_busy(st))) return rc;
    Fr i0;
    if (init) {
        i0 = fr_from_bytes(init);
    } else {
        memset(i0.v, 0, 32);
        if (op == 0) memcpy(i0.v, HR_ONE, 32);
    }
    if ((rc = scan_run(*ll.lane, op, d_in, n_in, i0, d_init, d_out, n_out, st))) return rc;
    if (stream) return ll.mark_busy(st);
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

int b2_prefix_scan(int op, const void* in, size_t n_in, const void* init, void* out, size_t n_out) {
    if (op < 0 || op > 1 || !out || (n_in && !in)) return fail(B2_ERR_ARG, "prefix_scan: bad arguments");
    if (n_out == 0) return B2_OK;
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->ntt_in.reserve(std::max<size_t>(1, n_in) * 32))) return rc;
    if ((rc = ctx->ntt_out.reserve(n_out * 32))) return rc;
    if (n_in) CK(cudaMemcpyAsync(ctx->ntt_in.p, in, n_in * 32, cudaMemcpyHostToDevice, ctx->stream));
    Fr i0;
    if (init) {
        i0 = fr_from_bytes(init);
    } else {
        memset(i0.v, 0, 32);
        if (op == 0) memcpy(i0.v, HR_ONE, 32);
    }
    if ((rc = scan_run(*ctx, op, ctx->ntt_in.p, n_in, i0, nullptr, ctx->ntt_out.p, n_out, ctx->stream))) return rc;
    CK(cudaMemcpyAsync(out, ctx->ntt_out.p, n_out * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B2_OK;
}

int b2_fr_vec_dev(int op, const void* d_a, const void* d_b, size_t n, void* d_out, void* stream) {
    if (op < 0 || op > 2 || (n && (!d_a || !d_b || !d_out))) return fail(B2_ERR_ARG, "fr_vec_dev: bad arguments");
    if (n == 0) return B2_OK;
    LaneLock ll;
    int rc = ll.acquire((cudaStream_t)stream);
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    LAUNCH(*ctx, fr_vec_dev_kernel, (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sms * 16), 256, 0, st,
           (const uint4*)d_a, (const uint4*)d_b, (uint4*)d_out, (unsigned long long)n, op);
    if (stream) return B2_OK;   // uses no lane workspace
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

}  // extern "C"

// ---- eval_polynomial / kate_division on device-resident coefficient forms ----------------------------
namespace {

int hr_inv(uint64_t r[4], const uint64_t a[4]) {   // a^(r - 2)
    uint64_t acc[4], base[4];
    memcpy(acc, HR_ONE, 32);
    memcpy(base, a, 32);
    uint64_t e[4] = {HR_P[0] - 2, HR_P[1], HR_P[2], HR_P[3]};
    for (int i = 0; i < 256; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) hr_mul(acc, acc, base);
        hr_mul(base, base, base);
    }
    memcpy(r, acc, 32);
    return 0;
}

}  // namespace

extern "C" {

static int eval_polys_impl(const void* d_polys, const void* const* h_ptrs, uint64_t columns, uint64_t stride, uint64_t n,
                           const void* point, void* out_host);

int b2_eval_polynomial_dev(const void* d_polys, uint64_t columns, uint64_t stride, uint64_t n, const void* point,
                           void* out_host) {
    if (!d_polys || !point || !out_host || n == 0 || stride < n) return fail(B2_ERR_ARG, "eval_polynomial: bad arguments");
    return eval_polys_impl(d_polys, nullptr, columns, stride, n, point, out_host);
}

int b2_eval_polynomials_dev(const void* const* d_poly_ptrs, uint64_t count, uint64_t n, const void* point, void* out_host) {
    if ((count && !d_poly_ptrs) || !point || !out_host || n == 0) return fail(B2_ERR_ARG, "eval_polynomials: bad arguments");
    for (uint64_t i = 0; i < count; i++)
        if (!d_poly_ptrs[i]) return fail(B2_ERR_ARG, "eval_polynomials: null polynomial pointer");
    return eval_polys_impl(nullptr, d_poly_ptrs, count, n, n, point, out_host);
}

static int eval_polys_impl(const void* d_polys, const void* const* h_ptrs, uint64_t columns, uint64_t stride, uint64_t n,
                           const void* point, void* out_host) {
    if (columns == 0) return B2_OK;
    if (columns > 65535) return fail(B2_ERR_ARG, "eval_polynomial: at most 65535 polynomials per call");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = ctx->stream;
    const uint64_t threads = (n + EVAL_C - 1) / EVAL_C;
    const uint32_t bpc = (uint32_t)((threads + EVAL_T - 1) / EVAL_T);
    const size_t lo_n = (size_t)1 << EVAL_LO, hi_n = (size_t)(threads >> EVAL_LO) + 1;
    if ((rc = ctx->scan_tmp.reserve((lo_n + hi_n) * 32 + (size_t)columns * bpc * 32 + (size_t)columns * 32 +
                                    (size_t)columns * 8)))
        return rc;
    Fr* lo = ctx->scan_tmp.as<Fr>();
    Fr* hi = lo + lo_n;
    uint4* partials = reinterpret_cast<uint4*>(hi + hi_n);
    uint4* d_out = partials + 2ull * columns * bpc;
    const uint4** d_ptrs = nullptr;
    if (h_ptrs) {
        d_ptrs = reinterpret_cast<const uint4**>(d_out + 2ull * columns);
        CK(cudaMemcpyAsync(d_ptrs, h_ptrs, (size_t)columns * 8, cudaMemcpyHostToDevice, st));
    }
    const Fr x = fr_from_bytes(point);
    Fr one;
    memcpy(one.v, HR_ONE, 32);
    LAUNCH(*ctx, ntt_pow_table_kernel, (unsigned)((lo_n + 127) / 128), 128, 0, st, lo, x, (unsigned long long)EVAL_C,
           (uint32_t)lo_n, 0, one);
    LAUNCH(*ctx, ntt_pow_table_kernel, (unsigned)((hi_n + 127) / 128), 128, 0, st, hi, x,
           (unsigned long long)EVAL_C << EVAL_LO, (uint32_t)hi_n, 0, one);
    LAUNCH(*ctx, eval_poly_partial_kernel, dim3(bpc, (unsigned)columns), EVAL_T, 0, st, (const uint4*)d_polys,
           (unsigned long long)stride, (unsigned long long)n, x, lo, hi, partials, bpc, (const uint4* const*)d_ptrs);
    LAUNCH(*ctx, eval_poly_final_kernel, (unsigned)columns, EVAL_T, 0, st, partials, bpc, d_out);
    CK(cudaMemcpyAsync(out_host, d_out, (size_t)columns * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

int b2_eval_polynomial(const void* poly, uint64_t n, const void* point, void* out) {
    if (!poly || !point || !out || n == 0) return fail(B2_ERR_ARG, "eval_polynomial: bad arguments");
    void* d = nullptr;
    CK(cudaMalloc(&d, n * 32));
    cudaError_t e = cudaMemcpy(d, poly, n * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    int rc = e == cudaSuccess ? b2_eval_polynomial_dev(d, 1, n, n, point, out) : fail(B2_ERR_CUDA, "eval_polynomial: upload failed");
    cudaFree(d);
    return rc;
}

int b2_kate_division_dev(const void* d_a, uint64_t n, const void* b, void* d_q, void* stream) {
    if (!d_a || !b || !d_q || n < 2) return fail(B2_ERR_ARG, "kate_division: bad arguments");
    LaneLock ll;
    int rc = ll.acquire((cudaStream_t)stream);
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if ((rc = ll.order_after_busy(st))) return rc;
    uint64_t bb[4];
    memcpy(bb, b, 32);
    if ((bb[0] | bb[1] | bb[2] | bb[3]) == 0) {
        // b = 0: q[j] = a[j + 1]
        CK(cudaMemcpyAsync(d_q, (const char*)d_a + 32, (n - 1) * 32, cudaMemcpyDeviceToDevice, st));
    } else {
        uint64_t binv[4];
        hr_inv(binv, bb);
        if ((rc = ctx->ntt_work.reserve((size_t)n * 32 * 3 + 64))) return rc;   // b^i | b^-i | t, then P in ntt_out
        if ((rc = ctx->ntt_out.reserve((size_t)(n + 1) * 32))) return rc;
        Fr* bpow = ctx->ntt_work.as<Fr>();
        Fr* binvpow = bpow + n;
        uint4* t = reinterpret_cast<uint4*>(binvpow + n);
        const unsigned long long threads = (n + POW_SEQ - 1) / POW_SEQ;
        LAUNCH(*ctx, ntt_pow_seq_kernel, (unsigned)((threads + 127) / 128), 128, 0, st, bpow, fr_from_bytes(bb),
               (unsigned long long)n);
        LAUNCH(*ctx, ntt_pow_seq_kernel, (unsigned)((threads + 127) / 128), 128, 0, st, binvpow, fr_from_bytes(binv),
               (unsigned long long)n);
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sms * 16);
        LAUNCH(*ctx, kate_scale_kernel, grid, 256, 0, st, (const uint4*)d_a, bpow, t, (unsigned long long)n);
        Fr zero;
        memset(zero.v, 0, 32);
        if ((rc = scan_run(*ctx, 1, t, n, zero, nullptr, ctx->ntt_out.p, n + 1, st))) return rc;
        LAUNCH(*ctx, kate_finish_kernel, grid, 256, 0, st, ctx->ntt_out.as<uint4>(), binvpow, (uint4*)d_q,
               (unsigned long long)n);
    }
    if (stream) return ll.mark_busy(st);
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

// (plain v, floor(v 2^256 / r)) of a Montgomery-form scalar, on the host (see fr_shoup_companion)
static void hr_shoup_pair(const void* v_mont, Fr* plain, Fr* comp) {
    uint64_t vm[4], one[4] = {1, 0, 0, 0}, pl[4];
    memcpy(vm, v_mont, 32);
    hr_mul(pl, vm, one);
    memcpy(plain->v, pl, 32);
    static const uint64_t NINV[4] = {0xc2e1f593efffffffULL, 0x6586864b4c6911b3ULL, 0xe39a982899062391ULL,
                                     0x73f82f1d0d8341b2ULL};   // -r^-1 mod 2^256
    uint64_t o[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        unsigned __int128 carry = 0;
        for (int j = 0; i + j < 4; j++) {
            unsigned __int128 t = (unsigned __int128)vm[i] * NINV[j] + o[i + j] + carry;
            o[i + j] = (uint64_t)t;
            carry = t >> 64;
        }
    }
    memcpy(comp->v, o, 32);
}

int b2_poly_combine_dev(const void* const* d_polys, uint32_t m, uint64_t n, const void* v, void* d_out, void* stream) {
    if (!d_polys || !v || !d_out || m == 0 || n == 0) return fail(B2_ERR_ARG, "poly_combine: bad arguments");
    for (uint32_t j = 0; j < m; j++)
        if (!d_polys[j]) return fail(B2_ERR_ARG, "poly_combine: null polynomial %u", j);
    LaneLock ll;
    int rc = ll.acquire((cudaStream_t)stream);
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if ((rc = ll.order_after_busy(st))) return rc;
    if ((rc = ctx->qtab.reserve((size_t)m * sizeof(void*)))) return rc;
    // the pointer table is read by the kernel after this call may have returned (stream != NULL): stage it in
    // pageable memory the runtime copies synchronously
    CK(cudaMemcpyAsync(ctx->qtab.p, d_polys, (size_t)m * sizeof(void*), cudaMemcpyHostToDevice, st));
    Fr vp, vc;
    hr_shoup_pair(v, &vp, &vc);
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sms * 16);
    LAUNCH(*ctx, poly_combine_kernel, grid, 256, 0, st, (const uint4* const*)ctx->qtab.p, m, (unsigned long long)n, vp, vc,
           (uint4*)d_out);
    if (stream) return ll.mark_busy(st);
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

int b2_poly_combine(const void* const* polys, uint32_t m, uint64_t n, const void* v, void* out) {
    if (!polys || !v || !out || m == 0 || n == 0) return fail(B2_ERR_ARG, "poly_combine: bad arguments");
    void* d = nullptr;
    CK(cudaMalloc(&d, (size_t)(m + 1) * n * 32));
    std::vector<const void*> ptrs(m);
    int rc = B2_OK;
    for (uint32_t j = 0; j < m && rc == B2_OK; j++) {
        ptrs[j] = (char*)d + (size_t)j * n * 32;
        if (!polys[j] || cudaMemcpy((void*)ptrs[j], polys[j], n * 32, cudaMemcpyHostToDevice) != cudaSuccess)
            rc = fail(B2_ERR_CUDA, "poly_combine: upload of polynomial %u failed", j);
    }
    char* d_out = (char*)d + (size_t)m * n * 32;
    if (rc == B2_OK && cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) rc = fail(B2_ERR_CUDA, "poly_combine: sync failed");
    if (rc == B2_OK) rc = b2_poly_combine_dev(ptrs.data(), m, n, v, d_out, nullptr);
    if (rc == B2_OK && cudaMemcpy(out, d_out, n * 32, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail(B2_ERR_CUDA, "poly_combine: download failed");
    cudaFree(d);
    return rc;
}

int b2_kate_division(const void* a, uint64_t n, const void* b, void* q) {
    if (!a || !b || !q || n < 2) return fail(B2_ERR_ARG, "kate_division: bad arguments");
    void* d = nullptr;
    CK(cudaMalloc(&d, (size_t)(2 * n) * 32));
    cudaError_t e = cudaMemcpy(d, a, n * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    int rc = e == cudaSuccess ? b2_kate_division_dev(d, n, b, (char*)d + n * 32, nullptr) : fail(B2_ERR_CUDA, "kate_division: upload failed");
    if (rc == B2_OK && cudaMemcpy(q, (char*)d + n * 32, (n - 1) * 32, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail(B2_ERR_CUDA, "kate_division: download failed");
    cudaFree(d);
    return rc;
}

int b2_vanishing_random_poly_dev(const void* key32, const void* random, uint32_t k, size_t n, void* d_out, void* stream) {
    if (!key32 || !random || !d_out || k == 0 || k > 64 || n == 0 || n > (1ull << 30))
        return fail(B2_ERR_ARG, "vanishing_random_poly: bad arguments");
    VanishKey key;
    memcpy(key.k, key32, 32);       // little-endian words, as RFC 8439 lays a key out
    LaneLock ll;
    int rc = ll.acquire((cudaStream_t)stream);
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if ((rc = ctx->out96.reserve((size_t)k * 32))) return rc;      // the k "random" elements, staged on the device
    CK(cudaMemcpyAsync(ctx->out96.p, random, (size_t)k * 32, cudaMemcpyHostToDevice, st));
    LAUNCH(*ctx, vanishing_random_poly_kernel, (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sms * 16), 256, 0, st,
           (uint4*)d_out, (const uint4*)ctx->out96.p, (unsigned)k, (unsigned long long)n, key);
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

int b2_fr_max_bits_dev(const void* d_a, size_t n, uint32_t* bits) {
    if (!bits || (n && !d_a)) return fail(B2_ERR_ARG, "fr_max_bits: bad arguments");
    *bits = 0;
    if (n == 0) return B2_OK;
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->out96.reserve(96))) return rc;
    CK(cudaMemsetAsync(ctx->out96.p, 0, 4, st));
    LAUNCH(*ctx, fr_max_bits_kernel, (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sms * 16), 256, 0, st,
           (const uint4*)d_a, (unsigned long long)n, (unsigned*)ctx->out96.p);
    CK(cudaMemcpyAsync(bits, ctx->out96.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

}  // extern "C"
