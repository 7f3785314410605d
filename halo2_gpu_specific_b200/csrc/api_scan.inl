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
    int rc = ll.acquire();
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
    int rc = ll.acquire();
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
    int rc = ll.acquire();
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
