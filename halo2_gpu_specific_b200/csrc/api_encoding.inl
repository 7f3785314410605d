// api_encoding.inl -- Params::read / Params::write point (de)compression on the device (included by api.cu).
// Reference: halo2_proofs/src/poly/commitment.rs:241-294.

namespace {

int decompress_to(Lane& ctx, const void* bytes32, size_t n, uint32_t sign_bit, char* d_out) {
    int rc;
    if ((rc = ctx.ntt_in.reserve(n * 32))) return rc;
    if ((rc = ctx.errflag.reserve(16))) return rc;
    cudaStream_t st = ctx.stream;
    unsigned long long* bad = reinterpret_cast<unsigned long long*>(ctx.errflag.as<char>() + 8);
    CK(cudaMemsetAsync(bad, 0xff, 8, st));
    CK(cudaMemcpyAsync(ctx.ntt_in.p, bytes32, n * 32, cudaMemcpyHostToDevice, st));
    LAUNCH(ctx, g1_decompress_kernel, (unsigned)((n + 127) / 128), 128, 0, st, ctx.ntt_in.as<uint4>(), (uint4*)d_out,
           (unsigned long long)n, sign_bit, bad);
    unsigned long long h_bad = 0;
    CK(cudaMemcpyAsync(&h_bad, bad, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_bad != ~0ull)
        return fail(B2_ERR_ARG, "g1_decompress: point %llu is not a valid encoding (x >= q or x^3 + 3 is not a square)",
                    h_bad - 1);
    return B2_OK;
}

}  // namespace

extern "C" {

int b2_g1_decompress(const void* bytes32, size_t n, uint32_t sign_bit, void* out_affine64) {
    if ((n && (!bytes32 || !out_affine64)) || sign_bit > 7) return fail(B2_ERR_ARG, "g1_decompress: bad arguments");
    if (n == 0) return B2_OK;
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->ntt_out.reserve(n * 64))) return rc;
    if ((rc = decompress_to(*ctx, bytes32, n, sign_bit, ctx->ntt_out.as<char>()))) return rc;
    CK(cudaMemcpy(out_affine64, ctx->ntt_out.p, n * 64, cudaMemcpyDeviceToHost));
    return B2_OK;
}

int b2_g1_compress(const void* affine64, size_t n, uint32_t sign_bit, void* out_bytes32) {
    if ((n && (!affine64 || !out_bytes32)) || sign_bit > 7) return fail(B2_ERR_ARG, "g1_compress: bad arguments");
    if (n == 0) return B2_OK;
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->ntt_out.reserve(n * 64))) return rc;
    if ((rc = ctx->ntt_in.reserve(n * 32))) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->ntt_out.p, affine64, n * 64, cudaMemcpyHostToDevice, st));
    LAUNCH(*ctx, g1_compress_kernel, (unsigned)((n + 127) / 128), 128, 0, st, ctx->ntt_out.as<uint4>(),
           ctx->ntt_in.as<uint4>(), (unsigned long long)n, sign_bit);
    CK(cudaMemcpyAsync(out_bytes32, ctx->ntt_in.p, n * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

int b2_srs_register_compressed(const void* bytes32, size_t n, uint32_t sign_bit, b2_handle_t* out) {
    if (!bytes32 || !out || n == 0 || sign_bit > 7) return fail(B2_ERR_ARG, "srs_register_compressed: bad arguments");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    char* d = nullptr;
    CK(cudaMalloc(&d, n * 64));
    if ((rc = decompress_to(*ctx, bytes32, n, sign_bit, d))) {
        cudaFree(d);
        return rc;
    }
    std::lock_guard<std::mutex> lk2(g_srs_mu);
    b2_handle_t h = g_next_handle++;
    g_srs[h] = Srs{ctx->dev->dev, d, n, nullptr, 0, 0};
    *out = h;
    return B2_OK;
}

int b2_srs_read_compressed(b2_handle_t srs, size_t offset, size_t count, uint32_t sign_bit, void* out_bytes32) {
    Srs s;
    int rc = srs_lookup(srs, &s);
    if (rc) return rc;
    if (offset + count > s.n || sign_bit > 7 || (count && !out_bytes32)) return fail(B2_ERR_ARG, "srs_read_compressed: bad arguments");
    if (count == 0) return B2_OK;
    int save = g_dev;
    g_dev = s.device;
    LaneLock ll;
    rc = ll.acquire();
    g_dev = save;
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->ntt_in.reserve(count * 32))) return rc;
    cudaStream_t st = ctx->stream;
    LAUNCH(*ctx, g1_compress_kernel, (unsigned)((count + 127) / 128), 128, 0, st, (const uint4*)(s.d + offset * 64),
           ctx->ntt_in.as<uint4>(), (unsigned long long)count, sign_bit);
    CK(cudaMemcpyAsync(out_bytes32, ctx->ntt_in.p, count * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

}  // extern "C"
