// api_lookup.inl -- logup multiplicities on device-resident columns (included by api.cu).
// Reference: logup::Argument::compress, halo2_proofs/src/plonk/logup/prover.rs:117-179 (kernels: lookup.cuh).

namespace {

// one stable radix pass over byte `byte` of limb `limb`: idx_in -> idx_out
int lk_radix_pass(Lane& ctx, const uint4* keys, const uint32_t* idx_in, uint32_t* idx_out, uint32_t count, uint32_t limb,
                  uint32_t byte, cudaStream_t st) {
    const uint32_t nblocks = (count + LK_TILE - 1) / LK_TILE;
    const size_t nh = (size_t)256 * nblocks;
    const uint32_t ntiles = (uint32_t)((nh + SCAN_TILE - 1) / SCAN_TILE);
    int rc;
    if ((rc = ctx.lk_hist.reserve(nh * 4))) return rc;
    if ((rc = ctx.lk_offs.reserve((nh + 1) * 4))) return rc;
    if ((rc = ctx.tile_sums.reserve((size_t)ntiles * 4 + 16))) return rc;
    LAUNCH(ctx, lk_hist_kernel, nblocks, LK_THREADS, 0, st, keys, idx_in, count, limb, byte, ctx.lk_hist.as<uint32_t>(),
           nblocks);
    LAUNCH(ctx, msm_scan_tile_kernel, ntiles, SCAN_THREADS, 0, st, ctx.lk_hist.as<uint32_t>(), ctx.lk_offs.as<uint32_t>(),
           ctx.tile_sums.as<uint32_t>(), (uint32_t)nh);
    LAUNCH(ctx, msm_scan_top_kernel, 1, SCAN_THREADS, 0, st, ctx.tile_sums.as<uint32_t>(), ntiles,
           ctx.lk_offs.as<uint32_t>() + nh);
    LAUNCH(ctx, msm_scan_add_kernel, ntiles, SCAN_THREADS, 0, st, ctx.lk_offs.as<uint32_t>(), ctx.lk_hist.as<uint32_t>(),
           ctx.tile_sums.as<uint32_t>(), (uint32_t)nh);
    LAUNCH(ctx, lk_scatter_kernel, nblocks, LK_THREADS, 0, st, keys, idx_in, idx_out, count, limb, byte,
           ctx.lk_offs.as<uint32_t>(), nblocks);
    return B2_OK;
}

// stable LSD sort of the row indices over limbs [limb_lo, limb_hi], skipping the bytes in which no two keys differ
int lk_sort(Lane& ctx, const uint4* keys, uint32_t count, const unsigned long long* limb_or, int limb_lo, int limb_hi,
            uint32_t** sorted_idx, cudaStream_t st) {
    uint32_t* a = ctx.lk_idx_a.as<uint32_t>();
    uint32_t* b = ctx.lk_idx_b.as<uint32_t>();
    LAUNCH(ctx, lk_iota_kernel, (count + 255) / 256, 256, 0, st, a, count);
    for (int l = limb_lo; l <= limb_hi; l++)
        for (uint32_t byte = 0; byte < 8; byte++) {
            if (((limb_or[l] >> (8 * byte)) & 0xffull) == 0) continue;
            int rc = lk_radix_pass(ctx, keys, a, b, count, (uint32_t)l, byte, st);
            if (rc) return rc;
            std::swap(a, b);
        }
    *sorted_idx = a;
    return B2_OK;
}

}  // namespace

extern "C" {

int b2_logup_multiplicity_dev(const void* d_inputs, uint32_t n_inputs, const void* d_table, uint64_t usable, uint64_t n,
                              void* d_m, uint64_t* largest_count) {
    if (!d_table || !d_m || (n_inputs && !d_inputs) || usable == 0 || usable > n || n >= (1ull << 31))
        return fail(B2_ERR_ARG, "logup_multiplicity: bad arguments");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane& ctx = *ll.lane;
    cudaStream_t st = ctx.stream;
    const uint32_t cnt = (uint32_t)usable;
    if ((rc = ctx.lk_keys.reserve(usable * 32))) return rc;
    if ((rc = ctx.lk_skeys.reserve(usable * 32))) return rc;
    if ((rc = ctx.lk_idx_a.reserve(usable * 4))) return rc;
    if ((rc = ctx.lk_idx_b.reserve(usable * 4))) return rc;
    if ((rc = ctx.lk_counts.reserve(n * 4))) return rc;
    if ((rc = ctx.lk_flags.reserve(64))) return rc;
    if ((rc = ctx.h_stage.reserve(64))) return rc;
    // flags: [0..3] limb ORs (u64), [4] tie flag / miss flag (2 x int), [5] largest count (u32)
    unsigned long long* d_or = ctx.lk_flags.as<unsigned long long>();
    int* d_tie = reinterpret_cast<int*>(d_or + 4);
    int* d_miss = d_tie + 1;
    uint32_t* d_largest = reinterpret_cast<uint32_t*>(d_or + 5);
    unsigned long long* h = ctx.h_stage.as<unsigned long long>();
    CK(cudaMemsetAsync(ctx.lk_flags.p, 0, 64, st));
    CK(cudaMemsetAsync(ctx.lk_counts.p, 0, n * 4, st));
    uint4* keys = ctx.lk_keys.as<uint4>();
    uint4* skeys = ctx.lk_skeys.as<uint4>();
    LAUNCH(ctx, lk_canon_kernel, (cnt + LK_THREADS - 1) / LK_THREADS, LK_THREADS, 0, st, (const uint4*)d_table, keys,
           (unsigned long long)usable);
    LAUNCH(ctx, lk_limb_or_kernel, std::min<uint32_t>((cnt + LK_THREADS - 1) / LK_THREADS, (uint32_t)ctx.sms * 8),
           LK_THREADS, 0, st, (const uint4*)keys, (unsigned long long)usable, d_or);
    CK(cudaMemcpyAsync(h, d_or, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    unsigned long long limb_or[4] = {h[0], h[1], h[2], h[3]};
    int top = -1;
    for (int l = 3; l >= 0; l--)
        if (limb_or[l]) { top = l; break; }
    uint32_t* sidx = nullptr;
    if (top < 0) {
        // every table value is the same: the stable order is the row order
        sidx = ctx.lk_idx_a.as<uint32_t>();
        LAUNCH(ctx, lk_iota_kernel, (cnt + 255) / 256, 256, 0, st, sidx, cnt);
    } else {
        // the most significant differing limb orders the keys unless two DIFFERENT keys agree on it
        if ((rc = lk_sort(ctx, keys, cnt, limb_or, top, top, &sidx, st))) return rc;
        bool lower = false;
        for (int l = 0; l < top; l++) lower = lower || limb_or[l] != 0;
        if (lower) {
            LAUNCH(ctx, lk_gather_kernel, (cnt + LK_THREADS - 1) / LK_THREADS, LK_THREADS, 0, st, (const uint4*)keys,
                   (const uint32_t*)sidx, skeys, cnt);
            LAUNCH(ctx, lk_tie_kernel, (cnt + LK_THREADS - 1) / LK_THREADS, LK_THREADS, 0, st, (const uint4*)skeys, cnt,
                   (uint32_t)top, d_tie);
            CK(cudaMemcpyAsync(h, d_tie, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (*reinterpret_cast<int*>(h)) {
                if ((rc = lk_sort(ctx, keys, cnt, limb_or, 0, top, &sidx, st))) return rc;
            }
        }
    }
    LAUNCH(ctx, lk_gather_kernel, (cnt + LK_THREADS - 1) / LK_THREADS, LK_THREADS, 0, st, (const uint4*)keys,
           (const uint32_t*)sidx, skeys, cnt);
    if (n_inputs) {
        const unsigned long long total = usable * n_inputs;
        LAUNCH(ctx, lk_search_kernel, (unsigned)((total + LK_THREADS - 1) / LK_THREADS), LK_THREADS, 0, st,
               (const uint4*)d_inputs, (unsigned long long)n, cnt, n_inputs, (const uint4*)skeys, (const uint32_t*)sidx,
               ctx.lk_counts.as<uint32_t>(), d_miss);
    }
    LAUNCH(ctx, lk_finish_kernel, (unsigned)((n + LK_THREADS - 1) / LK_THREADS), LK_THREADS, 0, st,
           (const uint32_t*)ctx.lk_counts.as<uint32_t>(), cnt, (unsigned long long)n, (uint4*)d_m, d_largest);
    CK(cudaMemcpyAsync(h, d_tie, 16, cudaMemcpyDeviceToHost, st));     // tie, miss, largest
    CK(cudaStreamSynchronize(st));
    const int miss = reinterpret_cast<int*>(h)[1];
    if (miss) return fail(B2_ERR_ARG, "logup binary_search_by_key should hit: an input value is not in the table");
    if (largest_count) *largest_count = reinterpret_cast<uint32_t*>(h)[2];
    return B2_OK;
}

}  // extern "C"
