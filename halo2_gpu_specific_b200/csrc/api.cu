// api.cu -- host runtime + extern "C" entry points declared in include/b2pcs.h.
//
// One DeviceCtx per GPU: a stream, a grow-only workspace (no cudaMalloc on the hot path
// after the first call of a given size), resident SRS buffers, cached NTT plans
// (twiddle tables per (omega, log_n, divisor)) and CUDA-event timing.  Calls on one device
// are serialised by the context mutex, which is what the reference does with
// acquire_gpu/release_gpu (halo2_proofs/src/arithmetic.rs:313-331).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <thread>

#include <fcntl.h>
#include <unistd.h>

#include "../../include/b2pcs.h"
#include "msm.cuh"
#include "probe_mixed.cuh"
#include "ntt.cuh"
#include "quotient.cuh"
#include "scan.cuh"
#include "encoding.cuh"
#include "lookup.cuh"

using namespace b2;

namespace {

thread_local std::string g_err;
thread_local int g_dev = 0;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(e_ == cudaErrorMemoryAllocation ? B2_ERR_OOM : B2_ERR_CUDA, "%s:%d %s: %s", \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e_));                         \
    } while (0)

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B2_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3);  // slack so sweeps do not realloc every size
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B2_ERR_OOM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        }
        cap = want;
        return B2_OK;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// page-locked host staging (grow-only): small results are copied here asynchronously and handed to the caller
// after the stream sync.  A cudaMemcpyAsync into PAGEABLE caller memory blocks the host until the producing
// kernels finish, which serialises every lane pipeline behind it.
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B2_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes < 4096 ? 4096 : bytes + (bytes >> 2);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B2_ERR_OOM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return B2_OK;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct Srs {
    int device;
    char* d;            // n affine points
    size_t n;
    char* table;        // optional window table: table[w * n + i] = 2^(c*w) * P_i, or nullptr
    uint32_t tc, tW;    // its window width and number of windows
};

struct NttPlan {
    uint32_t log_n = 0;
    int npass = 0;
    uint32_t mm[NTT_MAX_PASSES] = {0, 0, 0, 0};
    uint32_t tw_h = 0;
    Fr* tw_sub[NTT_MAX_PASSES] = {nullptr, nullptr, nullptr, nullptr};
    Fr* tw_sub_shoup[NTT_MAX_PASSES] = {nullptr, nullptr, nullptr, nullptr};   // (w, floor(w 2^256 / r)) pairs
    Fr* tw_lo = nullptr;
    Fr* tw_hi = nullptr;         // unscaled
    Fr* tw_hi_scaled = nullptr;  // times divisor (pass 0 of an iNTT); == tw_hi when no divisor
    Fr* tw_full = nullptr;       // w^e (times divisor), e < N/2: pass-0 twiddles with one product per element
    bool has_div = false;
    Fr div;
    std::vector<void*> owned;
};

struct DeviceCtx;

// One execution lane = a stream + its own grow-only workspace + events.  Concurrent host threads
// (the reference calls this path from rayon workers, plonk/prover.rs:293,470,535,561,643) each
// take a free lane, so the H2D copy of one call overlaps the kernels and the D2H of others
// instead of being serialised behind a single per-device lock.
struct Lane {
    DeviceCtx* dev = nullptr;
    int index = 0;
    cudaStream_t stream = nullptr;
    int sms = 0;
    int acc_blocks_per_sm = 1;
    std::atomic<uint64_t>* launch_counter = nullptr;
    // MSM workspace
    Buf scalars, codes, sorted, counts, offsets, cursor, buckets, part_pt, part_bucket, block_out, window_sums,
        out96, errflag, tmp_bases, partials, tile_sums, part_pt2, part_bucket2, part_pt3, part_bucket3, part_ids;
    // NTT workspace
    Buf ntt_in, ntt_work, ntt_out;
    // quotient evaluation: per-call tables, spilled slots
    Buf qtab, qspill;
    // batch inversion / prefix scans
    Buf scan_tmp, scan_tot;
    // logup multiplicities: canonical / sorted table keys, row permutations, radix histograms, counts
    Buf lk_keys, lk_skeys, lk_idx_a, lk_idx_b, lk_hist, lk_offs, lk_counts, lk_flags;
    // pinned landing zone of small device-to-host results (commitments of a batch)
    PinnedBuf h_stage;
    // timing
    cudaEvent_t ev[16];
    cudaEvent_t busy;            // last work enqueued on a caller-provided stream (async _dev calls)
    bool busy_valid = false;
    cudaStream_t busy_stream = nullptr;   // the caller stream that recorded `busy`
    double last_kernel_ms = 0, last_total_ms = 0;
    double phases[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

constexpr int MAX_LANES = 8;

struct DeviceCtx {
    int dev = -1;
    bool ready = false;
    std::mutex mu;                  // lane pool + lazy init
    std::condition_variable cv;
    int nlanes = 0;
    Lane lanes[MAX_LANES];
    bool lane_free[MAX_LANES];
    std::mutex plan_mu;             // NTT plan cache
    std::map<std::string, NttPlan*> plans;
    std::map<std::string, Fr*> pow_tables;   // g^j tables of coset transforms, keyed by (g, length)
    std::atomic<uint64_t> launches{0};
};

constexpr int MAX_DEV = 16;
DeviceCtx g_ctx[MAX_DEV];
std::mutex g_srs_mu;
std::map<b2_handle_t, Srs> g_srs;
b2_handle_t g_next_handle = 1;

struct LastCall {
    double kernel_ms = 0, total_ms = 0;
    Lane* lane = nullptr;
};
thread_local LastCall g_last;

int dev_get(DeviceCtx** out) {
    int dev = g_dev;
    if (dev < 0 || dev >= MAX_DEV) return fail(B2_ERR_ARG, "bad device %d", dev);
    DeviceCtx& c = g_ctx[dev];
    CK(cudaSetDevice(dev));
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.ready) {
        c.dev = dev;
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, dev));
        CK(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        CK(cudaFuncSetAttribute(ntt_pass_cluster2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        {
            const int tw_extra = (64 << NTT_TWSM) + 16;
            CK(cudaFuncSetAttribute(ntt_pass_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            CK(cudaFuncSetAttribute(ntt_pass_cluster2_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CK(cudaFuncSetAttribute(ntt_pass_cluster4_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CK(cudaFuncSetAttribute(ntt_pass_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024 + tw_extra));
            CK(cudaFuncSetAttribute(ntt_pass_cluster2_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + tw_extra));
            CK(cudaFuncSetAttribute(ntt_pass_cluster4_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + tw_extra));
            CK(cudaFuncSetAttribute(ntt_pass_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024 + tw_extra));
            CK(cudaFuncSetAttribute(ntt_pass_cluster2_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + tw_extra));
            CK(cudaFuncSetAttribute(ntt_pass_cluster4_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + tw_extra));
        }
        CK(cudaFuncSetAttribute(ntt_pass_cluster4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CK(cudaFuncSetAttribute(ntt_pass_mont_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        CK(cudaFuncSetAttribute(ntt_pass_mont_cluster2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CK(cudaFuncSetAttribute(ntt_pass_mont_cluster4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, msm_accumulate_kernel, 128, 0));
        int nl = 3;
        if (const char* e = getenv("B2_LANES")) {
            int v = atoi(e);
            if (v >= 1 && v <= MAX_LANES) nl = v;
        }
        for (int i = 0; i < nl; i++) {
            Lane& l = c.lanes[i];
            l.dev = &c;
            l.index = i;
            l.sms = prop.multiProcessorCount;
            l.acc_blocks_per_sm = nb > 0 ? nb : 1;
            l.launch_counter = &c.launches;
            // one step above the default (= lowest) priority: the small kernels of a pipelined batch (bound scans,
            // digit sorts) must not queue behind a caller's background work on a b2_stream_create stream
            CK(cudaStreamCreateWithPriority(&l.stream, cudaStreamNonBlocking, -1));
            for (auto& e : l.ev) CK(cudaEventCreate(&e));
            CK(cudaEventCreateWithFlags(&l.busy, cudaEventDisableTiming));
            c.lane_free[i] = true;
        }
        c.nlanes = nl;
        c.ready = true;
    }
    *out = &c;
    return B2_OK;
}

// RAII: take a free lane of the current device (blocks while all are taken).
// A lane may still carry asynchronous work that a _dev call enqueued on a caller's stream (its `busy` event): whoever
// takes it next orders after that work, so a synchronous pipeline that was handed such a lane would stall behind the
// caller's background stream.  Selection therefore goes: (1) for an asynchronous caller, the lane that already carries
// work of the same stream (ordering is free there); (2) a lane without pending asynchronous work -- asynchronous
// callers search from the top, synchronous ones from the bottom, so they stay out of each other's way; (3) any free lane.
struct LaneLock {
    Lane* lane = nullptr;
    static bool pending(Lane& l) {
        if (!l.busy_valid) return false;
        if (cudaEventQuery(l.busy) == cudaErrorNotReady) return true;
        cudaGetLastError();
        l.busy_valid = false;
        return false;
    }
    static int pick(DeviceCtx* c, cudaStream_t for_stream, bool idle_only) {
        if (for_stream)
            for (int i = 0; i < c->nlanes; i++)
                if (c->lane_free[i] && c->lanes[i].busy_valid && c->lanes[i].busy_stream == for_stream) return i;
        for (int j = 0; j < c->nlanes; j++) {
            const int i = for_stream ? c->nlanes - 1 - j : j;
            if (c->lane_free[i] && !pending(c->lanes[i])) return i;
        }
        if (idle_only) return -1;
        for (int i = 0; i < c->nlanes; i++)
            if (c->lane_free[i]) return i;
        return -1;
    }
    int acquire(cudaStream_t for_stream = nullptr) {
        DeviceCtx* c;
        int rc = dev_get(&c);
        if (rc) return rc;
        std::unique_lock<std::mutex> lk(c->mu);
        for (;;) {
            const int i = pick(c, for_stream, false);
            if (i >= 0) {
                c->lane_free[i] = false;
                lane = &c->lanes[i];
                break;
            }
            c->cv.wait(lk);
        }
        lk.unlock();
        if (lane->busy_valid) {  // async work of a previous _dev call may still use this workspace
            cudaError_t e = cudaStreamWaitEvent(lane->stream, lane->busy, 0);
            if (e != cudaSuccess) return fail(B2_ERR_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e));
        }
        return B2_OK;
    }
    // non-blocking, for the extra lanes of a pipeline: only a lane without pending asynchronous work
    bool try_acquire(DeviceCtx* c) {
        std::unique_lock<std::mutex> lk(c->mu);
        const int i = pick(c, nullptr, true);
        if (i >= 0) {
            c->lane_free[i] = false;
            lane = &c->lanes[i];
        }
        lk.unlock();
        if (!lane) return false;
        if (lane->busy_valid) cudaStreamWaitEvent(lane->stream, lane->busy, 0);
        return true;
    }
    // work was enqueued on a caller stream: the next user of the lane must order after it
    int order_after_busy(cudaStream_t user) {
        if (user != lane->stream && lane->busy_valid) CK(cudaStreamWaitEvent(user, lane->busy, 0));
        return B2_OK;
    }
    int mark_busy(cudaStream_t user) {
        if (user == lane->stream) return B2_OK;
        CK(cudaEventRecord(lane->busy, user));
        lane->busy_valid = true;
        lane->busy_stream = user;
        return B2_OK;
    }
    ~LaneLock() {
        if (!lane) return;
        DeviceCtx* c = lane->dev;
        {
            std::lock_guard<std::mutex> lk(c->mu);
            c->lane_free[lane->index] = true;
        }
        c->cv.notify_one();
    }
};

#define LAUNCH(ctx, kernel, grid, block, smem, st, ...)                                                    \
    do {                                                                                                   \
        kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                                            \
        (*(ctx).launch_counter)++;                                                                               \
        cudaError_t le_ = cudaGetLastError();                                                              \
        if (le_ != cudaSuccess)                                                                            \
            return fail(B2_ERR_CUDA, "%s:%d launch %s: %s", __FILE__, __LINE__, #kernel,                   \
                        cudaGetErrorString(le_));                                                          \
    } while (0)

// ---- host-side Fq (4 x 64-bit limbs): only used to normalise the single result point
typedef unsigned __int128 u128_t;
const uint64_t HQ_P[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
const uint64_t HQ_INV = 0x87d20782e4866389ULL;
const uint64_t HQ_ONE[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};

// generic 4-limb Montgomery product r = a * b * 2^-256 mod P (host; scalar bookkeeping only)
void hm_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4], const uint64_t P[4], uint64_t INV) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128_t c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128_t)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * INV;
        c = (u128_t)m * P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128_t)m * P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    bool ge = t[4] != 0;
    if (!ge) {
        ge = true;
        for (int i = 3; i >= 0; i--) {
            if (t[i] > P[i]) break;
            if (t[i] < P[i]) { ge = false; break; }
        }
    }
    if (ge) {
        u128_t br = 0;
        for (int i = 0; i < 4; i++) {
            u128_t d = (u128_t)t[i] - P[i] - (uint64_t)br;
            t[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
    }
    memcpy(r, t, 32);
}
void hq_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) { hm_mul(r, a, b, HQ_P, HQ_INV); }
// host Fr (challenge powers of the quotient programs)
const uint64_t HR_P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
const uint64_t HR_INV = 0xc2e1f593efffffffULL;
const uint64_t HR_ONE[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
void hr_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) { hm_mul(r, a, b, HR_P, HR_INV); }
void hr_pow(uint64_t r[4], const uint64_t a[4], uint64_t e) {
    uint64_t acc[4], base[4];
    memcpy(acc, HR_ONE, 32);
    memcpy(base, a, 32);
    while (e) {
        if (e & 1) hr_mul(acc, acc, base);
        hr_mul(base, base, base);
        e >>= 1;
    }
    memcpy(r, acc, 32);
}
void hq_inv(uint64_t r[4], const uint64_t a[4]) {
    uint64_t e[4] = {HQ_P[0] - 2, HQ_P[1], HQ_P[2], HQ_P[3]};
    uint64_t acc[4];
    memcpy(acc, HQ_ONE, 32);
    for (int i = 255; i >= 0; i--) {
        hq_mul(acc, acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) hq_mul(acc, acc, a);
    }
    memcpy(r, acc, 32);
}
// Jacobian (X, Y, Z) -> (x, y, 1) or the identity (0, 1, 0), in place (12 x u64, Montgomery)
void jac_normalise_host(void* p) {
    uint64_t v[12];
    memcpy(v, p, 96);
    uint64_t* X = v;
    uint64_t* Y = v + 4;
    uint64_t* Z = v + 8;
    if ((Z[0] | Z[1] | Z[2] | Z[3]) == 0) {
        memset(v, 0, 96);
        memcpy(v + 4, HQ_ONE, 32);
    } else {
        uint64_t zi[4], zi2[4], zi3[4];
        hq_inv(zi, Z);
        hq_mul(zi2, zi, zi);
        hq_mul(zi3, zi2, zi);
        hq_mul(X, X, zi2);
        hq_mul(Y, Y, zi3);
        memcpy(Z, HQ_ONE, 32);
    }
    memcpy(p, v, 96);
}

// A batch call pipelines its columns over every lane that is free right now: while one lane's
// kernels run, the next lane's H2D copy and the previous lane's D2H copy are in flight.
struct LaneSet {
    LaneLock primary;
    LaneLock extra[MAX_LANES];
    std::vector<Lane*> lanes;
    int acquire(int want, cudaStream_t for_stream = nullptr) {
        int rc = primary.acquire(for_stream);
        if (rc) return rc;
        lanes.push_back(primary.lane);
        for (int i = 0; i + 1 < want && i < MAX_LANES; i++)
            if (extra[i].try_acquire(primary.lane->dev)) lanes.push_back(extra[i].lane);
        return B2_OK;
    }
    int sync_all() {
        for (Lane* l : lanes) CK(cudaStreamSynchronize(l->stream));
        return B2_OK;
    }
};

Fr fr_from_bytes(const void* p) {
    Fr r;
    memcpy(r.v, p, 32);
    return r;
}

// ---------------------------------------------------------------------------- MSM
void msm_pick_config(size_t n, uint32_t max_bits, uint32_t* c_out, uint32_t* W_out) {
    if (max_bits > 254) max_bits = 254;
    uint32_t lg = 0;
    while ((2ull << lg) <= n) lg++;
    int c = (int)lg - 4;
    if (c < 8) c = 8;
    if (c > 16) c = 16;
    if (const char* e = getenv("B2_MSM_C")) {
        int v = atoi(e);
        if (v >= 8 && v <= 22) c = v;
    }
    uint32_t W = max_bits / (uint32_t)c + 1;
    if (W > 32) {  // cannot happen for c >= 8 and max_bits <= 254
        c = 8;
        W = max_bits / 8 + 1;
    }
    *c_out = (uint32_t)c;
    *W_out = W;
}

struct MsmBases {
    const char* d;        // plain affine bases (already offset), used when table == nullptr
    const char* table;    // window table of the whole SRS, or nullptr
    uint32_t tc, tW;
    size_t n_srs, offset; // table geometry: point (w, i) lives at table[w * n_srs + offset + i]
};

// device-pointer MSM on stream `st`; writes a 96 B Jacobian (not normalised) to d_out
int msm_run(Lane& ctx, const MsmBases& mb, const void* d_scalars, size_t n, uint32_t max_bits,
            void* d_out, cudaStream_t st, bool record_phases, bool reset_flag = true) {
    if (max_bits > 254) max_bits = 254;
    MsmGeom g;
    const bool pre = mb.table != nullptr;
    if (pre) {
        g.c = mb.tc;
        g.W = max_bits / g.c + 1;
        if (g.W > mb.tW) g.W = mb.tW;
    } else {
        msm_pick_config(n, max_bits, &g.c, &g.W);
    }
    g.B = 1u << (g.c - 1);
    g.n = (uint32_t)n;
    g.bucket_stride = pre ? 0u : g.B;
    g.point_stride = pre ? (uint32_t)mb.n_srs : 0u;
    g.point_offset = pre ? (uint32_t)mb.offset : 0u;
    const uint32_t nsets = pre ? 1u : g.W;
    const size_t nb = (size_t)nsets * g.B;
    const char* d_points = pre ? mb.table : mb.d;
    const uint32_t T = (uint32_t)ctx.sms * (uint32_t)ctx.acc_blocks_per_sm * 128u;
    // buckets that can be populated: with a single window (bounded scalars, e.g. 16-bit advice values)
    // digits stay below 2^max_bits, so the reduction only has to visit that prefix of the bucket set
    uint32_t b_used = g.B;
    if (g.W == 1 && max_bits < g.c - 1) b_used = 1u << max_bits;
    // buckets per reduce thread: enough blocks to fill the chip, fewer adds per bucket when there are many.  Swept on
    // B200 (profiles/r2_msm_rm.jsonl): at 2^19 buckets rm = 2 / 4 / 8 / 16 give a 1.76 / 1.14 / 0.91 / 0.55 ms reduce --
    // every block pays ~45 dependent point operations for its scan, tree and offset multiple, so fewer, longer blocks
    // that still fit one wave win; at 2^16 buckets rm = 2 (0.27 ms) beats 16 (0.43 ms).
    uint32_t rm = b_used >> 15;
    rm = rm < 2 ? 2 : (rm > 16 ? 16 : rm);

    const uint32_t bpw = (b_used + MSM_RT * rm - 1) / (MSM_RT * rm);
    const uint32_t ntiles = (uint32_t)((nb + SCAN_TILE - 1) / SCAN_TILE);

    int rc;
    if ((rc = ctx.sorted.reserve((size_t)g.W * n * 4))) return rc;
    if ((rc = ctx.counts.reserve(nb * 4))) return rc;
    if ((rc = ctx.offsets.reserve((nb + 1) * 4))) return rc;
    if ((rc = ctx.cursor.reserve(nb * 4))) return rc;
    if ((rc = ctx.tile_sums.reserve((size_t)ntiles * 4 + 16))) return rc;
    if ((rc = ctx.buckets.reserve(nb * 128))) return rc;
    if ((rc = ctx.part_pt.reserve((size_t)2 * T * 128))) return rc;
    if ((rc = ctx.part_bucket.reserve((size_t)2 * T * 4))) return rc;
    if ((rc = ctx.part_ids.reserve((size_t)2 * T * 4))) return rc;
    if ((rc = ctx.part_pt2.reserve((size_t)2 * (2 * T / PR_L + 2) * 128))) return rc;
    if ((rc = ctx.part_bucket2.reserve((size_t)2 * (2 * T / PR_L + 2) * 4))) return rc;
    if ((rc = ctx.part_pt3.reserve((size_t)2 * (2 * T / PR_L + 2) * 128))) return rc;
    if ((rc = ctx.part_bucket3.reserve((size_t)2 * (2 * T / PR_L + 2) * 4))) return rc;
    if ((rc = ctx.block_out.reserve((size_t)nsets * bpw * 128))) return rc;
    if ((rc = ctx.window_sums.reserve(32 * 128))) return rc;
    if ((rc = ctx.errflag.reserve(16))) return rc;

    cudaEvent_t* ev = ctx.ev;
    if (record_phases) {
        g_last.lane = &ctx;      // b2_last_msm_phases reads the events of the lane that ran the last recorded MSM
        CK(cudaEventRecord(ev[0], st));
    }
    if (reset_flag) CK(cudaMemsetAsync(ctx.errflag.p, 0, 4, st));
    // sort the (bucket, point) entries by bucket
    // (a two-level partitioned counting sort was measured in round 1: 2.3 ms against 1.1 ms at 2^22 for this
    // global-atomic counting sort, whose scatters already stay in L2; removed)
    {
        if ((rc = ctx.codes.reserve((size_t)g.W * n * 4))) return rc;
        CK(cudaMemsetAsync(ctx.counts.p, 0, nb * 4, st));
        LAUNCH(ctx, msm_digits_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (const uint4*)d_scalars,
               ctx.codes.as<uint32_t>(), ctx.counts.as<uint32_t>(), g, max_bits, ctx.errflag.as<int>());
        if (record_phases) CK(cudaEventRecord(ev[1], st));
        LAUNCH(ctx, msm_scan_tile_kernel, ntiles, SCAN_THREADS, 0, st, ctx.counts.as<uint32_t>(),
               ctx.offsets.as<uint32_t>(), ctx.tile_sums.as<uint32_t>(), (uint32_t)nb);
        LAUNCH(ctx, msm_scan_top_kernel, 1, SCAN_THREADS, 0, st, ctx.tile_sums.as<uint32_t>(), ntiles,
               ctx.offsets.as<uint32_t>() + nb);
        LAUNCH(ctx, msm_scan_add_kernel, ntiles, SCAN_THREADS, 0, st, ctx.offsets.as<uint32_t>(),
               ctx.cursor.as<uint32_t>(), ctx.tile_sums.as<uint32_t>(), (uint32_t)nb);
        if (record_phases) CK(cudaEventRecord(ev[2], st));
        LAUNCH(ctx, msm_scatter_kernel, dim3((unsigned)((n + 255) / 256), g.W), 256, 0, st,
               ctx.codes.as<uint32_t>(), ctx.cursor.as<uint32_t>(), ctx.sorted.as<uint32_t>(), g);
        if (record_phases) CK(cudaEventRecord(ev[3], st));
    }
    // every thread takes ceil(E / T) entries of the bucket-sorted list (E is read on the device)
    LAUNCH(ctx, msm_accumulate_kernel, T / 128, 128, 0, st, d_points, 64u, ctx.sorted.as<uint32_t>(),
           ctx.offsets.as<uint32_t>(), (uint32_t)nb, 16u, T, ctx.buckets.as<char>(), ctx.part_pt.as<char>(),
           ctx.part_bucket.as<uint32_t>());
    if (record_phases) CK(cudaEventRecord(ev[4], st));
    {
        // fast path (runs of <= FIX_G records), then levels for long runs only:
        // 2T records -> 2 * ceil(2T / PR_L) -> ... -> one thread; level kernels exit at once
        // unless the fast path raised the flag (second int of errflag)
        uint32_t nrec = 2 * T;
        int* need = ctx.errflag.as<int>() + 1;
        CK(cudaMemsetAsync(need, 0, 4, st));
        LAUNCH(ctx, msm_fixup_small_kernel, (nrec + 127) / 128, 128, 0, st, ctx.part_pt.as<char>(),
               ctx.part_bucket.as<uint32_t>(), nrec, ctx.part_ids.as<uint32_t>(), need, ctx.buckets.as<char>());
        char* in_pt = ctx.part_pt.as<char>();
        uint32_t* in_b = ctx.part_ids.as<uint32_t>();
        char* out_pt = ctx.part_pt2.as<char>();
        uint32_t* out_b = ctx.part_bucket2.as<uint32_t>();
        for (int level = 0;; level++) {
            const uint32_t nw = (nrec + PR_L - 1) / PR_L;   // one warp per PR_L records
            LAUNCH(ctx, msm_partial_reduce_kernel, (nw * 32 + 127) / 128, 128, 0, st, in_pt, in_b, nrec, out_pt, out_b,
                   nw, need, ctx.buckets.as<char>());
            if (nw == 1) break;
            nrec = 2 * nw;
            if (level == 0) {   // ping-pong between the two small buffers after the first level
                in_pt = out_pt;
                in_b = out_b;
                out_pt = ctx.part_pt3.as<char>();
                out_b = ctx.part_bucket3.as<uint32_t>();
            } else {
                std::swap(in_pt, out_pt);
                std::swap(in_b, out_b);
            }
        }
    }
    if (record_phases) CK(cudaEventRecord(ev[5], st));
    MsmGeom gr = g;
    gr.W = nsets;
    LAUNCH(ctx, msm_reduce_kernel, nsets * bpw, MSM_RT, MSM_RT * 128, st, ctx.buckets.as<char>(),
           ctx.offsets.as<uint32_t>(), gr, bpw, rm, ctx.block_out.as<char>());
    if (record_phases) CK(cudaEventRecord(ev[6], st));
    LAUNCH(ctx, msm_final_kernel, 1, MSM_FT, 0, st, ctx.block_out.as<char>(), g.c, nsets, bpw,
           ctx.window_sums.as<char>(), (char*)d_out);
    if (record_phases) CK(cudaEventRecord(ev[7], st));
    return B2_OK;
}

int msm_collect_phases(Lane& ctx) {
    CK(cudaEventSynchronize(ctx.ev[7]));
    for (int i = 0; i < 7; i++) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx.ev[i], ctx.ev[i + 1]));
        ctx.phases[i] = ms;
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx.ev[0], ctx.ev[7]));
    ctx.phases[7] = ms;
    ctx.last_kernel_ms = ms;
    g_last.kernel_ms = ms;
    return B2_OK;
}

int write_identity(Lane& ctx, void* d_out, cudaStream_t st) {
    // (0, 1, 0) in Montgomery form
    uint32_t h[24];
    memset(h, 0, sizeof h);
    const uint32_t one[8] = {FqParams::one0, FqParams::one1, FqParams::one2, FqParams::one3,
                             FqParams::one4, FqParams::one5, FqParams::one6, FqParams::one7};
    memcpy(h + 8, one, 32);
    CK(cudaMemcpyAsync(d_out, h, 96, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    (void)ctx;
    return B2_OK;
}

int srs_lookup(b2_handle_t h, Srs* out) {
    std::lock_guard<std::mutex> lk(g_srs_mu);
    auto it = g_srs.find(h);
    if (it == g_srs.end()) return fail(B2_ERR_HANDLE, "unknown SRS handle %llu", (unsigned long long)h);
    *out = it->second;
    return B2_OK;
}

// per launch (entry offsets are 32-bit); B2_MSM_MAX_N lowers it so that tests can exercise the split path
size_t msm_max_n() {
    static size_t v = 0;
    if (!v) {
        v = (size_t)1 << 26;
        if (const char* e = getenv("B2_MSM_MAX_N")) {
            long long x = atoll(e);
            if (x >= 64 && x <= (1ll << 26)) v = (size_t)x;
        }
    }
    return v;
}
#define MSM_MAX_N (msm_max_n())

// MSM over possibly > MSM_MAX_N points by splitting; all on stream `st`
int msm_run_split(Lane& ctx, const Srs& s, size_t offset, const char* d_scalars, size_t n, uint32_t max_bits,
                  void* d_out, cudaStream_t st, bool record, bool reset_flag = true) {
    MsmBases mb{s.d + offset * 64, s.table, s.tc, s.tW, s.n, offset};
    if (n <= MSM_MAX_N) return msm_run(ctx, mb, d_scalars, n, max_bits, d_out, st, record, reset_flag);
    size_t parts = (n + MSM_MAX_N - 1) / MSM_MAX_N;
    int rc;
    if ((rc = ctx.partials.reserve(parts * 96))) return rc;
    for (size_t p = 0; p < parts; p++) {
        size_t lo = p * MSM_MAX_N, cnt = std::min(MSM_MAX_N, n - lo);
        MsmBases pb{s.d + (offset + lo) * 64, s.table, s.tc, s.tW, s.n, offset + lo};
        if ((rc = msm_run(ctx, pb, d_scalars + lo * 32, cnt, max_bits, ctx.partials.as<char>() + p * 96, st,
                          record && p == 0, reset_flag && p == 0)))
            return rc;
    }
    LAUNCH(ctx, g1_sum_kernel, 1, 32, 0, st, ctx.partials.as<char>(), (uint32_t)parts, (char*)d_out);
    return B2_OK;
}

int check_bound_flag(Lane& ctx, cudaStream_t st) {
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx.errflag.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (flag) return fail(B2_ERR_BOUND, "a scalar exceeds the max_bits bound");
    return B2_OK;
}

// ---------------------------------------------------------------------------- NTT
int ntt_table(Lane& ctx, NttPlan* pl, Fr** out, const Fr& base, unsigned long long mult, uint32_t count,
              bool scaled, const Fr& scale) {
    void* p = nullptr;
    CK(cudaMalloc(&p, (size_t)count * 32));
    pl->owned.push_back(p);
    LAUNCH(ctx, ntt_pow_table_kernel, (count + 127) / 128, 128, 0, ctx.stream, (Fr*)p, base, mult, count,
           scaled ? 1 : 0, scale);
    *out = (Fr*)p;
    return B2_OK;
}

int ntt_get_plan(Lane& ctx, const void* omega, const void* divisor, uint32_t log_n, NttPlan** out) {
    std::string key((const char*)omega, 32);
    key.append((const char*)&log_n, 4);
    if (divisor) {
        key.push_back(1);
        key.append((const char*)divisor, 32);
    } else {
        key.push_back(0);
    }
    std::lock_guard<std::mutex> plk(ctx.dev->plan_mu);
    auto it = ctx.dev->plans.find(key);
    if (it != ctx.dev->plans.end()) {
        *out = it->second;
        return B2_OK;
    }
    NttPlan* pl = new NttPlan();
    pl->log_n = log_n;
    // digit sizes: up to 2^11 points per CTA (64 KB of shared memory, 2-3 CTAs per SM); digits of
    // 12 / 13 bits are owned by a cluster of 2 / 4 CTAs (distributed shared memory)
    uint32_t maxm = 11, maxc = 13;
    if (const char* e = getenv("B2_NTT_MAXM")) {
        int v = atoi(e);
        if (v >= 3 && v <= 12) maxm = (uint32_t)v;
    }
    if (const char* e = getenv("B2_NTT_MAXC")) {
        int v = atoi(e);
        if (v >= 11 && v <= 13) maxc = (uint32_t)v;
    }
    if (getenv("B2_NTT_CLUSTER") && atoi(getenv("B2_NTT_CLUSTER")) == 0 && maxc > 12) maxc = 12;  // 128 KB single CTA
    int P = (int)((log_n + maxm - 1) / maxm);
    if (maxm == 11) {
        const int Pc = (int)((log_n + maxc - 1) / maxc);   // passes when wide digits are allowed
        if (Pc < P) P = Pc;
    }
    if (P > NTT_MAX_PASSES) {
        delete pl;
        return fail(B2_ERR_ARG, "log_n %u needs more than %d passes", log_n, NTT_MAX_PASSES);
    }
    pl->npass = P;
    for (int i = 0; i < P; i++) pl->mm[i] = log_n / P + ((uint32_t)i < log_n % P ? 1 : 0);
    pl->tw_h = (log_n + 1) / 2;
    Fr w = fr_from_bytes(omega);
    Fr none = fr_from_bytes(omega);
    int rc;
    for (int i = 0; i < P; i++) {
        bool found = false;
        for (int j = 0; j < i; j++)
            if (pl->mm[j] == pl->mm[i]) {
                pl->tw_sub[i] = pl->tw_sub[j];
                pl->tw_sub_shoup[i] = pl->tw_sub_shoup[j];
                found = true;
                break;
            }
        if (found) continue;
        uint32_t cnt = pl->mm[i] >= 1 ? (1u << (pl->mm[i] - 1)) : 1u;
        if ((rc = ntt_table(ctx, pl, &pl->tw_sub[i], w, 1ull << (log_n - pl->mm[i]), cnt, false, none))) return rc;
        {
            void* p = nullptr;
            CK(cudaMalloc(&p, (size_t)64 << pl->mm[i]));      // four 16-byte planes of 2^m entries
            pl->owned.push_back(p);
            pl->tw_sub_shoup[i] = (Fr*)p;
            LAUNCH(ctx, ntt_shoup_table_kernel, ((1u << pl->mm[i]) + 127) / 128, 128, 0, ctx.stream,
                   (const Fr*)pl->tw_sub[i], (uint4*)p, pl->mm[i]);
        }
    }
    if (P > 1) {
        if ((rc = ntt_table(ctx, pl, &pl->tw_lo, w, 1ull, 1u << pl->tw_h, false, none))) return rc;
        if ((rc = ntt_table(ctx, pl, &pl->tw_hi, w, 1ull << pl->tw_h, 1u << (log_n - pl->tw_h), false, none)))
            return rc;
        pl->tw_hi_scaled = pl->tw_hi;
    }
    if (divisor) {
        pl->has_div = true;
        pl->div = fr_from_bytes(divisor);
        if (P > 1) {
            if ((rc = ntt_table(ctx, pl, &pl->tw_hi_scaled, w, 1ull << pl->tw_h, 1u << (log_n - pl->tw_h), true,
                                pl->div)))
                return rc;
        }
    }
    // full (half-length) pass-0 twiddle table while it stays small next to 180 GB of HBM
    size_t full_limit = (size_t)512 << 20;
    if (const char* e = getenv("B2_NTT_FULL_TW_MB")) full_limit = (size_t)atoll(e) << 20;
    if (P > 1 && (((size_t)16) << log_n) <= full_limit) {
        if ((rc = ntt_table(ctx, pl, &pl->tw_full, w, 1ull, 1u << (log_n - 1), divisor != nullptr,
                            divisor ? pl->div : none)))
            return rc;
    }
    CK(cudaStreamSynchronize(ctx.stream));
    ctx.dev->plans[key] = pl;
    *out = pl;
    return B2_OK;
}

// all-device NTT of `cols` columns.  in -> (work) -> out.  `work` must hold cols * 2^log_n
// elements when npass > 1.  out may alias in.
int ntt_run_dev(Lane& ctx, NttPlan* pl, const void* d_in, uint64_t in_stride, uint64_t n_in, void* d_out,
                uint64_t out_stride, uint64_t n_out, void* d_work, uint64_t cols, const Fr* coset_in,
                const Fr* coset_out, cudaStream_t st, const Fr* in_scale = nullptr) {
    const uint32_t k = pl->log_n;
    const uint64_t N = 1ull << k;
    uint32_t s_lo = k;
    for (int p = 0; p < pl->npass; p++) {
        NttPassArgs a;
        memset(&a, 0, sizeof a);
        const bool first = (p == 0), last = (p + 1 == pl->npass);
        s_lo -= pl->mm[p];
        a.in = (const uint4*)(first ? d_in : d_work);
        a.in_col_stride = first ? in_stride : N;
        a.out = (uint4*)(last ? d_out : d_work);
        a.out_col_stride = last ? out_stride : N;
        a.n_in = first ? n_in : N;
        a.n_out = last ? n_out : N;
        a.log_n = k;
        a.m = pl->mm[p];
        a.s_lo = s_lo;
        a.pass = (uint32_t)p;
        a.npass = (uint32_t)pl->npass;
        for (int i = 0; i < NTT_MAX_PASSES; i++) a.mm[i] = pl->mm[i];
        a.tw_h = pl->tw_h;
        static const bool use_shoup = !(getenv("B2_NTT_SHOUP") && atoi(getenv("B2_NTT_SHOUP")) == 0);
        a.tw_sub = use_shoup ? pl->tw_sub_shoup[p] : pl->tw_sub[p];
        a.tw_lo = pl->tw_lo;
        a.tw_hi = first ? pl->tw_hi_scaled : pl->tw_hi;
        a.tw_full = first ? pl->tw_full : nullptr;
        if (first && !last && pl->tw_full && pl->has_div) a.scale_out = 1;  // table entry 0 is the divisor, not 1
        if (first) a.in_scale = in_scale;
        if (first && coset_in) {
            a.coset_in = 1;
            a.zin1 = coset_in[0];
            a.zin2 = coset_in[1];
        }
        if (last && coset_out) {
            a.coset_out = 1;
            a.zout1 = coset_out[0];
            a.zout2 = coset_out[1];
        }
        if (last && pl->npass == 1 && pl->has_div) {
            a.scale_out = 1;
            a.scale = pl->div;
        }
        static const bool use_clusters = !(getenv("B2_NTT_CLUSTER") && atoi(getenv("B2_NTT_CLUSTER")) == 0);
        // measured on B200: 11- and 12-bit digits run best on a cluster of 2 (2^10 / 2^11 points per CTA),
        // 13-bit digits on a cluster of 4 (profiles/r1_ncu_summary.md)
        a.cl_log = !use_clusters ? 0 : (a.m >= 13 ? 2 : (a.m >= 11 ? 1 : 0));
        const uint32_t mloc = a.m - a.cl_log;
        const uint32_t threads = std::max(32u, 1u << (mloc >= NTT_RMAX ? mloc - NTT_RMAX : 0));
        // kernel variant (see ntt.cuh): 0 = strict butterflies, twiddles through L1; 1 = lazy; 2 = lazy + TMA-staged
        // shared-memory twiddles; 3 = shared-memory twiddles only
        static const int variant_env = getenv("B2_NTT_VARIANT") ? atoi(getenv("B2_NTT_VARIANT")) : NTT_DEFAULT_VARIANT;
        int variant = use_shoup ? variant_env : 0;
        if (variant < 0 || variant > 3) variant = 1;
        const bool tw_sm = (variant == 2 || variant == 3);
        const size_t smem = ((size_t)32 << mloc) + (tw_sm ? ((size_t)64 << NTT_TWSM) + 16 : 0);
        const uint64_t lines = N >> a.m;
        for (uint64_t c0 = 0; c0 < cols; c0 += 65535) {
            const uint64_t cc = std::min<uint64_t>(65535, cols - c0);
            NttPassArgs b = a;
            b.in = a.in + 2ull * c0 * a.in_col_stride;
            b.out = a.out + 2ull * c0 * a.out_col_stride;
            if (a.cl_log == 0) {
                const dim3 grid((unsigned)lines, (unsigned)cc);
                if (!use_shoup) LAUNCH(ctx, ntt_pass_mont_kernel, grid, threads, smem, st, b);
                else if (variant == 1) LAUNCH(ctx, ntt_pass_v1_kernel, grid, threads, smem, st, b);
                else if (variant == 2) LAUNCH(ctx, ntt_pass_v2_kernel, grid, threads, smem, st, b);
                else if (variant == 3) LAUNCH(ctx, ntt_pass_v3_kernel, grid, threads, smem, st, b);
                else LAUNCH(ctx, ntt_pass_kernel, grid, threads, smem, st, b);
            } else {
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof cfg);
                cfg.gridDim = dim3((unsigned)(lines << a.cl_log), (unsigned)cc);
                cfg.blockDim = dim3(threads);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = 1u << a.cl_log;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                cudaError_t le;
                if (use_shoup && variant == 1)
                    le = (a.cl_log == 1) ? cudaLaunchKernelEx(&cfg, ntt_pass_cluster2_v1_kernel, b)
                                         : cudaLaunchKernelEx(&cfg, ntt_pass_cluster4_v1_kernel, b);
                else if (use_shoup && variant == 2)
                    le = (a.cl_log == 1) ? cudaLaunchKernelEx(&cfg, ntt_pass_cluster2_v2_kernel, b)
                                         : cudaLaunchKernelEx(&cfg, ntt_pass_cluster4_v2_kernel, b);
                else if (use_shoup && variant == 3)
                    le = (a.cl_log == 1) ? cudaLaunchKernelEx(&cfg, ntt_pass_cluster2_v3_kernel, b)
                                         : cudaLaunchKernelEx(&cfg, ntt_pass_cluster4_v3_kernel, b);
                else if (use_shoup)
                    le = (a.cl_log == 1) ? cudaLaunchKernelEx(&cfg, ntt_pass_cluster2_kernel, b)
                                         : cudaLaunchKernelEx(&cfg, ntt_pass_cluster4_kernel, b);
                else
                    le = (a.cl_log == 1) ? cudaLaunchKernelEx(&cfg, ntt_pass_mont_cluster2_kernel, b)
                                         : cudaLaunchKernelEx(&cfg, ntt_pass_mont_cluster4_kernel, b);
                (*ctx.launch_counter)++;
                if (le != cudaSuccess) return fail(B2_ERR_CUDA, "cluster NTT launch: %s", cudaGetErrorString(le));
            }
        }
    }
    return B2_OK;
}

// g^j, j < count, resident and cached per device (one table per coset generator of a domain)
int ntt_pow_table(Lane& ctx, const void* gen, uint64_t count, const Fr** out) {
    std::string key((const char*)gen, 32);
    key.append((const char*)&count, 8);
    std::lock_guard<std::mutex> plk(ctx.dev->plan_mu);
    auto it = ctx.dev->pow_tables.find(key);
    if (it != ctx.dev->pow_tables.end()) {
        *out = it->second;
        return B2_OK;
    }
    void* p = nullptr;
    CK(cudaMalloc(&p, (size_t)count * 32));
    const unsigned long long threads = (count + POW_SEQ - 1) / POW_SEQ;
    LAUNCH(ctx, ntt_pow_seq_kernel, (unsigned)((threads + 127) / 128), 128, 0, ctx.stream, (Fr*)p, fr_from_bytes(gen),
           (unsigned long long)count);
    CK(cudaStreamSynchronize(ctx.stream));
    ctx.dev->pow_tables[key] = (Fr*)p;
    *out = (Fr*)p;
    return B2_OK;
}

size_t ntt_scratch_limit() {
    size_t gb = 24;
    if (const char* e = getenv("B2_NTT_SCRATCH_GB")) {
        int v = atoi(e);
        if (v >= 1 && v <= 160) gb = (size_t)v;
    }
    return gb << 30;
}

int copy2d(void* dst, uint64_t dst_stride, const void* src, uint64_t src_stride, uint64_t width, uint64_t rows,
           cudaMemcpyKind kind, cudaStream_t st) {
    if (rows == 0 || width == 0) return B2_OK;
    if (dst_stride == width && src_stride == width) {
        CK(cudaMemcpyAsync(dst, src, (size_t)width * rows * 32, kind, st));
    } else {
        CK(cudaMemcpy2DAsync(dst, (size_t)dst_stride * 32, src, (size_t)src_stride * 32, (size_t)width * 32,
                             (size_t)rows, kind, st));
    }
    return B2_OK;
}

// instruction-rate probes: best of three timed launches of a dependent-product kernel
template <class K>
int run_probe(Lane* ctx, K kernel, double* out) {
    cudaStream_t st = ctx->stream;
    const int iters = 2000, ILP = 4;
    const int blocks = ctx->sms * 8, threads = 256;
    kernel<<<blocks, threads, 0, st>>>(ctx->out96.as<uint4>(), 50);   // warm-up
    (*ctx->launch_counter)++;
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(ctx->ev[12], st));
        kernel<<<blocks, threads, 0, st>>>(ctx->out96.as<uint4>(), iters);
        (*ctx->launch_counter)++;
        CK(cudaEventRecord(ctx->ev[13], st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        const double rate = (double)blocks * threads * (double)iters * ILP / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    *out = best;
    return B2_OK;
}

}  // namespace

// ============================================================================ C ABI
template <int KIND>
static int run_pipe_probe(Lane* ctx, double* per_s) {
    cudaStream_t st = ctx->stream;
    const int iters = 4000, ILP = 8;
    const int blocks = ctx->sms * 8, threads = 256;
    unsigned long long* sink = reinterpret_cast<unsigned long long*>(ctx->out96.p);
    LAUNCH(*ctx, (pipe_probe_kernel<ILP, KIND>), blocks, threads, 0, st, sink, 50, 0x85EBCA77u);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(ctx->ev[12], st));
        LAUNCH(*ctx, (pipe_probe_kernel<ILP, KIND>), blocks, threads, 0, st, sink, iters, 0x85EBCA77u);
        CK(cudaEventRecord(ctx->ev[13], st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        const double rate = (double)blocks * threads * (double)iters * 8.0 * ILP / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    *per_s = best;
    return B2_OK;
}

extern "C" {

int b2_version(void) { return 100; }
const char* b2_last_error(void) { return g_err.c_str(); }

int b2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int b2_set_device(int device) {
    int n = b2_device_count();
    if (device < 0 || device >= n || device >= MAX_DEV) return fail(B2_ERR_ARG, "device %d of %d", device, n);
    g_dev = device;
    return B2_OK;
}
int b2_get_device(void) { return g_dev; }
int b2_synchronize(void) {
    DeviceCtx* dev;
    int rc = dev_get(&dev);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    return B2_OK;
}
int b2_stream_create(void** stream) {
    if (!stream) return fail(B2_ERR_ARG, "stream_create: null pointer");
    DeviceCtx* dev;
    int rc = dev_get(&dev);
    if (rc) return rc;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream = (void*)st;
    return B2_OK;
}
int b2_stream_synchronize(void* stream) {
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return B2_OK;
}
int b2_stream_destroy(void* stream) {
    if (stream) CK(cudaStreamDestroy((cudaStream_t)stream));
    return B2_OK;
}
uint64_t b2_launch_count(int reset) {
    DeviceCtx* dev;
    if (dev_get(&dev)) return 0;
    uint64_t v = dev->launches.load();
    if (reset) dev->launches.store(0);
    return v;
}

// ---- SRS
int b2_srs_register(const void* bases, size_t n, size_t stride_bytes, b2_handle_t* out) {
    if (!bases || !out || n == 0 || stride_bytes < 64) return fail(B2_ERR_ARG, "srs_register: bad arguments");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    char* d = nullptr;
    CK(cudaMalloc(&d, n * 64));
    cudaError_t e;
    if (stride_bytes == 64)
        e = cudaMemcpy(d, bases, n * 64, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    else
        e = cudaMemcpy2D(d, 64, bases, stride_bytes, 64, n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d);
        return fail(B2_ERR_CUDA, "srs upload: %s", cudaGetErrorString(e));
    }
    std::lock_guard<std::mutex> lk2(g_srs_mu);
    b2_handle_t h = g_next_handle++;
    g_srs[h] = Srs{ctx->dev->dev, d, n, nullptr, 0, 0};
    *out = h;
    return B2_OK;
}
int b2_srs_synthetic(size_t n, uint64_t first_index, uint64_t seed, b2_handle_t* out) {
    if (!out || n == 0) return fail(B2_ERR_ARG, "srs_synthetic: bad arguments");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    char* d = nullptr;
    CK(cudaMalloc(&d, n * 64));
    LAUNCH(*ctx, srs_synth_kernel, (unsigned)((n + 127) / 128), 128, 0, ctx->stream, d, (unsigned long long)n,
           (unsigned long long)first_index, (unsigned long long)seed);
    CK(cudaStreamSynchronize(ctx->stream));
    std::lock_guard<std::mutex> lk2(g_srs_mu);
    b2_handle_t h = g_next_handle++;
    g_srs[h] = Srs{ctx->dev->dev, d, n, nullptr, 0, 0};
    *out = h;
    return B2_OK;
}
int b2_srs_from_scalars_dev(const void* d_scalars, size_t n, b2_handle_t* out) {
    if (!out || !d_scalars || n == 0) return fail(B2_ERR_ARG, "srs_from_scalars: bad arguments");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    char* d = nullptr;
    CK(cudaMalloc(&d, n * 64));
    LAUNCH(*ctx, srs_from_scalars_kernel, (unsigned)((n + 127) / 128), 128, 0, ctx->stream, d, (const uint4*)d_scalars,
           (unsigned long long)n);
    CK(cudaStreamSynchronize(ctx->stream));
    std::lock_guard<std::mutex> lk2(g_srs_mu);
    b2_handle_t h = g_next_handle++;
    g_srs[h] = Srs{ctx->dev->dev, d, n, nullptr, 0, 0};
    *out = h;
    return B2_OK;
}
int b2_memcpy_d2d(void* dst_dev, const void* src_dev, size_t bytes) {
    DeviceCtx* dev;
    int rc = dev_get(&dev);
    if (rc) return rc;
    // a device-to-device cudaMemcpy returns before the copy has run, and the lanes' streams do not wait for the
    // legacy stream: wait here, so that whatever the caller launches next sees the data
    CK(cudaMemcpy(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaStreamSynchronize(cudaStreamLegacy));
    return B2_OK;
}
int b2_srs_precompute(b2_handle_t srs, uint32_t window_bits) {
    Srs s;
    int rc = srs_lookup(srs, &s);
    if (rc) return rc;
    if (s.table) return B2_OK;
    uint32_t c = window_bits;
    if (c == 0) {
        // measured on B200 (tools/window_variants.sh): 2^16 -> 16, 2^18 -> 17, 2^20..2^24 -> 20, 2^26 -> 22
        uint32_t lg = 0;
        while ((2ull << lg) <= s.n) lg++;
        int cc;
        if (lg >= 25) cc = 22;
        else if (lg >= 20) cc = 20;
        else if (lg >= 18) cc = 17;
        else if (lg >= 16) cc = 16;
        else cc = std::max(10, (int)lg - 2);
        c = (uint32_t)cc;
    }
    if (c < 8 || c > 24) return fail(B2_ERR_ARG, "srs_precompute: window_bits %u out of [8, 24]", c);
    const uint32_t W = 254 / c + 1;
    if ((unsigned long long)W * s.n >= (1ull << 31))
        return fail(B2_ERR_ARG, "srs_precompute: %u windows x %zu points exceed the 31-bit point index", W, s.n);
    int save = g_dev;
    g_dev = s.device;
    LaneLock ll;
    rc = ll.acquire();
    g_dev = save;
    if (rc) return rc;
    Lane* ctx = ll.lane;
    char* t = nullptr;
    cudaError_t e = cudaMalloc(&t, (size_t)W * s.n * 64);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B2_ERR_OOM, "srs_precompute: cudaMalloc(%zu): %s", (size_t)W * s.n * 64, cudaGetErrorString(e));
    }
    CK(cudaMemcpyAsync(t, s.d, s.n * 64, cudaMemcpyDeviceToDevice, ctx->stream));
    const unsigned blocks = (unsigned)((s.n + 128ull * PRE_K - 1) / (128ull * PRE_K));
    for (uint32_t w = 1; w < W; w++)
        LAUNCH(*ctx, srs_precompute_kernel, blocks, 128, 0, ctx->stream, t, (unsigned long long)s.n, w, c);
    CK(cudaStreamSynchronize(ctx->stream));
    std::lock_guard<std::mutex> lk2(g_srs_mu);
    auto it = g_srs.find(srs);
    if (it == g_srs.end()) {
        cudaFree(t);
        return fail(B2_ERR_HANDLE, "SRS freed during precompute");
    }
    it->second.table = t;
    it->second.tc = c;
    it->second.tW = W;
    return B2_OK;
}
int b2_srs_len(b2_handle_t srs, size_t* n) {
    Srs s;
    int rc = srs_lookup(srs, &s);
    if (rc) return rc;
    *n = s.n;
    return B2_OK;
}
int b2_srs_read(b2_handle_t srs, size_t offset, size_t count, void* out_affine64) {
    Srs s;
    int rc = srs_lookup(srs, &s);
    if (rc) return rc;
    if (offset + count > s.n) return fail(B2_ERR_ARG, "srs_read out of range");
    CK(cudaSetDevice(s.device));
    CK(cudaMemcpy(out_affine64, s.d + offset * 64, count * 64, cudaMemcpyDeviceToHost));
    return B2_OK;
}
int b2_srs_free(b2_handle_t srs) {
    std::lock_guard<std::mutex> lk(g_srs_mu);
    auto it = g_srs.find(srs);
    if (it == g_srs.end()) return fail(B2_ERR_HANDLE, "unknown SRS handle");
    cudaSetDevice(it->second.device);
    cudaFree(it->second.d);
    if (it->second.table) cudaFree(it->second.table);
    g_srs.erase(it);
    return B2_OK;
}

// ---- MSM
int b2_msm_config(b2_handle_t srs, size_t n, uint32_t max_bits, uint32_t* c, uint32_t* windows, uint32_t* bucket_sets) {
    if (max_bits > 254) max_bits = 254;
    Srs s;
    if (srs && srs_lookup(srs, &s) == B2_OK && s.table) {
        *c = s.tc;
        *windows = std::min(max_bits / s.tc + 1, s.tW);
        if (bucket_sets) *bucket_sets = 1;
        return B2_OK;
    }
    msm_pick_config(n, max_bits, c, windows);
    if (bucket_sets) *bucket_sets = *windows;
    return B2_OK;
}

int b2_msm_dev(b2_handle_t srs, size_t offset, const void* d_scalars, size_t n, uint32_t max_bits,
               void* d_out_jac96, void* stream) {
    Srs s;
    int rc = srs_lookup(srs, &s);
    if (rc) return rc;
    if (offset + n > s.n) return fail(B2_ERR_ARG, "msm: %zu scalars at offset %zu exceed SRS length %zu", n, offset, s.n);
    LaneLock ll;
    if ((rc = ll.acquire((cudaStream_t)stream))) return rc;
    Lane* ctx = ll.lane;
    if (s.device != ctx->dev->dev)
        return fail(B2_ERR_ARG, "SRS lives on device %d, current device is %d", s.device, ctx->dev->dev);
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if ((rc = ll.order_after_busy(st))) return rc;
    if (n == 0 || max_bits == 0) return write_identity(*ctx, d_out_jac96, st);
    rc = msm_run_split(*ctx, s, offset, (const char*)d_scalars, n, max_bits, d_out_jac96, st, true);
    if (rc) return rc;
    return ll.mark_busy(st);
}

int b2_msm(b2_handle_t srs, size_t offset, const void* scalars, size_t n, uint32_t max_bits, void* out_jac96) {
    if (!out_jac96 || (n && !scalars)) return fail(B2_ERR_ARG, "msm: null pointer");
    Srs s;
    int rc = srs_lookup(srs, &s);
    if (rc) return rc;
    if (offset + n > s.n) return fail(B2_ERR_ARG, "msm: %zu scalars at offset %zu exceed SRS length %zu", n, offset, s.n);
    LaneLock ll;
    if ((rc = ll.acquire())) return rc;
    Lane* ctx = ll.lane;
    if (s.device != ctx->dev->dev)
        return fail(B2_ERR_ARG, "SRS lives on device %d, current device is %d", s.device, ctx->dev->dev);
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->out96.reserve(96))) return rc;
    if (n == 0 || max_bits == 0) {
        if ((rc = write_identity(*ctx, ctx->out96.p, st))) return rc;
        CK(cudaMemcpy(out_jac96, ctx->out96.p, 96, cudaMemcpyDeviceToHost));
        return B2_OK;
    }
    if ((rc = ctx->scalars.reserve(n * 32))) return rc;
    CK(cudaEventRecord(ctx->ev[8], st));
    CK(cudaMemcpyAsync(ctx->scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
    if ((rc = msm_run_split(*ctx, s, offset, ctx->scalars.as<char>(), n, max_bits, ctx->out96.p, st, true)))
        return rc;
    CK(cudaMemcpyAsync(out_jac96, ctx->out96.p, 96, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[9], st));
    if ((rc = check_bound_flag(*ctx, st))) return rc;
    jac_normalise_host(out_jac96);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]));
    ctx->last_total_ms = ms;
    g_last.total_ms = ms;
    return B2_OK;
}

// ---- asynchronous host-pointer MSM: one caller thread keeps several MSMs in flight (the copy of column c + 1 runs
// under the kernels of column c on another lane), as the reference's rayon workers do with several threads
struct MsmTicket {
    LaneLock ll;
    void* out = nullptr;
    bool identity = false;
};
std::mutex g_ticket_mu;
std::map<uint64_t, MsmTicket*> g_tickets;
uint64_t g_next_ticket = 1;

int b2_msm_async(b2_handle_t srs, size_t offset, const void* scalars, size_t n, uint32_t max_bits, void* out_jac96,
                 uint64_t* ticket) {
    if (!out_jac96 || !ticket || (n && !scalars)) return fail(B2_ERR_ARG, "msm_async: null pointer");
    Srs s;
    int rc = srs_lookup(srs, &s);
    if (rc) return rc;
    if (offset + n > s.n)
        return fail(B2_ERR_ARG, "msm: %zu scalars at offset %zu exceed SRS length %zu", n, offset, s.n);
    std::unique_ptr<MsmTicket> t(new MsmTicket());
    if ((rc = t->ll.acquire())) return rc;
    Lane* ctx = t->ll.lane;
    if (s.device != ctx->dev->dev)
        return fail(B2_ERR_ARG, "SRS lives on device %d, current device is %d", s.device, ctx->dev->dev);
    cudaStream_t st = ctx->stream;
    t->out = out_jac96;
    if ((rc = ctx->out96.reserve(96))) return rc;
    if ((rc = ctx->h_stage.reserve(96))) return rc;
    if (n == 0 || max_bits == 0) {
        t->identity = true;
    } else {
        if ((rc = ctx->scalars.reserve(n * 32))) return rc;
        CK(cudaEventRecord(ctx->ev[8], st));
        CK(cudaMemcpyAsync(ctx->scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
        if ((rc = msm_run_split(*ctx, s, offset, ctx->scalars.as<char>(), n, max_bits, ctx->out96.p, st, true))) {
            cudaStreamSynchronize(st);      // nothing of this call may still read the caller's scalars
            return rc;
        }
        // the 96-byte result lands in the lane's pinned staging buffer (a copy into pageable caller memory would
        // block this call until the kernels are done)
        CK(cudaMemcpyAsync(ctx->h_stage.p, ctx->out96.p, 96, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev[9], st));
    }
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    const uint64_t id = g_next_ticket++;
    g_tickets[id] = t.release();
    *ticket = id;
    return B2_OK;
}

int b2_msm_wait(uint64_t ticket) {
    MsmTicket* raw = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_ticket_mu);
        auto it = g_tickets.find(ticket);
        if (it == g_tickets.end()) return fail(B2_ERR_HANDLE, "unknown MSM ticket %llu", (unsigned long long)ticket);
        raw = it->second;
        g_tickets.erase(it);
    }
    std::unique_ptr<MsmTicket> t(raw);       // releases the lane on every path out of here
    Lane* ctx = t->ll.lane;
    CK(cudaSetDevice(ctx->dev->dev));
    if (t->identity) {
        uint64_t v[12];
        memset(v, 0, sizeof v);
        memcpy(v + 4, HQ_ONE, 32);
        memcpy(t->out, v, 96);
        return B2_OK;
    }
    int rc = check_bound_flag(*ctx, ctx->stream);     // synchronises the lane's stream
    if (rc) return rc;
    memcpy(t->out, ctx->h_stage.p, 96);
    jac_normalise_host(t->out);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]));
    ctx->last_total_ms = ms;
    g_last.total_ms = ms;
    g_last.lane = ctx;
    return B2_OK;
}

int b2_best_multiexp(const void* coeffs, const void* bases, size_t n, void* out_jac96) {
    if (!out_jac96 || (n && (!coeffs || !bases))) return fail(B2_ERR_ARG, "best_multiexp: null pointer");
    b2_handle_t h = 0;
    if (n == 0) {
        LaneLock ll;
        int rc = ll.acquire();
        if (rc) return rc;
        Lane* ctx = ll.lane;
        if ((rc = ctx->out96.reserve(96))) return rc;
        if ((rc = write_identity(*ctx, ctx->out96.p, ctx->stream))) return rc;
        CK(cudaMemcpy(out_jac96, ctx->out96.p, 96, cudaMemcpyDeviceToHost));
        return B2_OK;
    }
    int rc = b2_srs_register(bases, n, 64, &h);
    if (rc) return rc;
    rc = b2_msm(h, 0, coeffs, n, 254, out_jac96);
    b2_srs_free(h);
    return rc;
}

int b2_g1_sum(const void* jac96, size_t count, void* out_jac96) {
    if (!out_jac96 || (count && !jac96)) return fail(B2_ERR_ARG, "g1_sum: null pointer");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->partials.reserve(std::max<size_t>(count, 1) * 96))) return rc;
    if ((rc = ctx->out96.reserve(96))) return rc;
    if (count) CK(cudaMemcpyAsync(ctx->partials.p, jac96, count * 96, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(*ctx, g1_sum_kernel, 1, 32, 0, ctx->stream, ctx->partials.as<char>(), (uint32_t)count,
           ctx->out96.as<char>());
    CK(cudaMemcpyAsync(out_jac96, ctx->out96.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    jac_normalise_host(out_jac96);
    return B2_OK;
}

int b2_g1_normalize(void* jac96, size_t count) {
    if (!jac96 && count) return fail(B2_ERR_ARG, "g1_normalize: null pointer");
    for (size_t i = 0; i < count; i++) jac_normalise_host((char*)jac96 + i * 96);
    return B2_OK;
}

int b2_g1_sum_groups_dev(const void* d_jac96, size_t count, size_t groups, void* d_out_jac96, void* stream) {
    if (!d_out_jac96 || (count && groups && !d_jac96)) return fail(B2_ERR_ARG, "g1_sum_groups_dev: null pointer");
    if (groups == 0) return B2_OK;
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    LAUNCH(*ctx, g1_sum_groups_kernel, (unsigned)((groups + 31) / 32), 32, 0, st, (const char*)d_jac96, (uint32_t)count,
           (uint32_t)groups, (char*)d_out_jac96);
    return B2_OK;   // uses no lane workspace
}

int b2_g1_sum_dev(const void* d_jac96, size_t count, void* d_out_jac96, void* stream) {
    if (!d_out_jac96 || (count && !d_jac96)) return fail(B2_ERR_ARG, "g1_sum_dev: null pointer");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    LAUNCH(*ctx, g1_sum_kernel, 1, 32, 0, st, (const char*)d_jac96, (uint32_t)count, (char*)d_out_jac96);
    return B2_OK;   // uses no lane workspace
}

// ---- NTT
int b2_ntt_exec(const b2_ntt_desc* d) {
    if (!d || !d->omega || !d->in || !d->out) return fail(B2_ERR_ARG, "ntt: null pointer");
    if (d->log_n < 1 || d->log_n > 28) return fail(B2_ERR_ARG, "ntt: log_n %u out of range [1, 28]", d->log_n);
    const uint64_t N = 1ull << d->log_n;
    if (d->n_in == 0 || d->n_in > N || d->n_out == 0 || d->n_out > N)
        return fail(B2_ERR_ARG, "ntt: n_in/n_out must be in [1, 2^log_n]");
    if (d->columns == 0) return B2_OK;
    if (d->in_stride < d->n_in || d->out_stride < d->n_out) return fail(B2_ERR_ARG, "ntt: stride shorter than column");
    if (d->location > 3) return fail(B2_ERR_ARG, "ntt: location must be 0..3");
    const bool wants_pipeline = d->location != 1 && d->columns > 1;
    LaneSet set;
    int rc = set.acquire(wants_pipeline ? MAX_LANES : 1, d->location == 1 ? (cudaStream_t)d->stream : nullptr);
    if (rc) return rc;
    Lane* ctx = set.primary.lane;
    NttPlan* pl;
    if ((rc = ntt_get_plan(*ctx, d->omega, d->divisor, d->log_n, &pl))) return rc;
    Fr cin[2], cout[2];
    if (d->coset_in) { cin[0] = fr_from_bytes(d->coset_in); cin[1] = fr_from_bytes((const char*)d->coset_in + 32); }
    if (d->coset_out) { cout[0] = fr_from_bytes(d->coset_out); cout[1] = fr_from_bytes((const char*)d->coset_out + 32); }
    const Fr* pcin = d->coset_in ? cin : nullptr;
    const Fr* pcout = d->coset_out ? cout : nullptr;
    const Fr* in_scale = nullptr;
    if (d->coset_gen && (rc = ntt_pow_table(*ctx, d->coset_gen, d->n_in, &in_scale))) return rc;

    // location: 0 host -> host, 1 device -> device, 2 host -> device, 3 device -> host
    const bool in_host = (d->location == 0 || d->location == 2);
    const bool out_host = (d->location == 0 || d->location == 3);
    const bool async = !in_host && !out_host && d->stream;
    const size_t col_bytes = (size_t)N * 32;
    // sub-batch so that the scratch stays bounded; host batches are cut finer so that the lanes
    // can overlap copy-in, transform and copy-out of neighbouring chunks
    uint64_t sub = std::max<uint64_t>(1, ntt_scratch_limit() / (col_bytes * 2));
    sub = std::min<uint64_t>(sub, d->columns);
    const size_t nl = set.lanes.size();
    if ((in_host || out_host) && nl > 1 && d->columns > 1) {
        uint64_t fine = std::max<uint64_t>(1, d->columns / (2 * nl));
        // keep chunks big enough to amortise launches: at least ~8 MiB of data
        const uint64_t min_cols = std::max<uint64_t>(1, ((size_t)8 << 20) / col_bytes);
        sub = std::min(sub, std::max(fine, min_cols));
    }
    cudaStream_t st0 = async ? (cudaStream_t)d->stream : ctx->stream;
    if ((rc = set.primary.order_after_busy(st0))) return rc;
    CK(cudaEventRecord(ctx->ev[8], st0));
    float k_ms_total = 0;
    uint64_t chunk_index = 0;
    for (uint64_t c0 = 0; c0 < d->columns; c0 += sub, chunk_index++) {
        const uint64_t cc = std::min<uint64_t>(sub, d->columns - c0);
        Lane* ln = async ? ctx : set.lanes[chunk_index % nl];
        cudaStream_t st = async ? st0 : ln->stream;
        const char* uin = (const char*)d->in + c0 * d->in_stride * 32;
        char* uout = (char*)d->out + c0 * d->out_stride * 32;
        if (pl->npass > 1 && (rc = ln->ntt_work.reserve(cc * col_bytes))) return rc;
        // input on the device
        const void* din = uin;
        uint64_t din_stride = d->in_stride;
        if (in_host) {
            // sized for N per column so that it can also receive the output of a multi-pass transform
            if ((rc = ln->ntt_in.reserve(cc * (out_host ? col_bytes : (size_t)d->n_in * 32)))) return rc;
            if ((rc = copy2d(ln->ntt_in.p, d->n_in, uin, d->in_stride, d->n_in, cc, cudaMemcpyHostToDevice, st))) return rc;
            din = ln->ntt_in.p;
            din_stride = d->n_in;
        }
        // output on the device
        void* dout = uout;
        uint64_t dout_stride = d->out_stride;
        if (out_host) {
            dout_stride = N;
            const bool reuse_in = in_host && (pl->npass > 1 || d->n_in == N);  // input is consumed by pass 0
            if (reuse_in) {
                dout = ln->ntt_in.p;
            } else {
                if ((rc = ln->ntt_out.reserve(cc * col_bytes))) return rc;
                dout = ln->ntt_out.p;
            }
        }
        const bool timed = (nl == 1 || !(in_host || out_host)) && !async;
        if (timed) CK(cudaEventRecord(ln->ev[10], st));
        if ((rc = ntt_run_dev(*ln, pl, din, din_stride, d->n_in, dout, dout_stride, d->n_out, ln->ntt_work.p, cc, pcin,
                              pcout, st, in_scale)))
            return rc;
        if (timed) CK(cudaEventRecord(ln->ev[11], st));
        if (out_host &&
            (rc = copy2d(uout, d->out_stride, dout, dout_stride, d->n_out, cc, cudaMemcpyDeviceToHost, st)))
            return rc;
        if (timed) {
            CK(cudaStreamSynchronize(st));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ln->ev[10], ln->ev[11]));
            k_ms_total += ms;
        }
    }
    if (!async) {
        if ((rc = set.sync_all())) return rc;
        CK(cudaEventRecord(ctx->ev[9], st0));
        CK(cudaStreamSynchronize(st0));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]));
        ctx->last_total_ms = ms;
        ctx->last_kernel_ms = k_ms_total;
        g_last.total_ms = ms;
        g_last.kernel_ms = k_ms_total;
    } else {
        if ((rc = set.primary.mark_busy(st0))) return rc;   // asynchronous on the caller's stream
    }
    return B2_OK;
}

int b2_best_fft(void* a, const void* omega, uint32_t log_n) {
    b2_ntt_desc d;
    memset(&d, 0, sizeof d);
    d.log_n = log_n;
    d.omega = omega;
    d.n_in = d.n_out = d.in_stride = d.out_stride = 1ull << log_n;
    d.columns = 1;
    d.in = a;
    d.out = a;
    return b2_ntt_exec(&d);
}
int b2_gpu_ifft(void* a, const void* omega_inv, uint32_t log_n, const void* divisor) {
    if (!divisor) return fail(B2_ERR_ARG, "gpu_ifft: divisor is required");
    b2_ntt_desc d;
    memset(&d, 0, sizeof d);
    d.log_n = log_n;
    d.omega = omega_inv;
    d.divisor = divisor;
    d.n_in = d.n_out = d.in_stride = d.out_stride = 1ull << log_n;
    d.columns = 1;
    d.in = a;
    d.out = a;
    return b2_ntt_exec(&d);
}
int b2_coeff_to_extended(const void* a, void* out, uint64_t columns, uint32_t k, uint32_t ext_k, const void* zeta,
                         const void* zeta_sq, const void* ext_omega) {
    if (!zeta || !zeta_sq || ext_k < k) return fail(B2_ERR_ARG, "coeff_to_extended: bad arguments");
    char z[64];
    memcpy(z, zeta, 32);
    memcpy(z + 32, zeta_sq, 32);
    b2_ntt_desc d;
    memset(&d, 0, sizeof d);
    d.log_n = ext_k;
    d.omega = ext_omega;
    d.coset_in = z;
    d.n_in = d.in_stride = 1ull << k;
    d.n_out = d.out_stride = 1ull << ext_k;
    d.columns = columns;
    d.in = a;
    d.out = out;
    return b2_ntt_exec(&d);
}
int b2_extended_to_coeff(const void* a, void* out, uint64_t n_out, uint32_t ext_k, const void* zeta,
                         const void* zeta_sq, const void* ext_omega_inv, const void* ext_divisor) {
    if (!zeta || !zeta_sq || !ext_divisor) return fail(B2_ERR_ARG, "extended_to_coeff: bad arguments");
    char z[64];  // moving out of the coset: {zeta^2, zeta}
    memcpy(z, zeta_sq, 32);
    memcpy(z + 32, zeta, 32);
    b2_ntt_desc d;
    memset(&d, 0, sizeof d);
    d.log_n = ext_k;
    d.omega = ext_omega_inv;
    d.divisor = ext_divisor;
    d.coset_out = z;
    d.n_in = d.in_stride = 1ull << ext_k;
    d.n_out = d.out_stride = n_out;
    d.columns = 1;
    d.in = a;
    d.out = out;
    return b2_ntt_exec(&d);
}
int b2_divide_by_vanishing_poly(void* a, uint32_t ext_k, const void* t_evaluations, uint32_t t_len) {
    if (!a || !t_evaluations || t_len == 0 || (t_len & (t_len - 1))) return fail(B2_ERR_ARG, "divide: bad arguments");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    const uint64_t N = 1ull << ext_k;
    if ((rc = ctx->ntt_in.reserve(N * 32))) return rc;
    if ((rc = ctx->ntt_out.reserve((size_t)t_len * 32))) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->ntt_in.p, a, N * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->ntt_out.p, t_evaluations, (size_t)t_len * 32, cudaMemcpyHostToDevice, st));
    LAUNCH(*ctx, fr_scale_periodic_kernel, (unsigned)std::min<uint64_t>((N + 255) / 256, ctx->sms * 16), 256, 0, st,
           ctx->ntt_in.as<uint4>(), ctx->ntt_out.as<Fr>(), (unsigned long long)N, t_len - 1);
    CK(cudaMemcpyAsync(a, ctx->ntt_in.p, N * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

// ---- fused commit + iNTT
static int commit_batch_impl(b2_handle_t srs, void* columns_data, uint64_t columns, size_t n, uint32_t max_bits, int do_ifft,
                             const void* omega_inv, const void* divisor, uint32_t log_n, void* out_jac96, void* d_keep,
                             int columns_on_device) {
    if (!columns_data || !out_jac96 || n == 0) return fail(B2_ERR_ARG, "commit_batch: bad arguments");
    if (do_ifft && (!omega_inv || !divisor || n != ((size_t)1 << log_n)))
        return fail(B2_ERR_ARG, "commit_batch: ifft needs n == 2^log_n, omega_inv and divisor");
    if (columns == 0) return B2_OK;
    Srs s;
    int rc = srs_lookup(srs, &s);
    if (rc) return rc;
    if (n > s.n) return fail(B2_ERR_ARG, "commit_batch: column length %zu exceeds SRS length %zu", n, s.n);
    LaneSet set;
    if ((rc = set.acquire(columns > 1 ? MAX_LANES : 1))) return rc;
    Lane* ctx = set.primary.lane;
    if (s.device != ctx->dev->dev)
        return fail(B2_ERR_ARG, "SRS lives on device %d, current device is %d", s.device, ctx->dev->dev);
    // An error in the middle of the batch (an allocation that fails, a CUDA error) must not return while other lanes
    // still copy from / into the caller's host buffers: whatever leaves this function early drains the lanes first
    // and clears their sticky bound flags, so the caller may free its buffers and the next call starts clean.
    struct Drain {
        LaneSet& set;
        bool armed = true;
        ~Drain() {
            if (!armed) return;
            for (Lane* l : set.lanes) {
                cudaStreamSynchronize(l->stream);
                if (l->errflag.p) cudaMemset(l->errflag.p, 0, 8);
            }
            cudaGetLastError();
        }
    } drain{set};
    NttPlan* pl = nullptr;
    if (do_ifft && (rc = ntt_get_plan(*ctx, omega_inv, divisor, log_n, &pl))) return rc;
    const size_t col_bytes = n * 32;
    // B2_MAX_BITS_AUTO: the bound of each column is the bit length of its largest scalar, found on the device right
    // after the column has landed (find_max_scalar_bits, plonk/prover.rs:945-962, 296)
    const bool auto_bits = max_bits == B2_MAX_BITS_AUTO;
    std::vector<char> lane_used(MAX_LANES + 1, 0);          // lanes whose bound flag was reset by an MSM of this batch
    if (max_bits > 254) max_bits = 254;
    const size_t nl = set.lanes.size();
    if ((rc = ctx->h_stage.reserve((size_t)columns * 96))) return rc;
    char* h_pts = ctx->h_stage.as<char>();
    CK(cudaEventRecord(ctx->ev[8], ctx->stream));
    // one column per step, lanes round-robin: copy-in of column i+1 and copy-out of column i-1 overlap
    // the MSM (+ iNTT) of column i
    for (uint64_t c = 0; c < columns; c++) {
        Lane* ln = set.lanes[c % nl];
        cudaStream_t st = ln->stream;
        char* h = (char*)columns_data + c * col_bytes;
        // where the column lives on the device: a lane staging buffer, or the caller's resident buffer
        char* dcol;
        if (columns_on_device) {
            dcol = h;
        } else if (d_keep) {
            dcol = (char*)d_keep + c * col_bytes;
        } else {
            if ((rc = ln->ntt_in.reserve(col_bytes))) return rc;
            dcol = ln->ntt_in.as<char>();
        }
        if ((rc = ln->out96.reserve(96))) return rc;
        if (do_ifft && pl->npass > 1 && (rc = ln->ntt_work.reserve(col_bytes))) return rc;
        if (!columns_on_device) CK(cudaMemcpyAsync(dcol, h, col_bytes, cudaMemcpyHostToDevice, st));
        uint32_t col_bits = max_bits;
        if (auto_bits) {
            // the host waits for THIS lane's copy + scan only; the MSM it then enqueues runs under the next column's
            // copy, which goes to another lane
            if ((rc = ln->scan_tot.reserve(64))) return rc;
            CK(cudaMemsetAsync(ln->scan_tot.p, 0, 4, st));
            LAUNCH(*ln, fr_max_bits_kernel, (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ln->sms * 16), 256, 0, st,
                   (const uint4*)dcol, (unsigned long long)n, (unsigned*)ln->scan_tot.p);
            CK(cudaMemcpyAsync(&col_bits, ln->scan_tot.p, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        if (col_bits == 0) {
            if ((rc = write_identity(*ln, ln->out96.p, st))) return rc;
        } else {
            // the bound flag is sticky per lane for the whole batch (reset once, read once)
            if ((rc = msm_run_split(*ln, s, 0, dcol, n, col_bits, ln->out96.p, st, false, !lane_used[c % nl])))
                return rc;
            lane_used[c % nl] = 1;
        }
        CK(cudaMemcpyAsync(h_pts + c * 96, ln->out96.p, 96, cudaMemcpyDeviceToHost, st));
        if (do_ifft) {
            if ((rc = ntt_run_dev(*ln, pl, dcol, n, n, dcol, n, n, ln->ntt_work.p, 1, nullptr, nullptr, st)))
                return rc;
            if (!d_keep && !columns_on_device) CK(cudaMemcpyAsync(h, dcol, col_bytes, cudaMemcpyDeviceToHost, st));
        }
    }
    int bound_flag_any = 0;
    if (max_bits != 0) {
        for (size_t l = 0; l < nl && l < columns; l++) {
            if (!lane_used[l]) continue;
            int flag = 0;
            CK(cudaMemcpyAsync(&flag, set.lanes[l]->errflag.p, 4, cudaMemcpyDeviceToHost, set.lanes[l]->stream));
            CK(cudaStreamSynchronize(set.lanes[l]->stream));
            bound_flag_any |= flag;
        }
    }
    if ((rc = set.sync_all())) return rc;
    CK(cudaEventRecord(ctx->ev[9], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]));
    ctx->last_total_ms = ms;
    ctx->last_kernel_ms = 0;
    g_last.total_ms = ms;
    g_last.kernel_ms = 0;
    drain.armed = false;      // everything has been synchronised above
    memcpy(out_jac96, h_pts, (size_t)columns * 96);
    for (uint64_t c = 0; c < columns; c++) jac_normalise_host((char*)out_jac96 + c * 96);
    if (bound_flag_any) return fail(B2_ERR_BOUND, "a scalar exceeds the max_bits bound");
    return B2_OK;
}

int b2_commit_batch(b2_handle_t srs, void* columns_data, uint64_t columns, size_t n, uint32_t max_bits, int do_ifft,
                    const void* omega_inv, const void* divisor, uint32_t log_n, void* out_jac96) {
    return commit_batch_impl(srs, columns_data, columns, n, max_bits, do_ifft, omega_inv, divisor, log_n, out_jac96, nullptr, 0);
}

int b2_commit_batch_resident(b2_handle_t srs, const void* columns_data, int columns_on_device, void* d_columns,
                             uint64_t columns, size_t n, uint32_t max_bits, int do_ifft, const void* omega_inv,
                             const void* divisor, uint32_t log_n, void* out_jac96) {
    if (columns_on_device) {
        if (!d_columns) return fail(B2_ERR_ARG, "commit_batch_resident: d_columns is NULL");
        return commit_batch_impl(srs, d_columns, columns, n, max_bits, do_ifft, omega_inv, divisor, log_n, out_jac96,
                                 nullptr, 1);
    }
    if (!d_columns) return fail(B2_ERR_ARG, "commit_batch_resident: d_columns is NULL");
    return commit_batch_impl(srs, const_cast<void*>(columns_data), columns, n, max_bits, do_ifft, omega_inv, divisor, log_n,
                             out_jac96, d_columns, 0);
}

// ---- witness file (halo2_proofs/src/helpers.rs:919-1015): u32 LE column count, then column i at byte offset
// 4 + i * 2^(k+5): 2^k field elements exactly as they sit in memory (Montgomery, 32 B).
static int witness_pread_all(int fd, char* dst, size_t bytes, uint64_t off) {
    while (bytes) {
        ssize_t got = pread(fd, dst, bytes, (off_t)off);
        if (got <= 0) return -1;
        dst += got;
        off += (uint64_t)got;
        bytes -= (size_t)got;
    }
    return 0;
}

int b2_witness_file_columns(const char* path, uint32_t* n_columns) {
    if (!path || !n_columns) return fail(B2_ERR_ARG, "witness_file_columns: bad arguments");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(B2_ERR_ARG, "witness file %s: cannot open", path);
    uint32_t len = 0;
    const int bad = witness_pread_all(fd, (char*)&len, 4, 0);
    close(fd);
    if (bad) return fail(B2_ERR_ARG, "witness file %s: no header", path);
    *n_columns = len;
    return B2_OK;
}

int b2_commit_witness_file(b2_handle_t srs, const char* path, uint32_t k, uint64_t first, uint64_t count, uint32_t max_bits,
                           void* d_keep, void* out_jac96) {
    if (!path || !out_jac96 || k > 26) return fail(B2_ERR_ARG, "commit_witness_file: bad arguments");
    if (count == 0) return B2_OK;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(B2_ERR_ARG, "witness file %s: cannot open", path);
    uint32_t len = 0;
    if (witness_pread_all(fd, (char*)&len, 4, 0) || first + count > len) {
        close(fd);
        return fail(B2_ERR_ARG, "witness file %s holds %u columns, asked for [%llu, %llu)", path, len,
                    (unsigned long long)first, (unsigned long long)(first + count));
    }
    const size_t n = (size_t)1 << k, col_bytes = n * 32;
    // two pinned staging buffers of `per` columns: the next group is read from the file (page cache / disk) while the
    // current one goes through the commit pipeline (H2D + MSM across the lanes)
    const uint64_t per = std::max<uint64_t>(1, std::min<uint64_t>(count, ((size_t)256 << 20) / col_bytes));
    char* stage[2] = {nullptr, nullptr};
    int rc = B2_OK;
    for (int i = 0; i < 2 && rc == B2_OK; i++)
        if (cudaMallocHost((void**)&stage[i], per * col_bytes) != cudaSuccess) {
            cudaGetLastError();
            rc = fail(B2_ERR_OOM, "commit_witness_file: cannot pin %zu bytes", (size_t)(per * col_bytes));
        }
    auto read_group = [&](uint64_t g0, char* dst) -> int {
        const uint64_t cnt = std::min<uint64_t>(per, count - g0);
        for (uint64_t c = 0; c < cnt; c++)
            if (witness_pread_all(fd, dst + c * col_bytes, col_bytes, 4ull + ((first + g0 + c) << (k + 5)))) return -1;
        return 0;
    };
    if (rc == B2_OK && read_group(0, stage[0])) rc = fail(B2_ERR_ARG, "witness file %s: short read", path);
    int cur = 0;
    for (uint64_t g0 = 0; g0 < count && rc == B2_OK; g0 += per, cur ^= 1) {
        const uint64_t cnt = std::min<uint64_t>(per, count - g0);
        int next_bad = 0;
        std::thread reader;
        if (g0 + per < count) reader = std::thread([&, g0] { next_bad = read_group(g0 + per, stage[cur ^ 1]); });
        rc = commit_batch_impl(srs, stage[cur], cnt, n, max_bits, 0, nullptr, nullptr, k, (char*)out_jac96 + g0 * 96,
                               d_keep ? (char*)d_keep + g0 * col_bytes : nullptr, 0);
        if (reader.joinable()) reader.join();
        if (rc == B2_OK && next_bad) rc = fail(B2_ERR_ARG, "witness file %s: short read", path);
    }
    for (int i = 0; i < 2; i++)
        if (stage[i]) cudaFreeHost(stage[i]);
    close(fd);
    return rc;
}

int b2_msm_and_ifft(b2_handle_t srs, void* coeffs, uint32_t max_bits, const void* omega_inv, const void* divisor,
                    uint32_t log_n, void* out_jac96) {
    return b2_commit_batch(srs, coeffs, 1, (size_t)1 << log_n, max_bits, 1, omega_inv, divisor, log_n, out_jac96);
}

// ---- memory helpers
int b2_host_alloc(size_t bytes, void** out) {
    DeviceCtx* dev;
    int rc = dev_get(&dev);
    if (rc) return rc;
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B2_ERR_OOM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return B2_OK;
}
int b2_host_register(void* p, size_t bytes) {
    DeviceCtx* dev;
    int rc = dev_get(&dev);
    if (rc) return rc;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B2_ERR_CUDA, "cudaHostRegister(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return B2_OK;
}
int b2_host_unregister(void* p) {
    CK(cudaHostUnregister(p));
    return B2_OK;
}
int b2_host_free(void* p) {
    CK(cudaFreeHost(p));
    return B2_OK;
}
int b2_dev_alloc(size_t bytes, void** out) {
    DeviceCtx* dev;
    int rc = dev_get(&dev);
    if (rc) return rc;
    CK(cudaMalloc(out, bytes));
    return B2_OK;
}
int b2_dev_free(void* p) {
    CK(cudaFree(p));
    return B2_OK;
}
int b2_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes) {
    DeviceCtx* dev;
    int rc = dev_get(&dev);
    if (rc) return rc;
    // from pageable memory cudaMemcpy may return once the data is staged, before the DMA has landed; the lanes'
    // streams do not wait for the legacy stream, so wait here
    CK(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    CK(cudaStreamSynchronize(cudaStreamLegacy));
    return B2_OK;
}
int b2_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes) {
    DeviceCtx* dev;
    int rc = dev_get(&dev);
    if (rc) return rc;
    CK(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return B2_OK;
}

// ---- diagnostics
int b2_field_vec(int field, int op, const void* a, const void* b, size_t n, void* out) {
    if (!a || !b || !out || op < 0 || op > 6 || field < 0 || field > 1 || (op == 4 && field != 0))
        return fail(B2_ERR_ARG, "field_vec: bad arguments");
    if (n == 0) return B2_OK;
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->ntt_in.reserve(n * 32))) return rc;
    if ((rc = ctx->ntt_work.reserve(n * 32))) return rc;
    if ((rc = ctx->ntt_out.reserve(n * 32))) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->ntt_in.p, a, n * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->ntt_work.p, b, n * 32, cudaMemcpyHostToDevice, st));
    if (field == 0)
        LAUNCH(*ctx, field_vec_kernel<FrParams>, (unsigned)((n + 127) / 128), 128, 0, st, ctx->ntt_in.as<uint4>(),
               ctx->ntt_work.as<uint4>(), ctx->ntt_out.as<uint4>(), (unsigned long long)n, op);
    else
        LAUNCH(*ctx, field_vec_kernel<FqParams>, (unsigned)((n + 127) / 128), 128, 0, st, ctx->ntt_in.as<uint4>(),
               ctx->ntt_work.as<uint4>(), ctx->ntt_out.as<uint4>(), (unsigned long long)n, op);
    CK(cudaMemcpyAsync(out, ctx->ntt_out.p, n * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B2_OK;
}

int b2_imad_probe(double* wide_macs_per_s, double* modmuls_per_s) {
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->out96.reserve(96))) return rc;
    cudaStream_t st = ctx->stream;
    const int iters = 2000, ILP = 4;
    const int blocks = ctx->sms * 8, threads = 256;
    LAUNCH(*ctx, imad_probe_kernel<ILP>, blocks, threads, 0, st, ctx->out96.as<uint4>(), 50);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(ctx->ev[12], st));
        LAUNCH(*ctx, imad_probe_kernel<ILP>, blocks, threads, 0, st, ctx->out96.as<uint4>(), iters);
        CK(cudaEventRecord(ctx->ev[13], st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        double muls = (double)blocks * threads * (double)iters * ILP;
        double rate = muls / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    if (modmuls_per_s) *modmuls_per_s = best;
    if (wide_macs_per_s) *wide_macs_per_s = best * 128.0;
    return B2_OK;
}

int b2_mul_probe(int kind, double* per_s) {
    if (!per_s || kind < 0 || kind > 4) return fail(B2_ERR_ARG, "mul_probe: kind 0..4");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->out96.reserve(96))) return rc;
    switch (kind) {
    case 0: return run_probe(ctx, imad_probe_kernel<4>, per_s);
    case 1: return run_probe(ctx, shoup_probe_kernel<4>, per_s);
    case 2:
    case 3: return fail(B2_ERR_ARG, "mul_probe: kinds 2 and 3 (generated squaring / Karatsuba) were removed after round 1");
    default: return run_probe(ctx, mul_probe_kernel<4, 4>, per_s);
    }
}

int b2_shoup_probe(double* muls_per_s) {
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->out96.reserve(96))) return rc;
    cudaStream_t st = ctx->stream;
    const int iters = 2000, ILP = 4;
    const int blocks = ctx->sms * 8, threads = 256;
    LAUNCH(*ctx, shoup_probe_kernel<ILP>, blocks, threads, 0, st, ctx->out96.as<uint4>(), 50);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(ctx->ev[12], st));
        LAUNCH(*ctx, shoup_probe_kernel<ILP>, blocks, threads, 0, st, ctx->out96.as<uint4>(), iters);
        CK(cudaEventRecord(ctx->ev[13], st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        const double rate = (double)blocks * threads * (double)iters * ILP / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    if (muls_per_s) *muls_per_s = best;
    return B2_OK;
}

int b2_dfma_probe(double* dfma_per_s) {
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->out96.reserve(96))) return rc;
    cudaStream_t st = ctx->stream;
    const int iters = 20000, ILP = 8;
    const int blocks = ctx->sms * 8, threads = 256;
    LAUNCH(*ctx, dfma_probe_kernel<ILP>, blocks, threads, 0, st, ctx->out96.as<double>(), 100);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(ctx->ev[12], st));
        LAUNCH(*ctx, dfma_probe_kernel<ILP>, blocks, threads, 0, st, ctx->out96.as<double>(), iters);
        CK(cudaEventRecord(ctx->ev[13], st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        double rate = (double)blocks * threads * (double)iters * ILP / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    *dfma_per_s = best;
    return B2_OK;
}

int b2_pipe_probe(int kind, double* macs_per_s) {
    if (!macs_per_s || kind < 0 || kind > 1) return fail(B2_ERR_ARG, "pipe_probe: kind 0..1");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->out96.reserve(96))) return rc;
    if (kind == 0) return run_pipe_probe<0>(ctx, macs_per_s);
    return run_pipe_probe<1>(ctx, macs_per_s);
}

int b2_mixed_probe(uint32_t imad_mask, uint32_t dfma_mask, int int_kind, int iters_int, int iters_f64, double* ms_out) {
    if (!ms_out || (imad_mask & dfma_mask) || int_kind < 0 || int_kind > 1 || iters_int < 0 || iters_f64 < 0)
        return fail(B2_ERR_ARG, "mixed_probe: disjoint warp masks (8 warps per block), int_kind 0..1");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    if ((rc = ctx->out96.reserve(96))) return rc;
    cudaStream_t st = ctx->stream;
    const int blocks = ctx->sms * 2, threads = 256;   // 2 blocks x 8 warps per SM = 4 warps per scheduler
    LAUNCH(*ctx, mixed_probe_kernel<2>, blocks, threads, 0, st, ctx->out96.as<uint4>(), 10, 10, imad_mask, dfma_mask,
           int_kind);
    double best = 1e30;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(ctx->ev[12], st));
        LAUNCH(*ctx, mixed_probe_kernel<2>, blocks, threads, 0, st, ctx->out96.as<uint4>(), iters_int, iters_f64,
               imad_mask, dfma_mask, int_kind);
        CK(cudaEventRecord(ctx->ev[13], st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        if (ms < best) best = ms;
    }
    *ms_out = best;
    return B2_OK;
}

int b2_affine_batch_probe(size_t n_pairs, uint32_t B, double* batch_ms, double* xyzz_ms, void* host_p, void* host_q,
                          void* host_out, size_t n_copy) {
    if (!batch_ms || !xyzz_ms || B == 0 || n_pairs == 0 || n_pairs % B || n_copy > n_pairs)
        return fail(B2_ERR_ARG, "affine_batch_probe: n_pairs must be a multiple of B");
    LaneLock ll;
    int rc = ll.acquire();
    if (rc) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = ctx->stream;
    const uint32_t nthreads = (uint32_t)(n_pairs / B);
    char *P = nullptr, *Q = nullptr, *out = nullptr, *xo = nullptr;
    uint4* prefix = nullptr;
    auto cleanup = [&]() { cudaFree(P); cudaFree(Q); cudaFree(out); cudaFree(xo); cudaFree(prefix); };
    if (cudaMalloc(&P, n_pairs * 64) != cudaSuccess || cudaMalloc(&Q, n_pairs * 64) != cudaSuccess ||
        cudaMalloc(&out, n_pairs * 64) != cudaSuccess || cudaMalloc(&xo, (size_t)nthreads * 128) != cudaSuccess ||
        cudaMalloc(&prefix, n_pairs * 32) != cudaSuccess) {
        cleanup();
        cudaGetLastError();
        return fail(B2_ERR_OOM, "affine_batch_probe: device allocation");
    }
    const unsigned gp = (unsigned)((n_pairs + 127) / 128), gt = (nthreads + 127) / 128;
    LAUNCH(*ctx, srs_synth_kernel, gp, 128, 0, st, P, (unsigned long long)n_pairs, 0ull, 0x1111ull);
    LAUNCH(*ctx, srs_synth_kernel, gp, 128, 0, st, Q, (unsigned long long)n_pairs, 0ull, 0x2222ull);
    double best_b = 1e30, best_x = 1e30;
    for (int rep = 0; rep < 3; rep++) {
        float ms = 0;
        CK(cudaEventRecord(ctx->ev[12], st));
        LAUNCH(*ctx, affine_batch_probe_kernel, gt, 128, 0, st, P, Q, out, prefix, B, nthreads);
        CK(cudaEventRecord(ctx->ev[13], st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        if (ms < best_b) best_b = ms;
        CK(cudaEventRecord(ctx->ev[12], st));
        LAUNCH(*ctx, xyzz_pair_probe_kernel, gt, 128, 0, st, P, Q, xo, B, nthreads);
        CK(cudaEventRecord(ctx->ev[13], st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
        if (ms < best_x) best_x = ms;
    }
    *batch_ms = best_b;
    *xyzz_ms = best_x;
    cudaError_t e = cudaSuccess;
    if (n_copy && host_p) e = cudaMemcpy(host_p, P, n_copy * 64, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && n_copy && host_q) e = cudaMemcpy(host_q, Q, n_copy * 64, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && n_copy && host_out) e = cudaMemcpy(host_out, out, n_copy * 64, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return fail(B2_ERR_CUDA, "affine_batch_probe: %s", cudaGetErrorString(e));
    return B2_OK;
}

int b2_last_timing(double* kernel_ms, double* total_ms) {
    // of the last host-pointer call made by THIS thread
    if (kernel_ms) *kernel_ms = g_last.kernel_ms;
    if (total_ms) *total_ms = g_last.total_ms;
    return B2_OK;
}
int b2_last_msm_phases(double* phases) {
    // CUDA-event times of the last MSM this thread ran (valid until another call reuses its lane)
    Lane* ln = g_last.lane;
    if (!ln) return fail(B2_ERR_ARG, "no MSM has run on this thread");
    int rc = msm_collect_phases(*ln);
    if (rc) return rc;
    for (int i = 0; i < 8; i++) phases[i] = ln->phases[i];
    return B2_OK;
}

}  // extern "C"

#include "api_quotient.inl"
#include "api_scan.inl"
#include "api_encoding.inl"
#include "api_lookup.inl"
