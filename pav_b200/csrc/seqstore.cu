// seqstore.cu -- context management and the packed sequence store (2-bit plane + N-mask plane).
//
// Replaces, for the hot path, what the reference does per alignment record on the CPU:
// pysam fetch of whole chromosomes/contigs, Bio reverse_complement and str.upper()
// (pavlib/cigarcall.py:58-75) and the per-base dict lookups of kanapy's k-mer stream
// (dep/svpop/dep/kanapy/util/kmer.py:50-69,206-221).
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include <algorithm>
#include <thread>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void pav_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" __attribute__((visibility("default"))) const char *pavgpu_last_error(void) { return g_err; }

extern "C" __attribute__((visibility("default"))) int pavgpu_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        pav_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return PAVGPU_ERR_CUDA;
    }
    return n;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_ctx_create(int device, pavgpu_ctx **ctx_out)
{
    if (!ctx_out) { pav_set_error("ctx_out is NULL"); return PAVGPU_ERR_ARG; }
    *ctx_out = nullptr;
    CUDA_TRY(cudaSetDevice(device));
    pavgpu_ctx *c = new pavgpu_ctx();
    c->device = device;
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    if (const char *mx = getenv("PAVGPU_POOL_MAX_MB")) c->pool_max_free = (size_t)strtoull(mx, nullptr, 10) << 20;
    // PAVGPU_L2_FETCH=32|64|128 (tuning, off by default): hint for the granularity at which L2 fetches from DRAM. The homology
    // gathers miss L1 on 3.8 M sectors but L2 is asked for 6.3 M and DRAM delivers 4.3 M (ncu, C2): neighbours of scattered
    // sectors are fetched along; 32 asks the device not to. A hint only -- errors are ignored.
    if (const char *fg = getenv("PAVGPU_L2_FETCH")) {
        const size_t v = (size_t)strtoull(fg, nullptr, 10);
        if (v == 32 || v == 64 || v == 128) { (void)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, v); (void)cudaGetLastError(); }
    }
    *ctx_out = c;
    return PAVGPU_OK;
}

// ---- pinned result buffers ----------------------------------------------------------------------
static std::mutex g_pinned_mu;
static std::unordered_map<void *, pavgpu_ctx *> g_pinned_owner;   // in-use pinned blocks -> owning context (nullptr: context gone)

int pav_pinned_take(pavgpu_ctx *ctx, size_t bytes, void **out)
{
    size_t need = bytes < 4096 ? 4096 : bytes;
    int best = -1;
    for (size_t i = 0; i < ctx->pinned.size(); i++) {
        const PavDevBlock &b = ctx->pinned[i];
        if (!b.used && b.bytes >= need && b.bytes / 4 <= need + ((size_t)1 << 20) && (best < 0 || b.bytes < ctx->pinned[best].bytes)) best = (int)i;
    }
    if (best < 0) {
        size_t sz = need < ((size_t)1 << 20) ? (need + 4095) / 4096 * 4096 : (need + need / 8 + ((size_t)1 << 20) - 1) >> 20 << 20;
        void *p = nullptr;
        cudaError_t e = cudaHostAlloc(&p, sz, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            pav_set_error("cudaHostAlloc(%zu) failed: %s", sz, cudaGetErrorString(e));
            return PAVGPU_ERR_NOMEM;
        }
        ctx->pinned.push_back(PavDevBlock{p, sz, false});
        best = (int)ctx->pinned.size() - 1;
    }
    ctx->pinned[best].used = true;
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        g_pinned_owner[ctx->pinned[best].ptr] = ctx;
    }
    *out = ctx->pinned[best].ptr;
    return PAVGPU_OK;
}

bool pav_pinned_give(void *ptr)
{
    pavgpu_ctx *ctx = nullptr;
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        auto it = g_pinned_owner.find(ptr);
        if (it == g_pinned_owner.end()) return false;
        ctx = it->second;
        g_pinned_owner.erase(it);
    }
    if (!ctx) { cudaFreeHost(ptr); return true; }   // the context was destroyed while the caller still held the buffer
    size_t free_bytes = 0;
    for (auto &b : ctx->pinned) if (!b.used) free_bytes += b.bytes;
    for (size_t i = 0; i < ctx->pinned.size(); i++) {
        if (ctx->pinned[i].ptr != ptr) continue;
        if (free_bytes + ctx->pinned[i].bytes > ((size_t)8 << 30)) {   // keep at most 8 GiB of idle pinned memory
            cudaFreeHost(ptr);
            ctx->pinned[i] = ctx->pinned.back();
            ctx->pinned.pop_back();
        } else ctx->pinned[i].used = false;
        break;
    }
    return true;
}

// ---- trace ---------------------------------------------------------------------------------------
static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

PavTrace::PavTrace(const char *n) : name(n)
{
    static int enabled = -1;
    if (enabled < 0) { const char *e = getenv("PAVGPU_TRACE"); enabled = (e && e[0] == '1') ? 1 : 0; }
    on = enabled == 1;
    if (on) t0 = last = now_ms();
}

void PavTrace::mark(const char *label)
{
    if (!on) return;
    double t = now_ms();
    fprintf(stderr, "[pavgpu trace] %-28s %-26s %9.3f ms\n", name, label, t - last);
    last = t;
}

PavTrace::~PavTrace()
{
    if (on) fprintf(stderr, "[pavgpu trace] %-28s %-26s %9.3f ms\n", name, "TOTAL", now_ms() - t0);
}

extern "C" __attribute__((visibility("default"))) void pavgpu_ctx_destroy(pavgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &e : c->ev) cudaEventDestroy(e);
    cudaFree(c->flush_buf);
    for (auto &b : c->pool) cudaFree(b.ptr);   // blocks still marked used belong to objects the caller leaked; the memory goes with the context
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        for (auto &b : c->pinned) {
            if (!b.used) cudaFreeHost(b.ptr);
            else g_pinned_owner[b.ptr] = nullptr;   // still held by the caller: released by pavgpu_free_host later
        }
    }
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_ctx_device(const pavgpu_ctx *c) { return c ? c->device : -1; }
extern "C" __attribute__((visibility("default"))) void pavgpu_free_host(void *p)
{
    if (p && !pav_pinned_give(p)) free(p);
}

extern "C" __attribute__((visibility("default"))) int pavgpu_host_alloc(pavgpu_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) { pav_set_error("host_alloc: bad argument"); return PAVGPU_ERR_ARG; }
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return pav_pinned_take(ctx, bytes, out);
}

extern "C" __attribute__((visibility("default"))) int pavgpu_l2_flush(pavgpu_ctx *ctx, size_t bytes)
{
    if (!ctx || bytes == 0) { pav_set_error("l2_flush: bad argument"); return PAVGPU_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->flush_bytes < bytes) {
        cudaFree(ctx->flush_buf);
        ctx->flush_buf = nullptr;
        ctx->flush_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->flush_buf, bytes));
        ctx->flush_bytes = bytes;
    }
    CUDA_TRY(cudaMemsetAsync(ctx->flush_buf, ++ctx->flush_val, bytes, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PAVGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// pack kernel: ASCII -> (2-bit plane, N-mask plane). One thread per 32 bases = one 64-bit plane word
// and one 32-bit mask word; two 128-bit loads per thread. Pure streaming: 1 B/base in, 0.375 B/base out.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ ascii, int64_t n_words,
                                                   uint64_t *__restrict__ pack2, uint32_t *__restrict__ nmask)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(ascii + w * 32);
    uint4 a = __ldcs(src), b = __ldcs(src + 1);
    uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint64_t word;
    uint32_t mask;
    pack_word32(v, word, mask);
    pack2[w] = word;
    nmask[w] = mask;
}

// N summary of the mask plane: a warp covers 32 consecutive mask words = 4 summary bits of one summary word.
__global__ void __launch_bounds__(256) nsum_kernel(const uint32_t *__restrict__ nmask, int64_t n_words, uint32_t *__restrict__ nsum)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned bal = __ballot_sync(0xffffffffu, w < n_words && nmask[w] != 0u);
    if ((threadIdx.x & 31) == 0 && bal) {
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) bits |= ((bal >> (8 * q)) & 0xffu) ? (1u << q) : 0u;
        atomicOr(nsum + (w >> 8), bits << ((w >> 3) & 31));
    }
}

int pav_build_nsum(pavgpu_seqstore *s)
{
    pavgpu_ctx *ctx = s->ctx;
    const int64_t n_words = s->total_bases / 32;
    CUDA_TRY(cudaMemsetAsync(s->d_nsum, 0, s->nsum_bytes, ctx->stream));
    nsum_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, ctx->stream>>>(s->d_nmask, n_words, s->d_nsum);
    CUDA_TRY(cudaGetLastError());
    return PAVGPU_OK;
}

static int alloc_store(pavgpu_ctx *ctx, int32_t n_seq, const int64_t *seq_len, pavgpu_seqstore **out)
{
    if (!ctx || n_seq < 0 || (n_seq > 0 && !seq_len) || !out) { pav_set_error("seqstore: bad argument"); return PAVGPU_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    static std::atomic<uint64_t> next_uid{1};
    pavgpu_seqstore *s = new pavgpu_seqstore();
    s->ctx = ctx;
    s->uid = next_uid.fetch_add(1);
    s->n_seq = n_seq;
    s->h_off.resize(n_seq);
    s->h_len.assign(seq_len, seq_len + n_seq);
    int64_t off = 0;
    for (int32_t i = 0; i < n_seq; i++) {
        if (seq_len[i] < 0 || seq_len[i] >= (int64_t)1 << 31) {
            pav_set_error("seqstore: sequence %d has unsupported length %lld", i, (long long)seq_len[i]);
            delete s;
            return PAVGPU_ERR_ARG;
        }
        s->h_off[i] = off;
        off += (seq_len[i] + SEQ_ALIGN - 1) / SEQ_ALIGN * SEQ_ALIGN;
    }
    off += SEQ_ALIGN;  // tail guard so k-mer windows may read one word past the last base
    s->total_bases = off;
    s->pack2_bytes = (size_t)(off / 32) * 8;
    s->nmask_bytes = (size_t)(off / 32) * 4;
    s->nsum_bytes = ((size_t)(off / 32 + 1) / 256 + 2) * 4;   // windows read mask words w and w+1: one summary word of slack
    s->d_off = s->d_len = nullptr;
    s->d_pack2 = nullptr;
    s->d_nmask = nullptr;
    s->d_nsum = nullptr;
    cudaError_t e;
    if ((e = pav_dev_alloc_t(ctx, s->pack2_bytes / 8, &s->d_pack2)) != cudaSuccess || (e = pav_dev_alloc_t(ctx, s->nmask_bytes / 4, &s->d_nmask)) != cudaSuccess ||
        (e = pav_dev_alloc_t(ctx, s->nsum_bytes / 4, &s->d_nsum)) != cudaSuccess ||
        (e = pav_dev_alloc_t(ctx, (size_t)n_seq + 1, &s->d_off)) != cudaSuccess ||
        (e = pav_dev_alloc_t(ctx, (size_t)n_seq + 1, &s->d_len)) != cudaSuccess) {
        pav_set_error("seqstore: cudaMalloc failed: %s", cudaGetErrorString(e));
        pavgpu_seqstore_free(s);
        return PAVGPU_ERR_NOMEM;
    }
    int rc = [&]() -> int {
        if (n_seq) {
            CUDA_TRY(cudaMemcpyAsync(s->d_off, s->h_off.data(), sizeof(int64_t) * n_seq, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(s->d_len, s->h_len.data(), sizeof(int64_t) * n_seq, cudaMemcpyHostToDevice, ctx->stream));
        }
        CUDA_TRY(cudaMemsetAsync(s->d_nsum, 0xFF, s->nsum_bytes, ctx->stream));   // "every block may hold N" until the planes are filled
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return PAVGPU_OK;
    }();
    if (rc) { pavgpu_seqstore_free(s); return rc; }
    *out = s;
    return PAVGPU_OK;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_seqstore_create_empty(pavgpu_ctx *ctx, int32_t n_seq, const int64_t *seq_len, pavgpu_seqstore **out)
{
    return alloc_store(ctx, n_seq, seq_len, out);
}

extern "C" __attribute__((visibility("default"))) int pavgpu_seqstore_create(pavgpu_ctx *ctx, int32_t n_seq, const uint8_t *const *seq_ascii, const int64_t *seq_len,
                                      pavgpu_seqstore **out)
{
    if (n_seq > 0 && !seq_ascii) { pav_set_error("seqstore: seq_ascii is NULL"); return PAVGPU_ERR_ARG; }
    PavTrace tr("seqstore_create");
    pavgpu_seqstore *s = nullptr;
    int rc = alloc_store(ctx, n_seq, seq_len, &s);
    if (rc) return rc;
    tr.mark("alloc planes");
    // Stage ASCII in HBM ('N' in the padding so that padding bases get mask = 1), then pack.
    uint8_t *d_ascii = nullptr;
    size_t ascii_cap = 0;
    cudaError_t e = ctx_arena_take(ctx, (size_t)s->total_bases, reinterpret_cast<void **>(&d_ascii), &ascii_cap);
    if (e != cudaSuccess) {
        pav_set_error("seqstore: cudaMalloc(%lld) for ASCII staging failed: %s", (long long)s->total_bases, cudaGetErrorString(e));
        pavgpu_seqstore_free(s);
        return PAVGPU_ERR_NOMEM;
    }
    rc = [&]() -> int {
        // Many small sequences (a batch of density windows): one pageable cudaMemcpyAsync each costs ~10 us of driver staging per
        // call. They are laid out -- padding included -- in one pinned buffer by a few host threads instead and go up in one copy.
        const bool many_small = n_seq >= 16 && s->total_bases <= ((int64_t)256 << 20) && s->total_bases / n_seq <= ((int64_t)1 << 20);
        uint8_t *stage = nullptr;
        if (many_small && pav_pinned_take(ctx, (size_t)s->total_bases, reinterpret_cast<void **>(&stage)) != PAVGPU_OK) stage = nullptr;
        struct StageGuard {   // back to the pool once the stream is done with it, on every way out
            uint8_t *p; cudaStream_t st;
            ~StageGuard() { if (p) { cudaStreamSynchronize(st); pavgpu_free_host(p); } }
        } guard{stage, ctx->stream};
        if (stage) {
            const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)8, (int64_t)std::thread::hardware_concurrency(), s->total_bases >> 20}));
            auto fill = [&](int32_t lo, int32_t hi) {
                for (int32_t i = lo; i < hi; i++) {
                    const int64_t end = (i + 1 < n_seq) ? s->h_off[i + 1] : s->total_bases;
                    if (seq_len[i] > 0) memcpy(stage + s->h_off[i], seq_ascii[i], (size_t)seq_len[i]);
                    memset(stage + s->h_off[i] + seq_len[i], 'N', (size_t)(end - s->h_off[i] - seq_len[i]));
                }
            };
            if (n_seq == 0) memset(stage, 'N', (size_t)s->total_bases);
            std::vector<std::thread> pool;
            for (int t = 1; t < n_thr; t++) pool.emplace_back(fill, (int32_t)((int64_t)n_seq * t / n_thr), (int32_t)((int64_t)n_seq * (t + 1) / n_thr));
            fill(0, (int32_t)((int64_t)n_seq / n_thr));
            for (auto &th : pool) th.join();
            CUDA_TRY(cudaMemcpyAsync(d_ascii, stage, (size_t)s->total_bases, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            CUDA_TRY(cudaMemsetAsync(d_ascii, 'N', (size_t)s->total_bases, ctx->stream));
            for (int32_t i = 0; i < n_seq; i++)
                if (seq_len[i] > 0)
                    CUDA_TRY(cudaMemcpyAsync(d_ascii + s->h_off[i], seq_ascii[i], (size_t)seq_len[i], cudaMemcpyHostToDevice, ctx->stream));
        }
        int64_t n_words = s->total_bases / 32;
        int64_t blocks = (n_words + 255) / 256;
        pack_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_ascii, n_words, s->d_pack2, s->d_nmask);
        CUDA_TRY(cudaGetLastError());
        { int nrc = pav_build_nsum(s); if (nrc) return nrc; }
        tr.mark("enqueue h2d + pack");
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        tr.mark("sync");
        return PAVGPU_OK;
    }();
    ctx_arena_give(ctx, d_ascii, ascii_cap);
    if (rc) { pavgpu_seqstore_free(s); return rc; }
    *out = s;
    return PAVGPU_OK;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_seqstore_create_packed(pavgpu_ctx *ctx, int32_t n_seq, const int64_t *seq_len, const uint64_t *pack2_host,
                                             size_t pack2_bytes, const uint32_t *nmask_host, size_t nmask_bytes, pavgpu_seqstore **out)
{
    if (!pack2_host || !nmask_host) { pav_set_error("seqstore_create_packed: planes are NULL"); return PAVGPU_ERR_ARG; }
    pavgpu_seqstore *s = nullptr;
    int rc = alloc_store(ctx, n_seq, seq_len, &s);
    if (rc) return rc;
    if (pack2_bytes != s->pack2_bytes || nmask_bytes != s->nmask_bytes) {   // planes packed under another layout (alignment, tail guard) or truncated
        pav_set_error("seqstore_create_packed: plane sizes (%zu, %zu) do not match this library's layout for these lengths (%zu, %zu)", pack2_bytes,
                      nmask_bytes, s->pack2_bytes, s->nmask_bytes);
        pavgpu_seqstore_free(s);
        return PAVGPU_ERR_ARG;
    }
    rc = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(s->d_pack2, pack2_host, s->pack2_bytes, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_nmask, nmask_host, s->nmask_bytes, cudaMemcpyHostToDevice, ctx->stream));
        { int nrc = pav_build_nsum(s); if (nrc) return nrc; }
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return PAVGPU_OK;
    }();
    if (rc) { pavgpu_seqstore_free(s); return rc; }
    *out = s;
    return PAVGPU_OK;
}

extern "C" __attribute__((visibility("default"))) void pavgpu_seqstore_free(pavgpu_seqstore *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    pav_dev_free(s->ctx, s->d_pack2);
    pav_dev_free(s->ctx, s->d_nmask);
    pav_dev_free(s->ctx, s->d_nsum);
    pav_dev_free(s->ctx, s->d_off);
    pav_dev_free(s->ctx, s->d_len);
    delete s;
}

extern "C" __attribute__((visibility("default"))) int32_t pavgpu_seqstore_n_seq(const pavgpu_seqstore *s) { return s ? s->n_seq : -1; }
extern "C" __attribute__((visibility("default"))) int64_t pavgpu_seqstore_total_bases(const pavgpu_seqstore *s) { return s ? s->total_bases : -1; }
extern "C" __attribute__((visibility("default"))) int64_t pavgpu_seqstore_offset(const pavgpu_seqstore *s, int32_t i)
{
    return (s && i >= 0 && i < s->n_seq) ? s->h_off[i] : -1;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_seqstore_planes(const pavgpu_seqstore *s, void **d_pack2, size_t *pack2_bytes, void **d_nmask, size_t *nmask_bytes)
{
    if (!s) { pav_set_error("seqstore is NULL"); return PAVGPU_ERR_ARG; }
    if (d_pack2) *d_pack2 = s->d_pack2;
    if (pack2_bytes) *pack2_bytes = s->pack2_bytes;
    if (d_nmask) *d_nmask = s->d_nmask;
    if (nmask_bytes) *nmask_bytes = s->nmask_bytes;
    return PAVGPU_OK;
}

// Order-independent checksum of the two planes: sum of (word * odd constant + word index) over all words, per plane, mod 2^64.
__global__ void __launch_bounds__(256) plane_checksum_kernel(const uint64_t *__restrict__ pack2, const uint32_t *__restrict__ nmask, int64_t n_words,
                                                             unsigned long long *__restrict__ out)
{
    unsigned long long a = 0, b = 0;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
        a += (pack2[w] ^ (unsigned long long)w) * 0x9E3779B97F4A7C15ull;
        b += ((unsigned long long)nmask[w] ^ ((unsigned long long)w << 32)) * 0xC2B2AE3D27D4EB4Full;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
}

extern "C" __attribute__((visibility("default"))) int pavgpu_seqstore_checksum(const pavgpu_seqstore *s, uint64_t sums_out[2])
{
    if (!s || !sums_out) { pav_set_error("seqstore_checksum: bad argument"); return PAVGPU_ERR_ARG; }
    pavgpu_ctx *ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    CUDA_TRY(pav_dev_alloc_t(ctx, 2, &d));
    int rc = [&]() -> int {
        CUDA_TRY(cudaMemsetAsync(d, 0, 16, ctx->stream));
        const int64_t n_words = s->total_bases / 32;
        const int blocks = (int)std::min<int64_t>((n_words + 255) / 256, (int64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 16);
        plane_checksum_kernel<<<std::max(blocks, 1), 256, 0, ctx->stream>>>(s->d_pack2, s->d_nmask, n_words, d);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(sums_out, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return PAVGPU_OK;
    }();
    pav_dev_free(ctx, d);
    return rc;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_seqstore_export(const pavgpu_seqstore *s, uint64_t *pack2_host, uint32_t *nmask_host)
{
    if (!s) { pav_set_error("seqstore is NULL"); return PAVGPU_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    if (pack2_host) CUDA_TRY(cudaMemcpy(pack2_host, s->d_pack2, s->pack2_bytes, cudaMemcpyDeviceToHost));
    if (nmask_host) CUDA_TRY(cudaMemcpy(nmask_host, s->d_nmask, s->nmask_bytes, cudaMemcpyDeviceToHost));
    return PAVGPU_OK;
}
