// common.cuh -- shared host/device definitions for libpavgpu (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "pavgpu.h"

void pav_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            pav_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #expr); \
            return PAVGPU_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)

// Device memory pool of a context. A Snakemake job calls the library many times with similar sizes; cudaMalloc and
// above all cudaFree of ~100 MB buffers cost milliseconds to hundreds of milliseconds each and synchronise the device,
// which made the host-buffer entry points slow and erratic. Blocks are handed back to the pool instead of the driver
// and reused best-fit; everything runs on the context's one stream and every entry point synchronises it before
// returning, so a freed block is never still in use by the device.
struct PavDevBlock {
    void *ptr;
    size_t bytes;
    bool used;
};

struct pavgpu_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    cudaEvent_t ev[8];
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    int flush_val = 0;
    std::vector<PavDevBlock> pool;
    size_t pool_free_bytes = 0;
    size_t pool_max_free = (size_t)32 << 30;   // PAVGPU_POOL_MAX_MB
    // pinned host staging for results (see pav_pinned_take)
    std::vector<PavDevBlock> pinned;
};

// Device buffer of at least `bytes` (never nullptr on success, even for bytes == 0). *got (optional) = its real size.
static inline cudaError_t pav_dev_alloc(pavgpu_ctx *ctx, size_t bytes, void **out, size_t *got = nullptr)
{
    size_t need = bytes < 512 ? 512 : bytes;
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); i++) {
        const PavDevBlock &b = ctx->pool[i];
        if (!b.used && b.bytes >= need && b.bytes / 4 <= need + ((size_t)1 << 20) && (best < 0 || b.bytes < ctx->pool[best].bytes)) best = (int)i;
    }
    if (best >= 0) {
        ctx->pool[best].used = true;
        ctx->pool_free_bytes -= ctx->pool[best].bytes;
        *out = ctx->pool[best].ptr;
        if (got) *got = ctx->pool[best].bytes;
        return cudaSuccess;
    }
    // new block, with some slack so that the next, slightly larger request still fits
    size_t sz = need < ((size_t)1 << 20) ? (need + 4095) / 4096 * 4096 : (need + need / 8 + ((size_t)1 << 20) - 1) >> 20 << 20;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {   // give everything cached back to the driver and retry once with the exact size
        (void)cudaGetLastError();
        for (size_t i = 0; i < ctx->pool.size();) {
            if (!ctx->pool[i].used) { cudaFree(ctx->pool[i].ptr); ctx->pool[i] = ctx->pool.back(); ctx->pool.pop_back(); }
            else i++;
        }
        ctx->pool_free_bytes = 0;
        sz = (need + 511) / 512 * 512;
        e = cudaMalloc(&p, sz);
        if (e != cudaSuccess) return e;
    }
    ctx->pool.push_back(PavDevBlock{p, sz, true});
    *out = p;
    if (got) *got = sz;
    return cudaSuccess;
}

static inline void pav_dev_free(pavgpu_ctx *ctx, void *ptr)
{
    if (!ptr) return;
    for (size_t i = 0; i < ctx->pool.size(); i++) {
        if (ctx->pool[i].ptr != ptr) continue;
        if (ctx->pool_free_bytes + ctx->pool[i].bytes > ctx->pool_max_free) {
            cudaFree(ptr);
            ctx->pool[i] = ctx->pool.back();
            ctx->pool.pop_back();
        } else {
            ctx->pool[i].used = false;
            ctx->pool_free_bytes += ctx->pool[i].bytes;
        }
        return;
    }
    cudaFree(ptr);   // not from the pool
}

template <typename T>
static inline cudaError_t pav_dev_alloc_t(pavgpu_ctx *ctx, size_t count, T **out)
{
    void *p = nullptr;
    cudaError_t e = pav_dev_alloc(ctx, count * sizeof(T), &p);
    *out = reinterpret_cast<T *>(p);
    return e;
}

// Scratch arenas (ASCII staging of the pack kernel, the density batch arena): same pool.
static inline cudaError_t ctx_arena_take(pavgpu_ctx *ctx, size_t bytes, void **out, size_t *got) { return pav_dev_alloc(ctx, bytes, out, got); }
static inline void ctx_arena_give(pavgpu_ctx *ctx, void *ptr, size_t) { pav_dev_free(ctx, ptr); }

// Pinned host buffers for results (library-allocated *_out arrays): D2H copies into pinned memory run at PCIe speed and
// need no first-touch page faults. pavgpu_free_host() recognises them and hands them back to the context they came from.
int pav_pinned_take(pavgpu_ctx *ctx, size_t bytes, void **out);
bool pav_pinned_give(void *ptr);   // false when ptr is not a pinned result buffer

// PAVGPU_TRACE=1: wall-clock phase timings of the host-buffer entry points on stderr.
struct PavTrace {
    bool on;
    const char *name;
    double t0, last;
    explicit PavTrace(const char *n);
    void mark(const char *label);
    ~PavTrace();
};

// Sequence planes in HBM.
//   pack2 : 64-bit words, 32 bases per word, base g at bits [62 - 2*(g%32), +2)  (first base most
//           significant, so a k-mer is a funnel-shifted window of the plane)
//   nmask : 32-bit words, bit (g%32) set when base g is not one of ACGTacgt
// Every sequence starts at a base offset that is a multiple of SEQ_ALIGN; padding bases have mask=1.
constexpr int64_t SEQ_ALIGN = 128;

struct pavgpu_seqstore {
    pavgpu_ctx *ctx;
    uint64_t uid;         // unique per store in this process (never 0): lets batches cache what they derived from a store
    int32_t n_seq;
    int64_t total_bases;  // padded
    std::vector<int64_t> h_off, h_len;
    int64_t *d_off;
    int64_t *d_len;
    uint64_t *d_pack2;
    uint32_t *d_nmask;
    uint32_t *d_nsum;     // N summary: bit b of word j set when mask words [(32 j + b) * 8, +8) hold a set bit (seqbits.cuh nsum_any2)
    size_t pack2_bytes, nmask_bytes, nsum_bytes;
};

// (Re)build the N summary from the mask plane on the store's stream: after packing, after an upload of packed planes, after a broadcast.
int pav_build_nsum(pavgpu_seqstore *store);

struct SeqPlanes {
    const uint64_t *pack2;
    const uint32_t *nmask;
    const int64_t *off;
    const int64_t *len;
    const uint32_t *nsum;
};

static inline SeqPlanes planes_of(const pavgpu_seqstore *s)
{
    return SeqPlanes{s->d_pack2, s->d_nmask, s->d_off, s->d_len, s->d_nsum};
}

#include "seqbits.cuh"

static inline float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
