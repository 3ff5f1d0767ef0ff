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
    size_t pack2_bytes, nmask_bytes;
};

struct SeqPlanes {
    const uint64_t *pack2;
    const uint32_t *nmask;
    const int64_t *off;
    const int64_t *len;
};

static inline SeqPlanes planes_of(const pavgpu_seqstore *s)
{
    return SeqPlanes{s->d_pack2, s->d_nmask, s->d_off, s->d_len};
}

// A sequence seen in alignment orientation: position t maps to the forward base t, or to the
// complement of forward base len-1-t when rev (what Bio.Seq.reverse_complement materialises in
// pavlib/cigarcall.py:69-70; here it is index arithmetic).
struct OSeq {
    const uint64_t *pack2;
    const uint32_t *nmask;
    int64_t base;  // offset of the sequence in the planes
    int64_t len;
    int rev;
    // Optional staged copy of plane words [t_w0, t_w0 + t_nw1 + 1) in shared memory (homology_tiled_kernel); windows whose two
    // words lie inside are served from it, all others from global memory. t_nw1 == 0: no tile.
    const uint64_t *t_pack2;
    const uint32_t *t_nmask;
    int64_t t_w0;
    int32_t t_nw1;
};

// Upper-cased base as 0..3 (ACGT) or 4 (anything else, or out of range).
__device__ __forceinline__ int oseq_base(const OSeq &s, int64_t t)
{
    if (t < 0 || t >= s.len) return 4;
    int64_t g = s.base + (s.rev ? (s.len - 1 - t) : t);
    uint32_t m = (__ldg(s.nmask + (g >> 5)) >> (g & 31)) & 1u;
    if (m) return 4;
    int c = (int)((__ldg(s.pack2 + (g >> 5)) >> (62 - 2 * (int)(g & 31))) & 3ull);
    return s.rev ? 3 - c : c;
}

// ---- 32-base windows --------------------------------------------------------------------------
// Reverse the 32 two-bit groups of a word and complement them (reverse complement of 32 bases).
__device__ __forceinline__ uint64_t revcomp32(uint64_t x)
{
    x = ~x;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    return ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | (uint64_t)__byte_perm((uint32_t)(x >> 32), 0, 0x0123);
}

// Forward-strand window: bases f .. f+31 of a sequence (base i of the window in bits [62-2i, 64-2i) of
// `bases`, bit i of `mask` set when that base is not ACGT or lies outside [0, len)). Positions inside a sequence
// are 32-bit (sequences are shorter than 2^31, checked when the store is built); only the plane offset is 64-bit.
template <bool TILED = false>
__device__ __forceinline__ void fwd_window(const OSeq &s, int32_t f, uint64_t &bases, uint32_t &mask)
{
    const int32_t len = (int32_t)s.len;
    if (f <= -32 || f >= len) { bases = 0; mask = 0xffffffffu; return; }
    const int lead = f < 0 ? -f : 0;           // window positions before the sequence start
    const int64_t g = s.base + (int64_t)(f + lead);
    const int64_t w = g >> 5;
    const int sh = (int)(g & 31);
    uint64_t hi, lo;
    uint32_t m0, m1;
    if (TILED && (uint64_t)(w - s.t_w0) < (uint64_t)s.t_nw1) {   // words w and w+1 are staged
        const int o = (int)(w - s.t_w0);
        hi = s.t_pack2[o]; lo = s.t_pack2[o + 1];
        m0 = s.t_nmask[o]; m1 = s.t_nmask[o + 1];
    } else {
        hi = __ldg(s.pack2 + w); lo = __ldg(s.pack2 + w + 1);
        m0 = __ldg(s.nmask + w); m1 = __ldg(s.nmask + w + 1);
    }
    uint64_t b = sh ? ((hi << (2 * sh)) | (lo >> (64 - 2 * sh))) : hi;
    uint32_t m = __funnelshift_r(m0, m1, sh);
    if (lead | (f + 32 > len)) {               // sequence edges only
        if (lead) { b >>= 2 * lead; m = (m << lead) | ((1u << lead) - 1u); }
        const int over = f + 32 - len;         // window positions past the sequence end
        if (over > 0) m |= ~0u << (32 - over);
    }
    bases = b; mask = m;
}

// Window of 32 bases starting at oriented position t (reverse-complement view when s.rev).
// (An out-of-line variant of this and of dev_homology_raw was measured on B200: 0.176 ms vs 0.148 ms inlined for the
// C2 homology kernel -- call overhead and spills cost more than the instruction-fetch stalls they remove. 16-base
// windows with 32-bit funnel shifts were measured too: 0.147 ms vs 0.102 ms, twice the loop trips for long scans.)
template <bool TILED = false>
__device__ __forceinline__ void oseq_window(const OSeq &s, int32_t t, uint64_t &bases, uint32_t &mask)
{
    if (!s.rev) { fwd_window<TILED>(s, t, bases, mask); return; }
    uint64_t b; uint32_t m;
    fwd_window<TILED>(s, (int32_t)s.len - t - 32, b, m);
    bases = revcomp32(b);
    mask = __brev(m);
}

// Longest common extension of A from a and B from b, 32 bases per step, capped at `limit`:
//   left == 0: common prefix of A[a..] and B[b..]        (window i covers a+32i .. a+32i+31)
//   left != 0: common suffix of A[..a] and B[..b]         (window i covers a-32i-31 .. a-32i)
// Stops at the first mismatch, non-ACGT base or sequence end on either side.
template <bool TILED = false>
__device__ __forceinline__ int32_t common_extension(const OSeq &A, int32_t a, const OSeq &B, int32_t b, int32_t limit, int left)
{
    int32_t h = 0;
    const int32_t a0 = left ? a - 31 : a, b0 = left ? b - 31 : b, step = left ? -32 : 32;
    int32_t pa = a0, pb = b0;
    while (h < limit) {
        uint64_t wa, wb; uint32_t ma, mb;
        oseq_window<TILED>(A, pa, wa, ma);
        oseq_window<TILED>(B, pb, wb, mb);
        uint64_t x = wa ^ wb;
        uint64_t d = (x | (x >> 1)) & 0x5555555555555555ull;      // one bit per differing base
        uint32_t m = ma | mb;
        int stop_d, stop_m;
        if (left) {   // last base of the window = least significant group / highest mask bit
            stop_d = d ? ((__ffsll((long long)d) - 1) >> 1) : 32;
            stop_m = m ? __clz((int)m) : 32;
        } else {      // first base = most significant group / lowest mask bit
            stop_d = d ? (__clzll((long long)d) >> 1) : 32;
            stop_m = m ? (__ffs((int)m) - 1) : 32;
        }
        int stop = min(stop_d, stop_m);
        if (stop < 32) { h += stop; return h < limit ? h : limit; }
        h += 32; pa += step; pb += step;
        if (h < 0) return limit;   // (cannot happen for sequences < 2^31; guards the 32-bit counter)
    }
    return limit;
}

// pavlib/call.py:542-592 (left != 0) and :595-647 (left == 0). T: flank searched from p away from the
// breakpoint; the SV sequence is V[v0 : v0+n], read circularly (leftwards from its end: sv[-((h+1) % n)],
// index -0 == 0; rightwards from its start: sv[h % n]). Word-parallel form: the first n steps are a common
// suffix/prefix of the flank with V; once a whole copy of V matched, step h compares T[p -/+ h] with
// V[...] = T[p -/+ h +/- n], i.e. the scan continues as the common extension of the flank with itself
// shifted by n.
static __device__ __forceinline__ int dev_homology_raw(const uint64_t *t_pack2, const uint32_t *t_nmask, int64_t t_base, int64_t t_len, int t_rev,
                                                    int64_t p, const uint64_t *v_pack2, const uint32_t *v_nmask, int64_t v_base, int64_t v_len,
                                                    int v_rev, int64_t v0, int n, int left)
{
    const OSeq T{t_pack2, t_nmask, t_base, t_len, t_rev, nullptr, nullptr, 0, 0};
    const OSeq V{v_pack2, v_nmask, v_base, v_len, v_rev, nullptr, nullptr, 0, 0};
    if (n <= 0 || p < 0 || p >= T.len) return 0;
    const int32_t p32 = (int32_t)p, v32 = (int32_t)v0;
    int32_t h = common_extension(T, p32, V, left ? v32 + n - 1 : v32, n, left);
    if (h < n) return h;
    // the flank is shorter than 2^31, so the self-comparison ends at a sequence edge long before the cap
    return n + common_extension(T, left ? p32 - n : p32 + n, T, p32, 0x7fffffff - n, left);
}

// The same scan over sequences that may carry a staged tile.
static __device__ __forceinline__ int dev_homology_tiled(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n, int left)
{
    if (n <= 0 || p < 0 || p >= T.len) return 0;
    const int32_t p32 = (int32_t)p, v32 = (int32_t)v0;
    int32_t h = common_extension<true>(T, p32, V, left ? v32 + n - 1 : v32, n, left);
    if (h < n) return h;
    return n + common_extension<true>(T, left ? p32 - n : p32 + n, T, p32, 0x7fffffff - n, left);
}

__device__ __forceinline__ int dev_left_homology(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n)
{
    return dev_homology_raw(T.pack2, T.nmask, T.base, T.len, T.rev, p, V.pack2, V.nmask, V.base, V.len, V.rev, v0, n, 1);
}

__device__ __forceinline__ int dev_right_homology(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n)
{
    return dev_homology_raw(T.pack2, T.nmask, T.base, T.len, T.rev, p, V.pack2, V.nmask, V.base, V.len, V.rev, v0, n, 0);
}

static inline float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
