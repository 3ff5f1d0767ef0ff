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

struct pavgpu_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    cudaEvent_t ev[8];
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    int flush_val = 0;
    // Scratch arenas handed back by finished batches / stores and reused by the next ones: a Snakemake job calls the
    // library many times with similar sizes, and cudaMalloc / cudaFree of ~GB buffers cost tens of milliseconds.
    static constexpr int N_CACHE = 4;
    void *cache_ptr[N_CACHE] = {nullptr, nullptr, nullptr, nullptr};
    size_t cache_bytes[N_CACHE] = {0, 0, 0, 0};
};

// Take a device buffer of at least `bytes` from the context cache (smallest fit) or allocate one. *got = its real size.
static inline cudaError_t ctx_arena_take(pavgpu_ctx *ctx, size_t bytes, void **out, size_t *got)
{
    int best = -1;
    for (int i = 0; i < pavgpu_ctx::N_CACHE; i++)
        if (ctx->cache_ptr[i] && ctx->cache_bytes[i] >= bytes && (best < 0 || ctx->cache_bytes[i] < ctx->cache_bytes[best])) best = i;
    if (best >= 0) {
        *out = ctx->cache_ptr[best]; *got = ctx->cache_bytes[best];
        ctx->cache_ptr[best] = nullptr; ctx->cache_bytes[best] = 0;
        return cudaSuccess;
    }
    *got = bytes;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {   // make room and retry once
        for (int i = 0; i < pavgpu_ctx::N_CACHE; i++) { cudaFree(ctx->cache_ptr[i]); ctx->cache_ptr[i] = nullptr; ctx->cache_bytes[i] = 0; }
        (void)cudaGetLastError();
        e = cudaMalloc(out, bytes);
    }
    return e;
}

// Give a buffer back: kept if a slot is free or it is larger than the smallest cached one (which is freed instead).
static inline void ctx_arena_give(pavgpu_ctx *ctx, void *ptr, size_t bytes)
{
    if (!ptr) return;
    int slot = -1, smallest = 0;
    for (int i = 0; i < pavgpu_ctx::N_CACHE; i++) {
        if (!ctx->cache_ptr[i]) { slot = i; break; }
        if (ctx->cache_bytes[i] < ctx->cache_bytes[smallest]) smallest = i;
    }
    if (slot < 0) {
        if (ctx->cache_bytes[smallest] >= bytes) { cudaFree(ptr); return; }
        cudaFree(ctx->cache_ptr[smallest]);
        slot = smallest;
    }
    ctx->cache_ptr[slot] = ptr; ctx->cache_bytes[slot] = bytes;
}

// Sequence planes in HBM.
//   pack2 : 64-bit words, 32 bases per word, base g at bits [62 - 2*(g%32), +2)  (first base most
//           significant, so a k-mer is a funnel-shifted window of the plane)
//   nmask : 32-bit words, bit (g%32) set when base g is not one of ACGTacgt
// Every sequence starts at a base offset that is a multiple of SEQ_ALIGN; padding bases have mask=1.
constexpr int64_t SEQ_ALIGN = 128;

struct pavgpu_seqstore {
    pavgpu_ctx *ctx;
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
__device__ __forceinline__ void fwd_window(const OSeq &s, int32_t f, uint64_t &bases, uint32_t &mask)
{
    const int32_t len = (int32_t)s.len;
    if (f <= -32 || f >= len) { bases = 0; mask = 0xffffffffu; return; }
    const int lead = f < 0 ? -f : 0;           // window positions before the sequence start
    const int64_t g = s.base + (int64_t)(f + lead);
    const int64_t w = g >> 5;
    const int sh = (int)(g & 31);
    const uint64_t hi = __ldg(s.pack2 + w), lo = __ldg(s.pack2 + w + 1);
    uint64_t b = sh ? ((hi << (2 * sh)) | (lo >> (64 - 2 * sh))) : hi;
    uint32_t m = __funnelshift_r(__ldg(s.nmask + w), __ldg(s.nmask + w + 1), sh);
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
__device__ __forceinline__ void oseq_window(const OSeq &s, int32_t t, uint64_t &bases, uint32_t &mask)
{
    if (!s.rev) { fwd_window(s, t, bases, mask); return; }
    uint64_t b; uint32_t m;
    fwd_window(s, (int32_t)s.len - t - 32, b, m);
    bases = revcomp32(b);
    mask = __brev(m);
}

// Longest common extension of A from a and B from b, 32 bases per step, capped at `limit`:
//   left == 0: common prefix of A[a..] and B[b..]        (window i covers a+32i .. a+32i+31)
//   left != 0: common suffix of A[..a] and B[..b]         (window i covers a-32i-31 .. a-32i)
// Stops at the first mismatch, non-ACGT base or sequence end on either side.
__device__ __forceinline__ int32_t common_extension(const OSeq &A, int32_t a, const OSeq &B, int32_t b, int32_t limit, int left)
{
    int32_t h = 0;
    const int32_t a0 = left ? a - 31 : a, b0 = left ? b - 31 : b, step = left ? -32 : 32;
    int32_t pa = a0, pb = b0;
    while (h < limit) {
        uint64_t wa, wb; uint32_t ma, mb;
        oseq_window(A, pa, wa, ma);
        oseq_window(B, pb, wb, mb);
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
    const OSeq T{t_pack2, t_nmask, t_base, t_len, t_rev};
    const OSeq V{v_pack2, v_nmask, v_base, v_len, v_rev};
    if (n <= 0 || p < 0 || p >= T.len) return 0;
    const int32_t p32 = (int32_t)p, v32 = (int32_t)v0;
    int32_t h = common_extension(T, p32, V, left ? v32 + n - 1 : v32, n, left);
    if (h < n) return h;
    // the flank is shorter than 2^31, so the self-comparison ends at a sequence edge long before the cap
    return n + common_extension(T, left ? p32 - n : p32 + n, T, p32, 0x7fffffff - n, left);
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
