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
};

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
// `bases`, bit i of `mask` set when that base is not ACGT or lies outside [0, len)).
__device__ __forceinline__ void fwd_window(const OSeq &s, int64_t f, uint64_t &bases, uint32_t &mask)
{
    if (f <= -32 || f >= s.len) { bases = 0; mask = 0xffffffffu; return; }
    int lead = f < 0 ? (int)(-f) : 0;          // window positions before the sequence start
    int64_t g = s.base + f + lead;
    int64_t w = g >> 5;
    int sh = (int)(g & 31);
    uint64_t hi = __ldg(s.pack2 + w), lo = __ldg(s.pack2 + w + 1);
    uint64_t b = sh ? ((hi << (2 * sh)) | (lo >> (64 - 2 * sh))) : hi;
    uint64_t m64 = (uint64_t)__ldg(s.nmask + w) | ((uint64_t)__ldg(s.nmask + w + 1) << 32);
    uint32_t m = (uint32_t)(m64 >> sh);
    if (lead) { b >>= 2 * lead; m = (m << lead) | ((1u << lead) - 1u); }
    int64_t over = f + 32 - s.len;             // window positions past the sequence end
    if (over > 0) m |= ~0u << (32 - (int)over);
    bases = b; mask = m;
}

// Window of 32 bases starting at oriented position t (reverse-complement view when s.rev).
__device__ __forceinline__ void oseq_window(const OSeq &s, int64_t t, uint64_t &bases, uint32_t &mask)
{
    if (!s.rev) { fwd_window(s, t, bases, mask); return; }
    uint64_t b; uint32_t m;
    fwd_window(s, s.len - t - 32, b, m);
    bases = revcomp32(b);
    mask = __brev(m);
}

// Longest common prefix of A[a..] and B[b..] (stops at the first mismatch, non-ACGT base or sequence end
// on either side), capped at `limit`. 32 bases per step.
__device__ __forceinline__ int64_t lcp_forward(const OSeq &A, int64_t a, const OSeq &B, int64_t b, int64_t limit)
{
    int64_t h = 0;
    while (h < limit) {
        uint64_t wa, wb; uint32_t ma, mb;
        oseq_window(A, a + h, wa, ma);
        oseq_window(B, b + h, wb, mb);
        uint64_t x = wa ^ wb;
        uint64_t d = (x | (x >> 1)) & 0x5555555555555555ull;      // one bit per differing base
        int stop_d = d ? (__clzll((long long)d) >> 1) : 32;       // first base = most significant group
        uint32_t m = ma | mb;
        int stop_m = m ? (__ffs((int)m) - 1) : 32;
        int stop = min(stop_d, stop_m);
        if (stop < 32) { h += stop; return h < limit ? h : limit; }
        h += 32;
    }
    return limit;
}

// Longest common suffix of A[..a] and B[..b] (positions a, b inclusive, walking towards position 0).
__device__ __forceinline__ int64_t lcs_backward(const OSeq &A, int64_t a, const OSeq &B, int64_t b, int64_t limit)
{
    int64_t h = 0;
    while (h < limit) {
        uint64_t wa, wb; uint32_t ma, mb;
        oseq_window(A, a - h - 31, wa, ma);
        oseq_window(B, b - h - 31, wb, mb);
        uint64_t x = wa ^ wb;
        uint64_t d = (x | (x >> 1)) & 0x5555555555555555ull;
        int stop_d = d ? ((__ffsll((long long)d) - 1) >> 1) : 32;  // last base = least significant group
        uint32_t m = ma | mb;
        int stop_m = m ? __clz((int)m) : 32;
        int stop = min(stop_d, stop_m);
        if (stop < 32) { h += stop; return h < limit ? h : limit; }
        h += 32;
    }
    return limit;
}

// pavlib/call.py:542-592. T: sequence searched leftwards from p; the SV sequence is V[v0 : v0+n], read
// circularly from its end (sv[-((h+1) % n)], index -0 == 0). Word-parallel form: the first n steps are a
// common-suffix of T[..p] with V[v0..v0+n-1]; once a whole copy of V matched, step h compares T[p-h] with
// V[(n-1-h) mod n] = T[p-h+n], i.e. the scan continues as a common suffix of T[..p-n] with T[..p].
__device__ __forceinline__ int dev_left_homology(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n)
{
    if (p < 0 || n <= 0) return 0;
    int64_t h = lcs_backward(T, p, V, v0 + n - 1, n);
    if (h < n) return (int)h;
    return (int)(n + lcs_backward(T, p - n, T, p, (int64_t)1 << 40));
}

// pavlib/call.py:595-647, same idea forwards: T[p+h] vs V[h mod n] = T[p+h-n] after the first copy.
__device__ __forceinline__ int dev_right_homology(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n)
{
    if (p >= T.len || n <= 0 || p < 0) return 0;
    int64_t h = lcp_forward(T, p, V, v0, n);
    if (h < n) return (int)h;
    return (int)(n + lcp_forward(T, p + n, T, p, (int64_t)1 << 40));
}

static inline float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
