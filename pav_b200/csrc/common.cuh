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

// pavlib/call.py:542-592. T: sequence searched leftwards from p; the SV sequence is V[v0 : v0+n],
// read circularly from its end (sv[-((h+1) % n)], index -0 == 0).
__device__ __forceinline__ int dev_left_homology(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n)
{
    int h = 0;
    int vi = n - 1;
    while ((int64_t)h <= p) {
        int b = oseq_base(T, p - h);
        if (b == 4) break;
        if (oseq_base(V, v0 + vi) != b) break;
        ++h;
        if (--vi < 0) vi = n - 1;
    }
    return h;
}

// pavlib/call.py:595-647
__device__ __forceinline__ int dev_right_homology(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n)
{
    int h = 0;
    int vi = 0;
    int64_t limit = T.len - p;
    while ((int64_t)h < limit) {
        int b = oseq_base(T, p + h);
        if (b == 4) break;
        if (oseq_base(V, v0 + vi) != b) break;
        ++h;
        if (++vi == n) vi = 0;
    }
    return h;
}

static inline float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
