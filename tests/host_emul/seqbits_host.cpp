// Host build of pav_b200/csrc/seqbits.cuh: the device functions that pack sequences, cut 32-base windows, run the homology
// scans and extract k-mers, compiled as plain C++ with the few CUDA intrinsics they use written out below. Test
// infrastructure (tests/test_device_logic_cpu.py): the same source text the kernels inline is run against the oracle and the
// golden vectors on a machine without a GPU. What it cannot cover: launch geometry, shuffles, shared memory, atomics.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#define PAV_DEV static inline
using std::min;
using std::max;

template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift)
{
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (shift & 31));
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline uint32_t __brev(uint32_t x)
{
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }

#ifndef HOM_BATCH_LOADS
#define HOM_BATCH_LOADS 1   // the host build exercises the batched-window forms too (EMU_BATCH=0 builds the plain ones)
#endif
#include "seqbits.cuh"

extern "C" {

// ASCII (n_words * 32 bytes, padded by the caller) -> planes, one pack_word32 per word like pack_kernel.
void emu_pack(const uint8_t *ascii, int64_t n_words, uint64_t *pack2, uint32_t *nmask)
{
    for (int64_t w = 0; w < n_words; w++) {
        uint32_t v[8];
        memcpy(v, ascii + w * 32, 32);
        pack_word32(v, pack2[w], nmask[w]);
    }
}

int emu_base(const uint64_t *pack2, const uint32_t *nmask, int64_t base, int64_t len, int rev, int64_t t)
{
    const OSeq s{pack2, nmask, base, len, rev, nullptr, nullptr, 0, 0, nullptr};
    return oseq_base(s, t);
}

void emu_window(const uint64_t *pack2, const uint32_t *nmask, int64_t base, int64_t len, int rev, int32_t t, uint64_t *bases, uint32_t *mask)
{
    const OSeq s{pack2, nmask, base, len, rev, nullptr, nullptr, 0, 0, nullptr};
    oseq_window(s, t, *bases, *mask);
}

// pavlib.call.left_homology / right_homology (pavgpu_homology's probe kernel): T = sequence 0, V = sequence 1 of the planes.
int emu_homology(const uint64_t *pack2, const uint32_t *nmask, int64_t t_base, int64_t t_len, int t_rev, int64_t p, int64_t v_base,
                 int64_t v_len, int v_rev, int64_t v0, int n, int left)
{
    return dev_homology_raw(pack2, nmask, t_base, t_len, t_rev, p, pack2, nmask, v_base, v_len, v_rev, v0, n, left);
}

// One indel through score_indel (version 1) or score_indel2 (version 2: convergent first trips), as homology_kernel (tile_words == 0) or with a staged copy of plane words
// [w0, w0 + tile_words) of each sequence, as homology_tiled_kernel / homology_nbr_kernel see it. out[10] = pos, end, qry_pos,
// qry_end, left_shift, hom_ref_l, hom_ref_r, hom_tig_l, hom_tig_r, seq_start.
void emu_score_indel(const uint64_t *r_pack2, const uint32_t *r_nmask, int64_t r_base, int64_t r_len, const uint64_t *q_pack2,
                     const uint32_t *q_nmask, int64_t q_base, int64_t q_len, int q_rev, int32_t svtype, int32_t n, int32_t pr, int32_t pq,
                     int32_t eqb, int64_t r_w0, int32_t r_tile_words, int64_t q_w0, int32_t q_tile_words, int32_t version,
                     const uint32_t *r_nsum, const uint32_t *q_nsum, int32_t *out)
{
    IndelScore o;
    if (r_tile_words <= 0 && q_tile_words <= 0) {
        const OSeq R{r_pack2, r_nmask, r_base, r_len, 0, nullptr, nullptr, 0, 0, r_nsum};     // N summaries optional (nullptr: masks always read)
        const OSeq Q{q_pack2, q_nmask, q_base, q_len, q_rev, nullptr, nullptr, 0, 0, q_nsum};
        if (version == 2) score_indel2<false>(R, Q, svtype, n, pr, pq, eqb, o);
        else score_indel<false>(R, Q, svtype, n, pr, pq, eqb, o);
    } else {
        // staged copies are poisoned outside the tile so that a window served from the wrong place cannot go unnoticed
        std::vector<uint64_t> tp_r(std::max(r_tile_words, 1)), tp_q(std::max(q_tile_words, 1));
        std::vector<uint32_t> tm_r(std::max(r_tile_words, 1)), tm_q(std::max(q_tile_words, 1));
        for (int k = 0; k < r_tile_words; k++) { tp_r[k] = r_pack2[r_w0 + k]; tm_r[k] = r_nmask[r_w0 + k]; }
        for (int k = 0; k < q_tile_words; k++) { tp_q[k] = q_pack2[q_w0 + k]; tm_q[k] = q_nmask[q_w0 + k]; }
        const OSeq R{r_pack2, r_nmask, r_base, r_len, 0, tp_r.data(), tm_r.data(), r_w0, std::max(r_tile_words - 1, 0), nullptr};
        const OSeq Q{q_pack2, q_nmask, q_base, q_len, q_rev, tp_q.data(), tm_q.data(), q_w0, std::max(q_tile_words - 1, 0), nullptr};
        if (version == 2) score_indel2<true>(R, Q, svtype, n, pr, pq, eqb, o);
        else score_indel<true>(R, Q, svtype, n, pr, pq, eqb, o);
    }
    out[0] = o.pos; out[1] = o.end; out[2] = o.qry_pos; out[3] = o.qry_end; out[4] = o.ls;
    out[5] = o.hom_rl; out[6] = o.hom_rr; out[7] = o.hom_tl; out[8] = o.hom_tr; out[9] = o.seq_start;
}

// N summary of a mask plane with nsum_kernel's layout: bit (w >> 3) & 31 of word w >> 8 set when mask word w is non-zero.
void emu_nsum(const uint32_t *nmask, int64_t n_words, uint32_t *nsum)
{
    for (int64_t w = 0; w < n_words; w++)
        if (nmask[w]) nsum[w >> 8] |= 1u << ((w >> 3) & 31);
}

// k-mers of a window the way ref_insert_kernel / tig_state_kernel read them: valid[i] = kmer_at(g0 + i), kmer[i], rc[i].
void emu_kmers(const uint64_t *pack2, const uint32_t *nmask, int64_t g0, int32_t n_pos, int k, uint64_t *kmer, uint64_t *rc, uint8_t *valid,
               const uint32_t *nsum)
{
    for (int32_t i = 0; i < n_pos; i++) {
        uint64_t x = 0;
        valid[i] = kmer_at(pack2, nmask, g0 + i, k, x, nsum) ? 1 : 0;
        kmer[i] = valid[i] ? x : 0;
        rc[i] = valid[i] ? kmer_revcomp(x, k) : 0;
    }
}

// Staging ranges of the opt-in kernels.
int64_t emu_nbr_first_word(int64_t c, int64_t plane_words) { return nbr_first_word(c, plane_words); }
void emu_tile_range(int64_t lo_g, int64_t hi_g, int64_t plane_words, int64_t *w0, int32_t *nw) { tile_range(lo_g, hi_g, plane_words, *w0, *nw); }
int emu_tile_words(void) { return HOM_TILE_WORDS; }
int emu_tile_margin(void) { return HOM_TILE_MARGIN; }

}  // extern "C"
