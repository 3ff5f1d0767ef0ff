// density.cu -- Path B: k-mer orientation states and smoothed state density for inversion windows.
//
// Reference semantics: scripts/density.py:423-571 (__main__) and :154-342 (get_smoothed_density),
// kanapy/util/kmer.py:61-69,118-133,186-221 (k-mer arithmetic / stream), pavlib/seq.py:305-325
// (reference k-mer Counter), scipy.stats.gaussian_kde (float64 Gaussian KDE). SURVEY appendix A.2.
//
// One *batch* of windows is processed per call (one window = one run of scripts/density.py):
//   D1+D2 kmer_window_kernel  windows of up to 53,248 reference k-mers: one CTA per window, the reference k-mer table in shared memory
//                          (16-bit positions into the staged planes instead of keys, canonical k-mers, one probe sequence per contig k-mer)
//   D1 ref_insert_kernel   larger windows: exact k-mers of the reference window -> per-window open-addressing table in HBM
//                          (62-bit keys + count of FURTHER copies: one atomic per distinct k-mer, two per repeat; more than 100
//                          copies or no k-mers => status 125)
//   D2 tig_state_kernel    contig k-mer + reverse complement, two probes -> STATE_MER per position
//   -- host: keep-mask (state count >= 20), N, dense row offsets --
//   D3 compact_kernel      order-preserving compaction of informative k-mers -> KMER / INDEX / STATE_MER
//   D4 runs_stats_kernel   one pass: maximal runs of equal STATE_MER in INDEX_DEN space (start, length, state) and, per
//                          state, n / mean / var(ddof=1) of INDEX_DEN (exact integer sums) -> bandwidth L_s and norm_s
//   D5 kde_table_kernel    T_s[d] = exp(-(d / L_s)^2 / 2), d in [0, N): data and evaluation points both live
//                          on the integer lattice, so N exps replace N * E exps; beside it the suffix sums S_s[d] = T_s[d] + ... + T_s[N-1]
//   D6 kde_eval_kernel     K_s(j) = norm_s * sum over runs [a,b] of state s of sum_{i=a..b} T_s[|i - j|]; the
//                          inner sum is one or two range sums: S_s[lo] - S_s[hi + 1] (runs shorter than 8 are summed from T_s),
//                          so a point costs O(#runs) instead of O(N); 8 lanes share one point
//   D7 gap_classify_kernel per gap: state change / argmax change / |delta| > 0.005 => list of points that
//                          need a full evaluation
//   D6' kde_eval_kernel    on that list
//   D8/D9 finish_rows_kernel  one pass that writes every row: sampled values (kept in dense arrays of their own), numpy.interp for the
//                          gaps not evaluated in full, spikes (> 1 -> reciprocal), STATE = argmax (first max wins)
//   D10 state_rle_kernel   run lengths of STATE (the rl_encoder tuples)
//
// The k-mer of position g is read straight out of the 2-bit plane with a funnel shift (no rolling
// state, so every position is independent); validity is k zero bits of the N-mask plane, which is
// exactly the stream's "load == k" condition.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>


#include "common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint64_t EMPTY_KEY = ~0ull;
constexpr int TILE = 1024;       // positions per block in D2 / D3
constexpr int EVAL_LANES = 8;    // lanes that share one evaluation point in D6
constexpr int EVAL_THREADS = 256;
constexpr int EVAL_GROUP = EVAL_THREADS / EVAL_LANES;  // evaluation points per block
constexpr int TREE_THREADS = 1024;

struct WinPlan {  // per window, device-visible
    // inputs
    int64_t ref_g0, tig_g0;       // global base offsets of the windows in the planes
    int32_t ref_len, tig_len;     // window lengths in bases
    int32_t rev, srs;
    // tables
    int64_t tab_off;              // slot offset of this window's hash table
    int32_t tab_log2;             // log2(capacity)
    int32_t mode;                 // k-mer part: 0 = global tables (ref_insert + tig_state), 1 = kmer_window_kernel, 2 = already done
    int64_t pos_off;              // offset of this window in the per-position scratch (tig positions)
    int64_t tile_off;             // offset of this window's tiles in the per-tile count array
    // after the first host sync
    int32_t status;               // 0 / 125
    int32_t keep_mask;            // bit s set when state s is kept
    int32_t n_rows;               // N
    int32_t smoothed;
    int64_t row_off;              // dense row offset
    int32_t n_samp;               // sampled lattice points
    int32_t npad;                 // leaves of the sum tree (power of two >= N)
    int64_t tree_off;             // offset (doubles) of this window's tree inside each per-state tree slab
};

struct WinCounts {  // written by D1 / D2
    unsigned long long ref_valid;
    unsigned int ref_max;
    unsigned int cnt[3];
    unsigned int overflow;        // kmer_window_kernel: a k-mer count went above what its 6 bits decide (the window is redone with the global tables)
    unsigned int pad;
};

struct KdeParams {  // per window, written by D4
    double L[3];
    double norm[3];
    int32_t n[3];
    int32_t pad;
};

// D1 ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ref_insert_kernel(const WinPlan *__restrict__ plan, int32_t win_base, SeqPlanes ref, int k, uint64_t *__restrict__ keys,
                  uint32_t *__restrict__ counts, WinCounts *__restrict__ wc)
{
    int32_t w = win_base + blockIdx.y;
    const WinPlan P = plan[w];
    if (P.mode != 0) return;
    int32_t n_pos = P.ref_len - k + 1;
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    uint64_t kmer = 0;
    if (i < n_pos) valid = kmer_at(ref.pack2, ref.nmask, P.ref_g0 + i, k, kmer, ref.nsum);
    unsigned int my_cnt = 0;
    if (valid) {
        if (P.rev) kmer = kmer_revcomp(kmer, k);  // density.py:538-539: the reference SET is reverse-complemented
        uint64_t *tk = keys + P.tab_off;
        uint32_t *tc = counts + P.tab_off;
        uint64_t mask = (1ull << P.tab_log2) - 1;
        uint64_t slot = hash_kmer(kmer, P.tab_log2);
        while (true) {
            unsigned long long old = atomicCAS((unsigned long long *)(tk + slot), (unsigned long long)EMPTY_KEY, (unsigned long long)kmer);
            if (old == EMPTY_KEY) { my_cnt = 1u; break; }            // first copy of this k-mer: the slot's count stays 0 = "no further copies"
            if (old == kmer) { my_cnt = atomicAdd(tc + slot, 1u) + 2u; break; }   // a further copy: only these pay a second atomic (rare outside repeats)
            slot = (slot + 1) & mask;
        }
    }
    // block aggregates -> per-window counters
    unsigned n_valid = __syncthreads_count(valid);
    unsigned mx = my_cnt;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
    __shared__ unsigned s_mx[8];
    if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned m2 = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) m2 = max(m2, s_mx[q]);
        if (n_valid) atomicAdd(&wc[w].ref_valid, (unsigned long long)n_valid);
        if (m2) atomicMax(&wc[w].ref_max, m2);
    }
}

// D2 ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool table_has(const uint64_t *__restrict__ tk, int log2cap, uint64_t key)
{
    uint64_t mask = (1ull << log2cap) - 1;
    uint64_t slot = hash_kmer(key, log2cap);
    while (true) {
        uint64_t v = __ldg(tk + slot);
        if (v == key) return true;
        if (v == EMPTY_KEY) return false;
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(TILE)
tig_state_kernel(const WinPlan *__restrict__ plan, int32_t win_base, SeqPlanes tig, int k, const uint64_t *__restrict__ keys,
                 int8_t *__restrict__ st_pos, uint32_t *__restrict__ tile_cnt, WinCounts *__restrict__ wc)
{
    int32_t w = win_base + blockIdx.y;
    const WinPlan P = plan[w];
    if (P.mode != 0) return;
    int32_t n_pos = P.tig_len - k + 1;
    int32_t n_tiles = (max(n_pos, 0) + TILE - 1) / TILE;
    if ((int)blockIdx.x >= n_tiles) return;
    int32_t i = blockIdx.x * TILE + threadIdx.x;
    int st = -1;
    if (i < n_pos) {
        uint64_t kmer;
        if (kmer_at(tig.pack2, tig.nmask, P.tig_g0 + i, k, kmer, tig.nsum)) {
            const uint64_t *tk = keys + P.tab_off;
            bool f = table_has(tk, P.tab_log2, kmer);
            bool r = table_has(tk, P.tab_log2, kmer_revcomp(kmer, k));
            st = f ? (r ? 1 : 0) : (r ? 2 : -1);  // KMER_ORIENTATION_STATE, density.py:38-43
        }
        st_pos[P.pos_off + i] = (int8_t)st;
    }
    unsigned c0 = __syncthreads_count(st == 0);
    unsigned c1 = __syncthreads_count(st == 1);
    unsigned c2 = __syncthreads_count(st == 2);
    if (threadIdx.x == 0) {
        uint32_t *tc = tile_cnt + (P.tile_off + blockIdx.x) * 3;
        tc[0] = c0; tc[1] = c1; tc[2] = c2;
        if (c0) atomicAdd(&wc[w].cnt[0], c0);
        if (c1) atomicAdd(&wc[w].cnt[1], c1);
        if (c2) atomicAdd(&wc[w].cnt[2], c2);
    }
}

// D1 + D2 on chip --------------------------------------------------------------------------------------
// One 1,024-thread CTA per window; the window's reference k-mer table lives in the shared memory of its SM and never touches L2
// or HBM. What makes it fit: the table does not hold keys. The CTA first stages the window's slice of the packed reference planes
// (12 bytes per 32 bases: 20 KB for 50 kbp); a slot then holds the 16-bit POSITION of the first copy of its k-mer, and a probe
// re-derives that k-mer from the staged planes to compare. k-mers are stored canonically (the smaller of the k-mer and its reverse
// complement), with one byte per slot beside the position: bit 0 / bit 1 = "seen in this / the other orientation", bits 2-7 = copies
// seen in either. A contig k-mer therefore costs ONE probe sequence, not one per orientation, and contig k-mers that are in the
// reference in either orientation -- nearly all of them -- end on a hit. 65,536 slots x 3 bytes = 192 KB + 20 KB of planes.
//   * windows with at most KS_MAX_REF reference k-mers take this kernel (WinPlan.mode 1), larger ones the global tables (mode 0);
//   * MAX_REF_KMER_COUNT (density.py:47, :510-527) asks whether some ORIENTED k-mer has more than that many copies. The 6-bit
//     count is of both orientations together, so "no count above min(limit, 60)" proves the answer is no; a window where some count
//     goes above that is flagged (WinCounts.overflow) and redone with the global tables, which count exactly.
// Measured on B200 (r02, 296 windows x 50 kbp): see DESIGN.md section 6.1; the r02 attempt with the table spread over the distributed shared
// memory of an 8-CTA cluster (8-byte keys, remote atomics) took 1.42 ms against 0.81 ms for the global tables and was removed.
constexpr int KS_THREADS = 1024;
constexpr int KS_MAX_LOG2 = 16;
constexpr int KS_MAX_REF = 53248;                                   // reference k-mers per window: load <= 0.8125
constexpr int KS_SEQ_WORDS = 1672;                                  // plane words staged: (31 + KS_MAX_REF + 30 + 31) / 32 + 1, rounded up
constexpr int KS_CNT_LIMIT = 60;
constexpr int KS_MAX_TILES = 256;                                   // contig tiles (of 1,024 positions) whose counters fit beside the table
constexpr unsigned KS_EMPTY = 0xFFFFu;
constexpr size_t KS_SMEM = ((size_t)3 << KS_MAX_LOG2) + (size_t)KS_SEQ_WORDS * 12;
static_assert((31 + KS_MAX_REF + 30 + 31) / 32 + 1 <= KS_SEQ_WORDS, "staging area too small");
static_assert(KS_SMEM + KS_MAX_TILES * 12 + 1024 <= 227 * 1024, "does not fit the shared memory of an SM");

// k-mer at base `g` of the staged slice (plain shared-memory loads; kmer_at is the same arithmetic on the planes in HBM)
__device__ __forceinline__ uint64_t kmer_staged(const uint64_t *sp, int32_t g, int k)
{
    const int32_t w = g >> 5;
    const int s = g & 31;
    const uint64_t hi = sp[w], lo = sp[w + 1];
    const uint64_t x = s ? ((hi << (2 * s)) | (lo >> (64 - 2 * s))) : hi;
    return x >> (64 - 2 * k);
}

__device__ __forceinline__ bool kmer_staged_valid(const uint32_t *sm, int32_t g, int k)
{
    const int32_t w = g >> 5;
    const uint64_t m = ((uint64_t)sm[w] | ((uint64_t)sm[w + 1] << 32)) >> (g & 31);
    return (m & ((1ull << k) - 1)) == 0;
}

__global__ void __launch_bounds__(KS_THREADS, 1)
kmer_window_kernel(const WinPlan *__restrict__ plan, SeqPlanes ref, SeqPlanes tig, int k, unsigned cnt_limit, int8_t *__restrict__ st_pos,
                   uint32_t *__restrict__ tile_cnt, WinCounts *__restrict__ wc)
{
    extern __shared__ __align__(16) unsigned char ks_smem[];
    unsigned short *pos16 = reinterpret_cast<unsigned short *>(ks_smem);
    uint32_t *meta = reinterpret_cast<uint32_t *>(ks_smem + ((size_t)2 << KS_MAX_LOG2));          // one byte per slot
    uint64_t *sp = reinterpret_cast<uint64_t *>(ks_smem + ((size_t)3 << KS_MAX_LOG2));
    uint32_t *sm = reinterpret_cast<uint32_t *>(ks_smem + ((size_t)3 << KS_MAX_LOG2) + (size_t)KS_SEQ_WORDS * 8);
    __shared__ unsigned s_tile[KS_MAX_TILES * 3];
    const int32_t w = blockIdx.x;
    const WinPlan P = plan[w];
    if (P.mode != 1) return;
    const int lg = max(min(P.tab_log2, KS_MAX_LOG2), 8);
    const unsigned slots = 1u << lg, mask = slots - 1;
    {
        const uint4 ff = make_uint4(~0u, ~0u, ~0u, ~0u), zz = make_uint4(0u, 0u, 0u, 0u);
        for (unsigned i = threadIdx.x; i < slots / 8; i += KS_THREADS) reinterpret_cast<uint4 *>(pos16)[i] = ff;
        for (unsigned i = threadIdx.x; i < slots / 16; i += KS_THREADS) reinterpret_cast<uint4 *>(meta)[i] = zz;
    }
    const int64_t w0 = P.ref_g0 >> 5;
    const int32_t off = (int32_t)(P.ref_g0 & 31);
    const int32_t n_ref = P.ref_len - k + 1;
    const int32_t nw = (off + P.ref_len + 31) / 32 + 1;
    for (int32_t i = threadIdx.x; i < nw; i += KS_THREADS) { sp[i] = __ldg(ref.pack2 + w0 + i); sm[i] = __ldg(ref.nmask + w0 + i); }
    __syncthreads();
    // Probing: double hashing (the step is an odd number from other bits of the hash, so a sequence visits every slot). Shared memory
    // has no lines to stay inside, and at the load this table runs at (up to 0.81) linear probing's clusters cost 9 probes per miss
    // and tails of 40 and more in a warp; double hashing 4 and ~15.
    // Lanes STREAM: a lane whose k-mer is settled takes its next position in the same trip of the loop instead of waiting for the
    // slowest probe sequence of the warp (ncu, r02, with every lane waiting: 10.8 of 32 lanes active per instruction, issue-bound).
    // ---- reference k-mers -> table
    unsigned n_valid = 0, my_max = 0, over = 0;
    {
        int32_t next = threadIdx.x, i = 0;
        bool act = false;
        uint64_t raw = 0, rcr = 0;
        unsigned slot = 0, step = 0, o = 0;
        while (true) {
            if (!act && next < n_ref) {
                i = next; next += KS_THREADS;
                if (kmer_staged_valid(sm, off + i, k)) {
                    n_valid++;
                    raw = kmer_staged(sp, off + i, k); rcr = kmer_revcomp(raw, k);
                    const uint64_t kmer = P.rev ? rcr : raw;          // density.py:538-539: the reference SET is reverse-complemented
                    const uint64_t other = P.rev ? raw : rcr;
                    o = kmer > other ? 1u : 0u;                       // orientation of `kmer` relative to the canonical form
                    const uint64_t h = (o ? other : kmer) * 0x9E3779B97F4A7C15ull;
                    slot = (unsigned)(h >> (64 - lg)); step = ((unsigned)(h >> 7) | 1u) & mask;
                    act = true;
                }
            }
            if (!__any_sync(FULL, act || next < n_ref)) break;
            if (act) {
                unsigned cur = *reinterpret_cast<volatile unsigned short *>(pos16 + slot);
                if (cur == KS_EMPTY) {
                    cur = atomicCAS(pos16 + slot, (unsigned short)KS_EMPTY, (unsigned short)i);
                    if (cur == KS_EMPTY) cur = (unsigned)i;
                }
                bool same = cur == (unsigned)i;
                if (!same) { const uint64_t k2 = kmer_staged(sp, off + (int32_t)cur, k); same = (k2 == raw) || (k2 == rcr); }
                if (same) {
                    const unsigned sh = 8u * (slot & 3u);
                    const unsigned old = atomicAdd(meta + (slot >> 2), 4u << sh) >> sh;
                    if (!((old >> o) & 1u)) atomicOr(meta + (slot >> 2), (1u << o) << sh);
                    const unsigned c = ((old >> 2) & 63u) + 1u;
                    my_max = max(my_max, c);
                    if (c > cnt_limit) over = 1u;             // flagged before the 6-bit field can carry into its neighbour
                    act = false;
                } else slot = (slot + step) & mask;
            }
        }
    }
    if (__syncthreads_or((int)over)) {                        // redone with the global tables; nothing of this window is used
        if (threadIdx.x == 0) atomicOr(&wc[w].overflow, 1u);
        return;
    }
    {
        const unsigned nv = __reduce_add_sync(FULL, n_valid), mx = __reduce_max_sync(FULL, my_max);
        if ((threadIdx.x & 31) == 0) {
            if (nv) atomicAdd(&wc[w].ref_valid, (unsigned long long)nv);
            if (mx) atomicMax(&wc[w].ref_max, mx);            // an upper bound of the oriented maximum, <= the limit here
        }
    }
    // ---- contig k-mers: state per position, counts per 1,024-position tile (tig_state_kernel's outputs)
    const int32_t n_pos = P.tig_len - k + 1;
    const int32_t n_tiles = (max(n_pos, 0) + TILE - 1) / TILE;
    for (int32_t t = threadIdx.x; t < n_tiles * 3; t += KS_THREADS) s_tile[t] = 0u;
    __syncthreads();
    {
        int32_t next = threadIdx.x, i = 0;
        bool act = false;
        uint64_t km = 0, rc = 0;
        unsigned slot = 0, step = 0;
        while (true) {
            if (!act && next < n_pos) {
                i = next; next += KS_THREADS;
                if (kmer_at(tig.pack2, tig.nmask, P.tig_g0 + i, k, km, tig.nsum)) {
                    rc = kmer_revcomp(km, k);
                    const uint64_t h = (km > rc ? rc : km) * 0x9E3779B97F4A7C15ull;
                    slot = (unsigned)(h >> (64 - lg)); step = ((unsigned)(h >> 7) | 1u) & mask;
                    act = true;
                } else st_pos[P.pos_off + i] = (int8_t)-1;
            }
            if (!__any_sync(FULL, act || next < n_pos)) break;
            if (act) {
                const unsigned cur = pos16[slot];
                int st = -2;                                   // -2: keep probing
                if (cur == KS_EMPTY) st = -1;                  // in neither orientation
                else {
                    const uint64_t k2 = kmer_staged(sp, off + (int32_t)cur, k);
                    if (k2 == km || k2 == rc) {
                        const unsigned bits = (meta[slot >> 2] >> (8u * (slot & 3u))) & 3u;
                        const unsigned o = km > rc ? 1u : 0u;
                        const bool f = (bits >> o) & 1u;
                        const bool r = (km == rc) ? f : ((bits >> (o ^ 1u)) & 1u);
                        st = f ? (r ? 1 : 0) : (r ? 2 : -1);  // KMER_ORIENTATION_STATE, density.py:38-43
                    }
                }
                if (st == -2) slot = (slot + step) & mask;
                else {
                    st_pos[P.pos_off + i] = (int8_t)st;
                    if (st >= 0) atomicAdd(&s_tile[(i / TILE) * 3 + st], 1u);
                    act = false;
                }
            }
        }
    }
    __syncthreads();
    unsigned t0 = 0, t1 = 0, t2 = 0;
    for (int32_t t = threadIdx.x; t < n_tiles; t += KS_THREADS) {
        uint32_t *tc = tile_cnt + (P.tile_off + t) * 3;
        const unsigned c0 = s_tile[t * 3], c1 = s_tile[t * 3 + 1], c2 = s_tile[t * 3 + 2];
        tc[0] = c0; tc[1] = c1; tc[2] = c2;
        t0 += c0; t1 += c1; t2 += c2;
    }
    t0 = __reduce_add_sync(FULL, t0); t1 = __reduce_add_sync(FULL, t1); t2 = __reduce_add_sync(FULL, t2);
    if ((threadIdx.x & 31) == 0 && (t0 | t1 | t2)) {
        if (t0) atomicAdd(&wc[w].cnt[0], t0);
        if (t1) atomicAdd(&wc[w].cnt[1], t1);
        if (t2) atomicAdd(&wc[w].cnt[2], t2);
    }
}

// D3 ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE)
compact_kernel(const WinPlan *__restrict__ plan, int32_t win_base, SeqPlanes tig, int k, const int8_t *__restrict__ st_pos,
               const uint32_t *__restrict__ tile_cnt, uint64_t *__restrict__ o_kmer, int32_t *__restrict__ o_index,
               int8_t *__restrict__ o_state_mer)
{
    int32_t w = win_base + blockIdx.y;
    const WinPlan P = plan[w];
    if (P.status != 0 || P.n_rows == 0) return;
    int32_t n_pos = P.tig_len - k + 1;
    int32_t n_tiles = (max(n_pos, 0) + TILE - 1) / TILE;
    if ((int)blockIdx.x >= n_tiles) return;
    __shared__ unsigned s_base;
    __shared__ unsigned s_warp[TILE / 32];
    // this position's state and k-mer first: their loads do not depend on the prefix below and overlap with it (the CTA is one chain of
    // dependent loads -- plan, tile counts, state, planes -- and two of them are resident per SM)
    const int32_t i = blockIdx.x * TILE + threadIdx.x;
    const int st = (i < n_pos) ? (int)st_pos[P.pos_off + i] : -1;
    uint64_t kmer = 0;
    if (i < n_pos) kmer_at(tig.pack2, tig.nmask, P.tig_g0 + i, k, kmer, tig.nsum);
    // rows written by earlier tiles of this window
    unsigned part = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += blockDim.x) {
        const uint32_t *tc = tile_cnt + (P.tile_off + t) * 3;
        part += ((P.keep_mask & 1) ? tc[0] : 0) + ((P.keep_mask & 2) ? tc[1] : 0) + ((P.keep_mask & 4) ? tc[2] : 0);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(FULL, part, d);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && part) atomicAdd(&s_base, part);
    __syncthreads();
    bool keep = st >= 0 && ((P.keep_mask >> st) & 1);
    unsigned bal = __ballot_sync(FULL, keep);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {   // exclusive scan of the 32 warp counts by one warp (was: every thread summing up to 31 counts)
        unsigned c = s_warp[lane], inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += u; }
        s_warp[lane] = inc - c;
    }
    __syncthreads();
    const unsigned before = s_warp[wid];
    if (keep) {
        int64_t row = P.row_off + s_base + before + __popc(bal & ((1u << lane) - 1));
        o_kmer[row] = kmer;
        o_index[row] = i;
        o_state_mer[row] = (int8_t)st;
    }
}

// D4 + D5a ---------------------------------------------------------------------------------------
// One block per window, one pass over STATE_MER (INDEX_DEN = row number):
//   * maximal runs of equal state, in order, at run_start/run_len/run_state[row_off + r] (ballot compaction);
//   * per state n, sum(i), sum(i^2) as exact integers -> var(ddof=1) = (n*S2 - S1^2) / (n*(n-1)) with a 128-bit
//     numerator, bandwidth L_s = sqrt(var) * N^(-1/5) * smooth and norm_s = (2 pi)^(-1/2) / L_s (scipy gaussian_kde:
//     covariance = np.cov(bias=False) * factor^2, cho_cov = cholesky(cov)).
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long *s_buf)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += s_buf[q];
    return t;
}

__global__ void __launch_bounds__(1024)
runs_stats_kernel(const WinPlan *__restrict__ plan, const int8_t *__restrict__ state_mer, double smooth, int32_t *__restrict__ run_start,
                  int32_t *__restrict__ run_len, int8_t *__restrict__ run_state, int32_t *__restrict__ n_runs, KdeParams *__restrict__ kp)
{
    int32_t w = blockIdx.x;
    const WinPlan P = plan[w];
    if (!P.smoothed) { if (threadIdx.x == 0) n_runs[w] = 0; return; }
    const int8_t *sm = state_mer + P.row_off;
    const int32_t N = P.n_rows;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ int s_warp[32];
    __shared__ int s_total;
    __shared__ unsigned long long s_buf[32];
    // every thread owns a contiguous slice of the rows: its run heads are counted, ranked by a block scan and written; the sums
    // for the bandwidths come from the same pass (r01 walked the column 1,024 rows at a time with three barriers per step)
    const int32_t per = (N + (int32_t)blockDim.x - 1) / (int32_t)blockDim.x;
    const int32_t lo = min((int32_t)threadIdx.x * per, N), hi = min(lo + per, N);
    unsigned long long cn[3] = {0, 0, 0}, c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
    int cnt = 0;
    {
        int prev = lo > 0 ? (int)sm[lo - 1] : -2;
        for (int32_t i = lo; i < hi; i++) {
            const int st = (int)sm[i];
            cnt += (st != prev) ? 1 : 0;
            prev = st;
#pragma unroll
            for (int s = 0; s < 3; s++)
                if (st == s) { cn[s] += 1; c1[s] += (unsigned long long)i; c2[s] += (unsigned long long)i * (unsigned long long)i; }
        }
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int before = 0;
    for (int q = 0; q < wid; q++) before += s_warp[q];
    if (threadIdx.x == blockDim.x - 1) s_total = before + inc;
    {
        int32_t r = before + inc - cnt;
        int prev = lo > 0 ? (int)sm[lo - 1] : -2;
        for (int32_t i = lo; i < hi; i++) {
            const int st = (int)sm[i];
            if (st != prev) { run_start[P.row_off + r] = i; run_state[P.row_off + r] = (int8_t)st; r++; }
            prev = st;
        }
    }
    __syncthreads();
    int32_t nr = s_total;
    for (int32_t r = threadIdx.x; r < nr; r += blockDim.x) {
        int32_t a = run_start[P.row_off + r];
        int32_t e = (r + 1 < nr) ? run_start[P.row_off + r + 1] : N;
        run_len[P.row_off + r] = e - a;
    }
    if (threadIdx.x == 0) n_runs[w] = nr;
    // ---- per-state bandwidths
    double bw = pow((double)N, -1.0 / 5.0) * smooth;  // density.py:198
    for (int s = 0; s < 3; s++) {
        unsigned long long n = block_sum_u64(cn[s], s_buf);
        unsigned long long S1 = block_sum_u64(c1[s], s_buf);
        unsigned long long S2 = block_sum_u64(c2[s], s_buf);
        if (threadIdx.x == 0) {
            double L = 1.0, norm = 0.0;
            if (n > 0) {
                unsigned __int128 num = (unsigned __int128)n * S2 - (unsigned __int128)S1 * S1;   // exact: n*sum(i^2) - sum(i)^2 >= 0
                double var = (double)num / ((double)n * (double)(n - 1));                        // np.cov(bias=False)
                L = sqrt(var) * bw;                  // cho_cov = cholesky(cov) * factor
                norm = pow(2.0 * M_PI, -0.5) / L;    // (2 pi)^(-d/2) / cho_cov[0,0]
            }
            kp[w].L[s] = L; kp[w].norm[s] = norm; kp[w].n[s] = (int32_t)n;
        }
    }
}

// D5b ---------------------------------------------------------------------------------------------
// One block per (window, state): the Gaussian table T[d] = exp(-(d / L)^2 / 2), d in [0, N), and its suffix sums
// S[d] = T[d] + T[d + 1] + ... + T[N - 1] (S[N] = 0), accumulated from the far tail towards d = 0, i.e. from the smallest terms to
// the largest. Layout inside the window's slot of the per-state slab: T at [0, N), S at [N, 2 N + 1).
// (Round 1 kept a binary sum tree here -- 2 log2 N reads per range sum, 1.06 GB written for 296 windows; a range sum over distances
// lo..hi is S[lo] - S[hi + 1]: two reads. The subtraction loses about log10(S[lo] / result) digits, which for runs of 8 or more
// k-mers is one to two digits of sixteen; shorter runs are summed term by term from T.)
constexpr int SCAN_ITEMS = 4;

__global__ void __launch_bounds__(TREE_THREADS)
kde_table_kernel(const WinPlan *__restrict__ plan, const KdeParams *__restrict__ kp, double *__restrict__ tab0, double *__restrict__ tab1,
                 double *__restrict__ tab2)
{
    const int32_t w = blockIdx.x / 3, s = blockIdx.x % 3;
    const WinPlan P = plan[w];
    if (!P.smoothed) return;
    if (kp[w].n[s] == 0) return;   // state absent: never read
    double *T = (s == 0 ? tab0 : (s == 1 ? tab1 : tab2)) + P.tree_off;
    const int32_t N = P.n_rows;
    double *S = T + N;
    // exp(-(d / L)^2 / 2) as exp(-(d * d) * c), c = 1 / (2 L^2): d * d is an exact integer, so the argument carries the rounding of c and of
    // one product -- as many roundings as (d / L) squared, without a float64 division per element (this kernel is bound by float64
    // arithmetic: ~38 M exp for 296 windows)
    const double L = kp[w].L[s];
    const double c = 1.0 / (2.0 * (L * L));
    for (int32_t d = threadIdx.x; d < N; d += blockDim.x) T[d] = exp(-((double)((int64_t)d * d) * c));
    __shared__ double s_warp[TREE_THREADS / 32];
    __shared__ double s_carry;
    if (threadIdx.x == 0) { s_carry = 0.0; S[N] = 0.0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int CH = TREE_THREADS * SCAN_ITEMS;
    // chunks from the tail; inside a chunk thread t owns elements [base + t * ITEMS, +ITEMS), later threads hold the farther distances
    for (int32_t base = ((N - 1) / CH) * CH; base >= 0; base -= CH) {
        const int32_t d0 = base + threadIdx.x * SCAN_ITEMS;
        double v[SCAN_ITEMS];
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) v[i] = (d0 + i < N) ? T[d0 + i] : 0.0;
        // suffix sums inside the thread, far element first
#pragma unroll
        for (int i = SCAN_ITEMS - 2; i >= 0; i--) v[i] += v[i + 1];
        // exclusive suffix over threads: sum of the totals of all threads with a larger index
        double tot = v[0], inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_down_sync(FULL, inc, o);
            if (lane + o < 32) inc += u;
        }
        if (lane == 0) s_warp[wid] = inc;           // warp total
        // what lies behind me inside the warp = the inclusive suffix of the next lane. (Not inc - tot: with a narrow bandwidth the
        // table falls by many orders of magnitude per element and the difference of two nearly equal sums is all rounding error.)
        const double nxt = __shfl_down_sync(FULL, inc, 1);
        const double behind = lane < 31 ? nxt : 0.0;
        __syncthreads();
        double after = 0.0;                          // totals of the warps behind mine, farthest first
        for (int q = TREE_THREADS / 32 - 1; q > wid; q--) after += s_warp[q];
        const double excl = behind + (after + s_carry);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++)
            if (d0 + i < N) S[d0 + i] = v[i] + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = v[0] + excl;  // S[base]
        __syncthreads();
    }
}

// Sum of T over distances lo..hi (inclusive, 0 <= lo <= hi < N); all terms are >= 0.
__device__ __forceinline__ double table_range_sum(const double *__restrict__ T, const double *__restrict__ S, int32_t lo, int32_t hi)
{
    if (hi - lo < 8) {
        double acc = 0.0;
        for (int32_t d = hi; d >= lo; d--) acc += __ldg(T + d);
        return acc;
    }
    return __ldg(S + lo) - __ldg(S + hi + 1);
}

// D6 ---------------------------------------------------------------------------------------------
// grp_off[w] = first block of window w (prefix over ceil(n_eval_w / EVAL_GROUP)). mode 0: sampled lattice
// (e-th sample = min(e * srs, N-1)); mode 1: explicit list fill_list[row_off + e]. EVAL_LANES lanes share a point.
__global__ void __launch_bounds__(EVAL_THREADS)
kde_eval_kernel(const WinPlan *__restrict__ plan, int32_t n_win, const int64_t *__restrict__ grp_off, const int32_t *__restrict__ n_eval_w,
                int mode, const int32_t *__restrict__ fill_list, const KdeParams *__restrict__ kp, const int32_t *__restrict__ run_start,
                const int32_t *__restrict__ run_len, const int8_t *__restrict__ run_state, const int32_t *__restrict__ n_runs,
                const double *__restrict__ tree0, const double *__restrict__ tree1, const double *__restrict__ tree2,
                double *__restrict__ k0, double *__restrict__ k1, double *__restrict__ k2)
{
    int64_t blk = blockIdx.x;
    int32_t lo = 0, hi = n_win;
    while (hi - lo > 1) {
        int32_t mid = (lo + hi) >> 1;
        if (grp_off[mid] <= blk) lo = mid; else hi = mid;
    }
    const int32_t w = lo;
    const WinPlan P = plan[w];
    const int32_t N = P.n_rows;
    const int sub = threadIdx.x % EVAL_LANES;
    int32_t e = (int32_t)(blk - grp_off[w]) * EVAL_GROUP + threadIdx.x / EVAL_LANES;
    int32_t j = -1;
    if (e < n_eval_w[w]) {
        if (mode == 0) { int64_t jj = (int64_t)e * P.srs; j = jj > N - 1 ? N - 1 : (int32_t)jj; }
        else j = fill_list[P.row_off + e];
    }
    const double *t0 = tree0 + P.tree_off, *t1 = tree1 + P.tree_off, *t2 = tree2 + P.tree_off;   // per state: T at [0, N), S at [N, 2 N + 1)
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (j >= 0) {
        const int32_t nr = n_runs[w];
        const int32_t *rs = run_start + P.row_off, *rl = run_len + P.row_off;
        const int8_t *rst = run_state + P.row_off;
        for (int32_t r = sub; r < nr; r += EVAL_LANES) {
            int32_t a = __ldg(rs + r), b = a + __ldg(rl + r) - 1;
            int s = __ldg(rst + r);
            const double *t = s == 0 ? t0 : (s == 1 ? t1 : t2);
            const double *sfx = t + N;
            double v;
            if (j < a) v = table_range_sum(t, sfx, a - j, b - j);
            else if (j > b) v = table_range_sum(t, sfx, j - b, j - a);
            else {
                v = table_range_sum(t, sfx, 0, j - a);
                if (b > j) v += table_range_sum(t, sfx, 1, b - j);
            }
            a0 += (s == 0) ? v : 0.0;
            a1 += (s == 1) ? v : 0.0;
            a2 += (s == 2) ? v : 0.0;
        }
    }
#pragma unroll
    for (int d = 1; d < EVAL_LANES; d <<= 1) {
        a0 += __shfl_xor_sync(FULL, a0, d);
        a1 += __shfl_xor_sync(FULL, a1, d);
        a2 += __shfl_xor_sync(FULL, a2, d);
    }
    if (j >= 0 && sub == 0) {   // mode 0: the e-th sample, kept densely (row_off + e); mode 1: the row itself
        const int64_t o = P.row_off + (mode == 0 ? e : j);
        k0[o] = a0 * kp[w].norm[0];
        k1[o] = a1 * kp[w].norm[1];
        k2[o] = a2 * kp[w].norm[2];
    }
}

__device__ __forceinline__ int argmax3(double a, double b, double c)
{
    int m = 0;
    double v = a;
    if (b > v) { m = 1; v = b; }
    if (c > v) m = 2;
    return m;
}

// D7 ---------------------------------------------------------------------------------------------
// One block per window. Gap g spans samples a = g * srs and b = min(a + srs, N - 1).
__global__ void __launch_bounds__(256)
gap_classify_kernel(const WinPlan *__restrict__ plan, double delta, const int8_t *__restrict__ state_mer, const double *__restrict__ k0,
                    const double *__restrict__ k1, const double *__restrict__ k2, uint8_t *__restrict__ gap_full,
                    int32_t *__restrict__ fill_list, int32_t *__restrict__ n_fill)
{
    int32_t w = blockIdx.x;
    const WinPlan P = plan[w];
    if (!P.smoothed) { if (threadIdx.x == 0) n_fill[w] = 0; return; }
    int32_t N = P.n_rows, srs = P.srs;
    int32_t n_gap = P.n_samp - 1;
    const int8_t *sm = state_mer + P.row_off;
    const double *a0 = k0 + P.row_off, *a1 = k1 + P.row_off, *a2 = k2 + P.row_off;   // the SAMPLED values: entry g = row g * srs, the last = row N - 1
    __shared__ int s_scan[256];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int32_t g0 = 0; g0 < n_gap; g0 += blockDim.x) {
        int32_t g = g0 + threadIdx.x;
        int cnt = 0;
        int32_t a = 0, b = 0;
        if (g < n_gap) {
            a = g * srs;
            b = min(a + srs, N - 1);
            if (b > a + 1) {
                bool change = argmax3(a0[g], a1[g], a2[g]) != argmax3(a0[g + 1], a1[g + 1], a2[g + 1]);
                int8_t s0 = sm[a];
                for (int32_t i = a + 1; i <= b && !change; i++) change = sm[i] != s0;
                double dm = fmax(fabs(a0[g] - a0[g + 1]), fmax(fabs(a1[g] - a1[g + 1]), fabs(a2[g] - a2[g + 1])));
                bool full = change || dm > delta;
                gap_full[P.row_off + g] = full ? 1 : 0;
                cnt = full ? (b - a - 1) : 0;
            } else {
                gap_full[P.row_off + g] = 1;  // nothing to fill
            }
        }
        // block exclusive scan of cnt
        s_scan[threadIdx.x] = cnt;
        __syncthreads();
        for (int d = 1; d < (int)blockDim.x; d <<= 1) {
            int v = (threadIdx.x >= (unsigned)d) ? s_scan[threadIdx.x - d] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        int off = s_base + s_scan[threadIdx.x] - cnt;
        for (int t = 0; t < cnt; t++) fill_list[P.row_off + off + t] = a + 1 + t;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_base += s_scan[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) n_fill[w] = s_base;
}

// D8 / D9 ------------------------------------------------------------------------------------------
// One pass that WRITES every row of the three KERN columns and STATE: a sampled row takes its value from the dense sample arrays
// kde_eval_kernel (mode 0) filled, a row of a gap evaluated in full keeps the value kde_eval_kernel (mode 1) wrote, every other row is
// numpy.interp between the two samples around it (`slope * (x - x0) + y0`, from the RAW sampled values); then the spike rule
// (`> 1 -> 1 / x`, density.py:330-332; pandas aligns the right-hand frame on the column, so it is an element-wise reciprocal) and
// STATE = argmax, first maximum wins (:335-338).
// (Why the samples live in their own arrays: with the samples inside the columns, reading one row in `srs` = 20 pulled nearly every
// 128-byte line of the columns through L2 -- ncu, 296 windows: 248 MB read by this kernel, 243 MB by the separate spike pass over the
// sampled rows and 256 MB by gap_classify_kernel, for 15 MB of samples.)
// y / x for an integer-valued x with r = 1 / x already at hand: q = y * r, then one correction with the exact remainder (Markstein: with r the
// correctly rounded reciprocal, fma(fma(-q, x, y), r, q) is y / x correctly rounded -- bit for bit what the division returns). The remainder
// must not underflow for that, so differences below 2^-900 (densities thousands of bandwidths from their data) take the division itself.
__device__ __forceinline__ double div_by_int(double y, double x, double r)
{
    if (fabs(y) < 0x1p-900 && y != 0.0) return y / x;
    const double q = y * r;
    return fma(fma(-q, x, y), r, q);
}

#ifndef FIN_ROWS_N
#define FIN_ROWS_N 2
#endif
constexpr int FIN_ROWS = FIN_ROWS_N;      // rows per thread: the dependent loads (plan, gap flag, values) of FIN_ROWS rows overlap (r02, 296 windows: 1 row 0.205 ms, 2 rows 0.16, 4 rows 0.215, 8 rows 0.38 -- registers)

__global__ void __launch_bounds__(256)
finish_rows_kernel(const WinPlan *__restrict__ plan, int32_t win_base, const uint8_t *__restrict__ gap_full, const double *__restrict__ s0,
                   const double *__restrict__ s1, const double *__restrict__ s2, double *__restrict__ k0, double *__restrict__ k1,
                   double *__restrict__ k2, int8_t *__restrict__ state)
{
    const int32_t w = win_base + blockIdx.y;
    const WinPlan P = plan[w];
    if (P.status != 0) return;
    const int32_t N = P.n_rows;
    const int32_t j0 = blockIdx.x * (256 * FIN_ROWS) + threadIdx.x;
    if (j0 >= N) return;
    if (!P.smoothed) {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
        for (int q = 0; q < FIN_ROWS; q++) {
            const int32_t j = j0 + q * 256;
            if (j < N) { const int64_t r = P.row_off + j; state[r] = -1; k0[r] = nan; k1[r] = nan; k2[r] = nan; }
        }
        return;
    }
    double *ks[3] = {k0 + P.row_off, k1 + P.row_off, k2 + P.row_off};
    const double *ss[3] = {s0 + P.row_off, s1 + P.row_off, s2 + P.row_off};
    int32_t g[FIN_ROWS], a[FIN_ROWS], b[FIN_ROWS];
    bool in[FIN_ROWS], sampled[FIN_ROWS], interp[FIN_ROWS];
#pragma unroll
    for (int q = 0; q < FIN_ROWS; q++) {
        const int32_t j = j0 + q * 256;
        in[q] = j < N;
        g[q] = j / P.srs;
        a[q] = g[q] * P.srs; b[q] = min(a[q] + P.srs, N - 1);
        sampled[q] = (j == a[q]) || (j == N - 1);
        interp[q] = in[q] && !sampled[q] && !gap_full[P.row_off + g[q]];
    }
    double ya[FIN_ROWS][3], yb[FIN_ROWS][3];
#pragma unroll
    for (int q = 0; q < FIN_ROWS; q++) {
        const int32_t j = j0 + q * 256;
        const int32_t e = (j == a[q]) ? g[q] : g[q] + 1;          // a sampled row's entry: its lattice point, or the last sample for row N - 1
#pragma unroll
        for (int s = 0; s < 3; s++) {
            ya[q][s] = !in[q] ? 0.0 : (sampled[q] ? ss[s][e] : (interp[q] ? ss[s][g[q]] : ks[s][j]));
            yb[q][s] = interp[q] ? ss[s][g[q] + 1] : 0.0;
        }
    }
    // x1 - x0 is `srs` for every gap but the last one of a window: one reciprocal per thread instead of three divisions per row
    // (ncu, r02: this kernel was issuing 350 instructions per row, most of them float64 division sequences)
    const double r_srs = 1.0 / (double)P.srs;
    double dx[FIN_ROWS], rdx[FIN_ROWS];
#pragma unroll
    for (int q = 0; q < FIN_ROWS; q++) {
        dx[q] = (double)(b[q] - a[q]);
        rdx[q] = r_srs;
        if (interp[q] && b[q] - a[q] != P.srs) rdx[q] = 1.0 / dx[q];
    }
#pragma unroll
    for (int q = 0; q < FIN_ROWS; q++) {
        const int32_t j = j0 + q * 256;
        if (!in[q]) continue;
        double v[3];
#pragma unroll
        for (int s = 0; s < 3; s++) {
            v[s] = ya[q][s];
            if (interp[q]) {   // numpy.interp: slope = (y1 - y0) / (x1 - x0); slope * (x - x0) + y0, a product and a sum (no FMA)
                const double dy = yb[q][s] - ya[q][s];
                v[s] = __dadd_rn(__dmul_rn(div_by_int(dy, dx[q], rdx[q]), (double)(j - a[q])), ya[q][s]);
            }
            if (v[s] > 1.0) v[s] = 1.0 / v[s];   // density.py:330-332
            ks[s][j] = v[s];
        }
        state[P.row_off + j] = (int8_t)argmax3(v[0], v[1], v[2]);
    }
}

// D10 -----------------------------------------------------------------------------------------------
// Run-length encoding of the final STATE column in row order: (state, count, first INDEX, last INDEX) per run -- exactly the
// tuples pavlib.density.rl_encoder yields (pavlib/density.py:330-361) and all that scan_for_inv looks at to decide whether a
// locus is expanded again (pavlib/inv.py:294-342). One CTA per window; at most STATE_RUN_CAP runs are stored per window (a
// smoothed STATE column has a handful), the true number is always reported. Un-smoothed windows are one run of state -1.
constexpr int STATE_RUN_CAP = 512;    // <= the block size of state_rle_kernel

__global__ void __launch_bounds__(1024)
state_rle_kernel(const WinPlan *__restrict__ plan, const int8_t *__restrict__ state, const int32_t *__restrict__ index,
                 pavgpu_state_run *__restrict__ runs, int32_t *__restrict__ n_state_runs)
{
    const int32_t w = blockIdx.x;
    const WinPlan P = plan[w];
    const int32_t N = (P.status == 0) ? P.n_rows : 0;
    pavgpu_state_run *out = runs + (int64_t)w * STATE_RUN_CAP;
    if (N == 0) { if (threadIdx.x == 0) n_state_runs[w] = 0; return; }
    const int32_t *ix = index + P.row_off;
    if (!P.smoothed) {
        if (threadIdx.x == 0) { out[0] = pavgpu_state_run{-1, N, ix[0], ix[N - 1]}; n_state_runs[w] = 1; }
        return;
    }
    const int8_t *st = state + P.row_off;
    // every thread owns a contiguous slice of the rows: count the run heads in it, scan the counts over the block, write them
    const int32_t per = (N + (int32_t)blockDim.x - 1) / (int32_t)blockDim.x;
    const int32_t lo = min((int32_t)threadIdx.x * per, N), hi = min(lo + per, N);
    int cnt = 0;
    for (int32_t i = lo; i < hi; i++) cnt += (i == 0 || st[i] != st[i - 1]) ? 1 : 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ int s_warp[32];
    __shared__ int s_total;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int before = 0;
    for (int q = 0; q < wid; q++) before += s_warp[q];
    if (threadIdx.x == blockDim.x - 1) s_total = before + inc;
    int32_t r = before + inc - cnt;                        // rank of my first head
    for (int32_t i = lo; i < hi; i++) {
        if (i == 0 || st[i] != st[i - 1]) {
            if (r < STATE_RUN_CAP) { out[r].state = (int32_t)st[i]; out[r].first_index = ix[i]; out[r].count = i; }   // count holds the first row for now
            if (r > 0 && r - 1 < STATE_RUN_CAP) out[r - 1].last_index = ix[i - 1];
            r++;
        }
    }
    __syncthreads();
    const int32_t nr = s_total;
    const int32_t kept = min(nr, STATE_RUN_CAP);
    if (threadIdx.x == 0 && nr <= STATE_RUN_CAP) out[nr - 1].last_index = ix[N - 1];
    // first rows -> counts (run r ends where run r + 1 starts)
    int32_t first = 0, next = -1;
    if ((int32_t)threadIdx.x < kept) {
        first = out[threadIdx.x].count;
        next = ((int32_t)threadIdx.x + 1 < kept) ? out[threadIdx.x + 1].count : -1;
    }
    __syncthreads();
    if ((int32_t)threadIdx.x < kept) out[threadIdx.x].count = (next >= 0 ? next : N) - first;   // (the last stored run of an overflowing window is not used)
    if (threadIdx.x == 0) n_state_runs[w] = nr;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
struct pavgpu_density_batch {
    pavgpu_ctx *ctx;
    int32_t n_win;
    pavgpu_density_params prm;
    std::vector<pavgpu_density_window> win;
    std::vector<WinPlan> plan;
    std::vector<pavgpu_density_result> res;
    int64_t tab_slots, pos_total, tile_total, rows_total, rows_cap;
    int32_t max_ref_tiles, max_tig_tiles, max_row_tiles;
    // device
    WinPlan *d_plan; WinCounts *d_wc; KdeParams *d_kp;
    uint64_t *d_keys; uint32_t *d_counts;
    int8_t *d_st_pos; uint32_t *d_tile_cnt;
    uint64_t *d_kmer; int32_t *d_index; int8_t *d_state_mer, *d_state;
    double *d_k[3], *d_samp[3], *d_tree[3];   // d_samp: the sampled values of a window, densely at [row_off, row_off + n_samp)
    int32_t *d_run_start, *d_run_len, *d_n_runs; int8_t *d_run_state;
    int64_t tree_total;
    uint8_t *d_gap_full; int32_t *d_fill_list, *d_n_fill, *d_n_eval; int64_t *d_grp_off;
    pavgpu_state_run *d_state_runs; int32_t *d_n_state_runs;
    bool ran;
    void *d_arena;    // one allocation backs every device buffer below
    size_t arena_bytes;
    bool allocated;   // device buffers live for the lifetime of the batch (sized from upper bounds on the first run)
    pavgpu_density_stats stats;
};

extern "C" __attribute__((visibility("default"))) void pavgpu_density_default_params(pavgpu_density_params *p)
{
    if (!p) return;
    p->k = 31; p->min_informative = 2000; p->min_state_count = 20; p->max_ref_kmer_count = 100;
    p->smooth = 1.0; p->delta = 0.005;
}

static void dens_release(pavgpu_density_batch *b)
{
    ctx_arena_give(b->ctx, b->d_arena, b->arena_bytes);   // back to the context cache for the next batch
    b->d_arena = nullptr;
}

extern "C" __attribute__((visibility("default"))) void pavgpu_density_batch_free(pavgpu_density_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    dens_release(b);
    delete b;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_density_batch_create(pavgpu_ctx *ctx, int32_t n_win, const pavgpu_density_window *win,
                                                                                    const pavgpu_density_params *params, pavgpu_density_batch **out)
{
    if (!ctx || !out || n_win < 0 || (n_win > 0 && !win) || !params) { pav_set_error("density_batch_create: bad argument"); return PAVGPU_ERR_ARG; }
    if (params->k < 1 || params->k > 31) {
        pav_set_error("density: k=%d is not supported on the GPU path (1 <= k <= 31; exact 62-bit keys)", params->k);
        return PAVGPU_ERR_ARG;
    }
    for (int32_t i = 0; i < n_win; i++) {
        if (win[i].ref_pos < 0 || win[i].ref_end < win[i].ref_pos || win[i].tig_pos < 0 || win[i].tig_end < win[i].tig_pos || win[i].srs < 1) {
            pav_set_error("density: window %d has invalid coordinates or srs", i);
            return PAVGPU_ERR_ARG;
        }
    }
    pavgpu_density_batch *b = new pavgpu_density_batch();
    b->ctx = ctx; b->n_win = n_win; b->prm = *params;
    b->win.assign(win, win + n_win);
    b->plan.resize(n_win);
    b->res.resize(n_win);
    *out = b;
    return PAVGPU_OK;
}

static int log2_ceil(int64_t v)
{
    int l = 0;
    while (((int64_t)1 << l) < v) l++;
    return l;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_density_batch_run(pavgpu_density_batch *b, const pavgpu_seqstore *ref_store,
                                                                                 const pavgpu_seqstore *tig_store, pavgpu_density_stats *stats)
{
    if (!b || !ref_store || !tig_store) { pav_set_error("density_batch_run: bad argument"); return PAVGPU_ERR_ARG; }
    pavgpu_ctx *ctx = b->ctx;
    if (ref_store->ctx->device != ctx->device || tig_store->ctx->device != ctx->device) {
        pav_set_error("density_batch_run: stores live on another device");
        return PAVGPU_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int k = b->prm.k;
    const int32_t n_win = b->n_win;
    memset(&b->stats, 0, sizeof b->stats);
    b->ran = false;
    if (n_win == 0) { b->ran = true; b->rows_total = 0; if (stats) *stats = b->stats; return PAVGPU_OK; }
    PavTrace tr("density_batch_run");

    // ---- plan, part 1 (host)
    int64_t tab = 0, pos = 0, tiles = 0, bases = 0, tree_cap = 0;
    // PAVGPU_DENSITY_ONCHIP=0 sends every window to the global tables (A/B runs and the tests of that path)
    const char *oc_env = getenv("PAVGPU_DENSITY_ONCHIP");
    const bool onchip_on = !(oc_env && oc_env[0] == '0');
    int32_t n_onchip = 0;
    int32_t max_ref_blocks = 1, max_tig_tiles = 1;
    for (int32_t w = 0; w < n_win; w++) {
        const pavgpu_density_window &W = b->win[w];
        if (W.ref_seq_id < 0 || W.ref_seq_id >= ref_store->n_seq || W.tig_seq_id < 0 || W.tig_seq_id >= tig_store->n_seq ||
            W.ref_end > ref_store->h_len[W.ref_seq_id] || W.tig_end > tig_store->h_len[W.tig_seq_id]) {
            pav_set_error("density: window %d is outside its sequence", w);
            return PAVGPU_ERR_ARG;
        }
        WinPlan &P = b->plan[w];
        memset(&P, 0, sizeof P);
        P.ref_g0 = ref_store->h_off[W.ref_seq_id] + W.ref_pos;
        P.tig_g0 = tig_store->h_off[W.tig_seq_id] + W.tig_pos;
        P.ref_len = W.ref_end - W.ref_pos;
        P.tig_len = W.tig_end - W.tig_pos;
        P.rev = W.rev; P.srs = W.srs;
        int32_t n_ref = std::max(P.ref_len - k + 1, 0), n_tig = std::max(P.tig_len - k + 1, 0);
        P.tab_log2 = std::max(log2_ceil(2 * (int64_t)std::max(n_ref, 1)), 4);
        P.tab_off = tab; tab += (int64_t)1 << P.tab_log2;
        P.mode = (onchip_on && n_ref >= 1 && n_ref <= KS_MAX_REF && (n_tig + TILE - 1) / TILE <= KS_MAX_TILES) ? 1 : 0;
        n_onchip += P.mode;
        P.pos_off = pos; pos += n_tig;
        P.tile_off = tiles; tiles += (n_tig + TILE - 1) / TILE;
        max_ref_blocks = std::max(max_ref_blocks, (n_ref + 255) / 256);
        max_tig_tiles = std::max(max_tig_tiles, (n_tig + TILE - 1) / TILE);
        bases += P.tig_len;
        P.tree_off = tree_cap;                                   // upper bound: N <= n_tig
        tree_cap += 2 * (int64_t)std::max(n_tig, 1) + 2;         // T[N] + S[N + 1], N <= n_tig
    }
    b->tab_slots = tab; b->pos_total = pos; b->tile_total = tiles; b->tree_total = tree_cap;
    b->stats.bases = bases;

    int launches = 0;
    if (!b->allocated) {   // everything is sized from upper bounds (rows <= contig k-mer positions), so runs never allocate;
                           // one arena, one cudaMalloc (a dozen separate allocations cost milliseconds per call)
        size_t rcap = (size_t)std::max<int64_t>(pos, 1), tcap = (size_t)std::max<int64_t>(tab, 1);
        size_t off = 0;
        auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        size_t o_plan = carve(sizeof(WinPlan) * n_win), o_wc = carve(sizeof(WinCounts) * n_win), o_kp = carve(sizeof(KdeParams) * n_win);
        size_t o_keys = carve(8 * tcap), o_counts = carve(4 * tcap), o_st = carve(rcap), o_tile = carve(12 * (size_t)std::max<int64_t>(tiles, 1));
        size_t o_kmer = carve(8 * rcap), o_index = carve(4 * rcap), o_sm = carve(rcap), o_state = carve(rcap);
        size_t o_k[3], o_samp[3], o_tree[3];
        for (int s = 0; s < 3; s++) { o_k[s] = carve(8 * rcap); o_samp[s] = carve(8 * rcap); o_tree[s] = carve(8 * (size_t)std::max<int64_t>(tree_cap, 1)); }
        size_t o_rs = carve(4 * rcap), o_rl = carve(4 * rcap), o_rst = carve(rcap), o_nr = carve(4 * (size_t)n_win);
        size_t o_gap = carve(rcap), o_fill = carve(4 * rcap), o_nf = carve(4 * (size_t)n_win), o_ne = carve(4 * (size_t)n_win);
        size_t o_grp = carve(8 * ((size_t)n_win + 1));
        size_t o_sruns = carve(sizeof(pavgpu_state_run) * STATE_RUN_CAP * (size_t)n_win), o_nsr = carve(4 * (size_t)n_win);
        CUDA_TRY(ctx_arena_take(ctx, off, &b->d_arena, &b->arena_bytes));
        char *base = static_cast<char *>(b->d_arena);
        b->d_plan = (WinPlan *)(base + o_plan); b->d_wc = (WinCounts *)(base + o_wc); b->d_kp = (KdeParams *)(base + o_kp);
        b->d_keys = (uint64_t *)(base + o_keys); b->d_counts = (uint32_t *)(base + o_counts); b->d_st_pos = (int8_t *)(base + o_st);
        b->d_tile_cnt = (uint32_t *)(base + o_tile);
        b->d_kmer = (uint64_t *)(base + o_kmer); b->d_index = (int32_t *)(base + o_index);
        b->d_state_mer = (int8_t *)(base + o_sm); b->d_state = (int8_t *)(base + o_state);
        for (int s = 0; s < 3; s++) { b->d_k[s] = (double *)(base + o_k[s]); b->d_samp[s] = (double *)(base + o_samp[s]); b->d_tree[s] = (double *)(base + o_tree[s]); }
        b->d_run_start = (int32_t *)(base + o_rs); b->d_run_len = (int32_t *)(base + o_rl); b->d_run_state = (int8_t *)(base + o_rst);
        b->d_n_runs = (int32_t *)(base + o_nr);
        b->d_gap_full = (uint8_t *)(base + o_gap); b->d_fill_list = (int32_t *)(base + o_fill);
        b->d_n_fill = (int32_t *)(base + o_nf); b->d_n_eval = (int32_t *)(base + o_ne); b->d_grp_off = (int64_t *)(base + o_grp);
        b->d_state_runs = (pavgpu_state_run *)(base + o_sruns); b->d_n_state_runs = (int32_t *)(base + o_nsr);
        b->allocated = true;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev[0], st));
    CUDA_TRY(cudaEventRecord(ctx->ev[1], st));   // the table initialisation below is part of the step
    CUDA_TRY(cudaMemcpyAsync(b->d_plan, b->plan.data(), sizeof(WinPlan) * n_win, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(b->d_wc, 0, sizeof(WinCounts) * n_win, st));
    const int32_t YMAX = 32768;
    // k-mer part: windows small enough for kmer_window_kernel build and probe their table in shared memory (mode 1); the others, and
    // the windows that kernel hands back (a k-mer count its 6 bits cannot decide), go through the tables in HBM (mode 0).
    std::vector<WinCounts> wc(n_win);
    auto run_global = [&](bool whole_slab) -> int {
        if (whole_slab) {
            CUDA_TRY(cudaMemsetAsync(b->d_keys, 0xFF, sizeof(uint64_t) * std::max<int64_t>(tab, 1), st));
            CUDA_TRY(cudaMemsetAsync(b->d_counts, 0, sizeof(uint32_t) * std::max<int64_t>(tab, 1), st));
        } else {
            for (int32_t w = 0; w < n_win; w++) {
                const WinPlan &P = b->plan[w];
                if (P.mode != 0) continue;
                CUDA_TRY(cudaMemsetAsync(b->d_keys + P.tab_off, 0xFF, sizeof(uint64_t) << P.tab_log2, st));
                CUDA_TRY(cudaMemsetAsync(b->d_counts + P.tab_off, 0, sizeof(uint32_t) << P.tab_log2, st));
            }
        }
        for (int32_t w0 = 0; w0 < n_win; w0 += YMAX) {
            int32_t ny = std::min(YMAX, n_win - w0);
            ref_insert_kernel<<<dim3(max_ref_blocks, ny), 256, 0, st>>>(b->d_plan, w0, planes_of(ref_store), k, b->d_keys, b->d_counts, b->d_wc);
            launches++;
        }
        for (int32_t w0 = 0; w0 < n_win; w0 += YMAX) {
            int32_t ny = std::min(YMAX, n_win - w0);
            tig_state_kernel<<<dim3(max_tig_tiles, ny), TILE, 0, st>>>(b->d_plan, w0, planes_of(tig_store), k, b->d_keys, b->d_st_pos, b->d_tile_cnt, b->d_wc);
            launches++;
        }
        CUDA_TRY(cudaGetLastError());
        return PAVGPU_OK;
    };
    if (n_onchip > 0) {
        static bool attr_done[64] = {};
        const int dev = ctx->device;
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(kmer_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_SMEM));
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
        const unsigned cnt_limit = (unsigned)std::max<int64_t>(std::min<int64_t>(b->prm.max_ref_kmer_count, KS_CNT_LIMIT), 0);
        kmer_window_kernel<<<(unsigned)n_win, KS_THREADS, KS_SMEM, st>>>(b->d_plan, planes_of(ref_store), planes_of(tig_store), k, cnt_limit,
                                                                         b->d_st_pos, b->d_tile_cnt, b->d_wc);
        launches++;
        CUDA_TRY(cudaGetLastError());
    }
    if (n_onchip < n_win) { int rc = run_global(n_onchip == 0 || (n_win - n_onchip) > 64); if (rc != PAVGPU_OK) return rc; }
    CUDA_TRY(cudaMemcpyAsync(wc.data(), b->d_wc, sizeof(WinCounts) * n_win, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int32_t n_redo = 0;
    for (int32_t w = 0; w < n_win; w++) n_redo += (b->plan[w].mode == 1 && wc[w].overflow) ? 1 : 0;
    if (n_redo > 0) {
        for (int32_t w = 0; w < n_win; w++) {
            WinPlan &P = b->plan[w];
            if (P.mode == 1 && wc[w].overflow) { P.mode = 0; memset(&wc[w], 0, sizeof(WinCounts)); }
            else P.mode = 2;
        }
        CUDA_TRY(cudaMemcpyAsync(b->d_plan, b->plan.data(), sizeof(WinPlan) * n_win, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b->d_wc, wc.data(), sizeof(WinCounts) * n_win, cudaMemcpyHostToDevice, st));
        { int rc = run_global(n_redo > 64); if (rc != PAVGPU_OK) return rc; }
        CUDA_TRY(cudaMemcpyAsync(wc.data(), b->d_wc, sizeof(WinCounts) * n_win, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    b->stats.kmer_tables_on_chip = n_onchip - n_redo;
    tr.mark("plan + k-mer kernels");

    // ---- plan, part 2 (host): status, keep mask, N, dense row offsets, sample counts
    int64_t rows = 0;
    int32_t max_rows = 1;
    std::vector<int64_t> grp_off(n_win + 1, 0);
    std::vector<int32_t> n_eval(n_win, 0);
    for (int32_t w = 0; w < n_win; w++) {
        WinPlan &P = b->plan[w];
        pavgpu_density_result &R = b->res[w];
        memset(&R, 0, sizeof R);
        if (wc[w].ref_valid == 0 || (int64_t)wc[w].ref_max > (int64_t)b->prm.max_ref_kmer_count) {  // density.py:510-527
            P.status = PAVGPU_INV_FAIL; P.n_rows = 0; P.smoothed = 0; P.keep_mask = 0;
        } else {
            P.status = 0; P.keep_mask = 0;
            int64_t N = 0;
            for (int s = 0; s < 3; s++)
                if (wc[w].cnt[s] > 0 && (int64_t)wc[w].cnt[s] >= (int64_t)b->prm.min_state_count) { P.keep_mask |= 1 << s; N += wc[w].cnt[s]; }
            P.n_rows = (int32_t)N;
            P.smoothed = (N >= b->prm.min_informative && N > 0) ? 1 : 0;  // density.py:193-194
        }
        P.row_off = rows;
        rows += P.n_rows;
        max_rows = std::max(max_rows, P.n_rows);
        P.n_samp = 0; P.npad = 0;
        if (P.smoothed) {
            int32_t N = P.n_rows;
            P.npad = 1 << log2_ceil(N);   // <= the capacity reserved at tree_off
            P.n_samp = (N - 1) / P.srs + 1 + (((N - 1) % P.srs) ? 1 : 0);  // density.py:211-214
        }
        n_eval[w] = P.n_samp;
        grp_off[w + 1] = grp_off[w] + (P.n_samp + EVAL_GROUP - 1) / EVAL_GROUP;
        R.status = P.status; R.smoothed = P.smoothed; R.row_off = P.row_off; R.n_rows = P.n_rows; R.n_eval = P.n_samp;
    }
    b->rows_total = rows;
    b->stats.rows = rows;
    CUDA_TRY(cudaMemcpyAsync(b->d_plan, b->plan.data(), sizeof(WinPlan) * n_win, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_n_eval, n_eval.data(), sizeof(int32_t) * n_win, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_grp_off, grp_off.data(), sizeof(int64_t) * (n_win + 1), cudaMemcpyHostToDevice, st));

    int32_t row_blocks = (max_rows + 255) / 256;
    for (int32_t w0 = 0; w0 < n_win; w0 += YMAX) {
        int32_t ny = std::min(YMAX, n_win - w0);
        compact_kernel<<<dim3(max_tig_tiles, ny), TILE, 0, st>>>(b->d_plan, w0, planes_of(tig_store), k, b->d_st_pos, b->d_tile_cnt, b->d_kmer,
                                                                b->d_index, b->d_state_mer);
        launches++;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev[2], st));
    runs_stats_kernel<<<n_win, 1024, 0, st>>>(b->d_plan, b->d_state_mer, b->prm.smooth, b->d_run_start, b->d_run_len, b->d_run_state, b->d_n_runs,
                                              b->d_kp);
    kde_table_kernel<<<n_win * 3, TREE_THREADS, 0, st>>>(b->d_plan, b->d_kp, b->d_tree[0], b->d_tree[1], b->d_tree[2]);
    launches += 2;
    CUDA_TRY(cudaGetLastError());
    int64_t pairs = 0;
    if (grp_off[n_win] > 0) {
        kde_eval_kernel<<<(unsigned)grp_off[n_win], EVAL_THREADS, 0, st>>>(b->d_plan, n_win, b->d_grp_off, b->d_n_eval, 0, nullptr, b->d_kp,
                                                                           b->d_run_start, b->d_run_len, b->d_run_state, b->d_n_runs, b->d_tree[0],
                                                                           b->d_tree[1], b->d_tree[2], b->d_samp[0], b->d_samp[1], b->d_samp[2]);
        launches++;
        CUDA_TRY(cudaGetLastError());
    }
    gap_classify_kernel<<<n_win, 256, 0, st>>>(b->d_plan, b->prm.delta, b->d_state_mer, b->d_samp[0], b->d_samp[1], b->d_samp[2], b->d_gap_full,
                                               b->d_fill_list, b->d_n_fill);
    launches++;
    CUDA_TRY(cudaGetLastError());
    std::vector<int32_t> n_fill(n_win, 0);
    CUDA_TRY(cudaMemcpyAsync(n_fill.data(), b->d_n_fill, sizeof(int32_t) * n_win, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    tr.mark("compact + KDE sampled");
    for (int32_t w = 0; w < n_win; w++) {
        pairs += ((int64_t)b->plan[w].n_samp + n_fill[w]) * b->plan[w].n_rows;
        b->res[w].n_eval += n_fill[w];
        grp_off[w + 1] = grp_off[w] + (n_fill[w] + EVAL_GROUP - 1) / EVAL_GROUP;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev[3], st));
    if (grp_off[n_win] > 0) {
        CUDA_TRY(cudaMemcpyAsync(b->d_grp_off, grp_off.data(), sizeof(int64_t) * (n_win + 1), cudaMemcpyHostToDevice, st));
        kde_eval_kernel<<<(unsigned)grp_off[n_win], EVAL_THREADS, 0, st>>>(b->d_plan, n_win, b->d_grp_off, b->d_n_fill, 1, b->d_fill_list,
                                                                           b->d_kp, b->d_run_start, b->d_run_len, b->d_run_state, b->d_n_runs,
                                                                           b->d_tree[0], b->d_tree[1], b->d_tree[2], b->d_k[0], b->d_k[1], b->d_k[2]);
        launches++;
        CUDA_TRY(cudaGetLastError());
    }
    for (int32_t w0 = 0; w0 < n_win; w0 += YMAX) {
        int32_t ny = std::min(YMAX, n_win - w0);
        finish_rows_kernel<<<dim3((row_blocks + FIN_ROWS - 1) / FIN_ROWS, ny), 256, 0, st>>>(b->d_plan, w0, b->d_gap_full, b->d_samp[0], b->d_samp[1],
                                                                                              b->d_samp[2], b->d_k[0], b->d_k[1], b->d_k[2], b->d_state);
        launches++;
    }
    CUDA_TRY(cudaGetLastError());
    state_rle_kernel<<<n_win, 1024, 0, st>>>(b->d_plan, b->d_state, b->d_index, b->d_state_runs, b->d_n_state_runs);
    launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    b->stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    b->stats.ms_kmer = ev_ms(ctx->ev[1], ctx->ev[2]);
    b->stats.ms_kde = ev_ms(ctx->ev[2], ctx->ev[4]);
    b->stats.ms_fill = ev_ms(ctx->ev[3], ctx->ev[4]);
    b->stats.ms_kernels = ev_ms(ctx->ev[1], ctx->ev[4]);
    b->stats.kde_pairs = pairs;
    b->stats.kernel_launches = launches;
    b->ran = true;
    if (stats) *stats = b->stats;
    return PAVGPU_OK;
}

template <typename T>
static int fetch_col(pavgpu_ctx *ctx, const T *d, int64_t n, T **out)
{
    *out = nullptr;
    if (n == 0) return PAVGPU_OK;
    T *h = nullptr;   // pinned result buffer from the context's pool (released with pavgpu_free_host)
    int prc = pav_pinned_take(ctx, (size_t)n * sizeof(T), reinterpret_cast<void **>(&h));
    if (prc) return prc;
    cudaError_t e = cudaMemcpyAsync(h, d, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) { pavgpu_free_host(h); pav_set_error("density fetch: %s", cudaGetErrorString(e)); return PAVGPU_ERR_CUDA; }
    *out = h;
    return PAVGPU_OK;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_density_batch_fetch(pavgpu_density_batch *b, pavgpu_density_result *res, uint64_t **kmer_out,
                                                                                   int32_t **index_out, int8_t **state_mer_out, int8_t **state_out,
                                                                                   double **kern_fwd_out, double **kern_fwdrev_out,
                                                                                   double **kern_rev_out, int64_t *n_rows_total)
{
    if (!b || !b->ran || !res || !kmer_out || !index_out || !state_mer_out || !state_out || !kern_fwd_out || !kern_fwdrev_out || !kern_rev_out ||
        !n_rows_total) {
        pav_set_error("density_batch_fetch: bad argument or batch not run");
        return PAVGPU_ERR_ARG;
    }
    pavgpu_ctx *ctx = b->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (int32_t w = 0; w < b->n_win; w++) res[w] = b->res[w];
    int64_t n = b->rows_total;
    *n_rows_total = n;
    PavTrace tr("density_batch_fetch");
    CUDA_TRY(cudaEventRecord(ctx->ev[5], ctx->stream));
    int rc = 0;
    rc |= fetch_col(ctx, b->d_kmer, n, kmer_out);
    rc |= fetch_col(ctx, b->d_index, n, index_out);
    rc |= fetch_col(ctx, b->d_state_mer, n, state_mer_out);
    rc |= fetch_col(ctx, b->d_state, n, state_out);
    rc |= fetch_col(ctx, b->d_k[0], n, kern_fwd_out);
    rc |= fetch_col(ctx, b->d_k[1], n, kern_fwdrev_out);
    rc |= fetch_col(ctx, b->d_k[2], n, kern_rev_out);
    CUDA_TRY(cudaEventRecord(ctx->ev[6], ctx->stream));
    tr.mark("pinned buffers + enqueue");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    tr.mark("d2h");
    b->stats.ms_d2h = ev_ms(ctx->ev[5], ctx->ev[6]);
    if (rc) {
        pavgpu_free_host(*kmer_out); pavgpu_free_host(*index_out); pavgpu_free_host(*state_mer_out); pavgpu_free_host(*state_out);
        pavgpu_free_host(*kern_fwd_out); pavgpu_free_host(*kern_fwdrev_out); pavgpu_free_host(*kern_rev_out);
        return PAVGPU_ERR_NOMEM;
    }
    return PAVGPU_OK;
}

// Run lengths of STATE for every window (what scan_for_inv decides from): 16 bytes per run instead of 38 bytes per row.
extern "C" __attribute__((visibility("default"))) int pavgpu_density_batch_fetch_runs(pavgpu_density_batch *b, pavgpu_density_result *res,
                                                                                        pavgpu_state_run **runs_out, int64_t *run_off, int64_t *n_runs_total)
{
    if (!b || !b->ran || !res || !runs_out || !run_off || !n_runs_total) { pav_set_error("density_batch_fetch_runs: bad argument or batch not run"); return PAVGPU_ERR_ARG; }
    pavgpu_ctx *ctx = b->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int32_t n_win = b->n_win;
    *runs_out = nullptr; *n_runs_total = 0;
    for (int32_t w = 0; w < n_win; w++) res[w] = b->res[w];
    run_off[0] = 0;
    if (n_win == 0) return PAVGPU_OK;
    PavTrace tr("density_batch_fetch_runs");
    std::vector<int32_t> n_sr(n_win);
    CUDA_TRY(cudaEventRecord(ctx->ev[5], ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(n_sr.data(), b->d_n_state_runs, 4 * (size_t)n_win, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    int64_t total = 0;
    for (int32_t w = 0; w < n_win; w++) { run_off[w] = total; total += n_sr[w]; }
    run_off[n_win] = total;
    *n_runs_total = total;
    if (total == 0) return PAVGPU_OK;
    pavgpu_state_run *h = nullptr, *blk = nullptr;
    int prc = pav_pinned_take(ctx, (size_t)total * sizeof(pavgpu_state_run), reinterpret_cast<void **>(&h));
    if (prc) return prc;
    // one copy of the whole per-window run block (n_win x STATE_RUN_CAP x 16 B: a few MB) instead of one small copy per window
    prc = pav_pinned_take(ctx, (size_t)n_win * STATE_RUN_CAP * sizeof(pavgpu_state_run), reinterpret_cast<void **>(&blk));
    if (prc) { pavgpu_free_host(h); return prc; }
    int rc = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(blk, b->d_state_runs, (size_t)n_win * STATE_RUN_CAP * sizeof(pavgpu_state_run), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev[6], ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        for (int32_t w = 0; w < n_win; w++)
            if (n_sr[w] > 0 && n_sr[w] <= STATE_RUN_CAP)
                memcpy(h + run_off[w], blk + (size_t)w * STATE_RUN_CAP, (size_t)n_sr[w] * sizeof(pavgpu_state_run));
        // a window with more runs than the device keeps (never seen on smoothed columns): encode its STATE column here
        for (int32_t w = 0; w < n_win; w++) {
            if (n_sr[w] <= STATE_RUN_CAP) continue;
            const int64_t N = b->res[w].n_rows, r0 = b->res[w].row_off;
            std::vector<int8_t> st((size_t)N);
            std::vector<int32_t> ix((size_t)N);
            CUDA_TRY(cudaMemcpy(st.data(), b->d_state + r0, (size_t)N, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(ix.data(), b->d_index + r0, (size_t)N * 4, cudaMemcpyDeviceToHost));
            pavgpu_state_run *o = h + run_off[w];
            int64_t r = -1;
            for (int64_t i = 0; i < N; i++) {
                if (i == 0 || st[i] != st[i - 1]) { r++; o[r] = pavgpu_state_run{(int32_t)st[i], 0, ix[i], ix[i]}; }
                o[r].count++; o[r].last_index = ix[i];
            }
        }
        return PAVGPU_OK;
    }();
    tr.mark("d2h");
    pavgpu_free_host(blk);
    if (rc) { pavgpu_free_host(h); return rc; }
    b->stats.ms_d2h = ev_ms(ctx->ev[5], ctx->ev[6]);
    *runs_out = h;
    return PAVGPU_OK;
}

// All columns of one window into caller-owned arrays of res[win].n_rows entries (any pointer may be NULL): the window that becomes a call.
extern "C" __attribute__((visibility("default"))) int pavgpu_density_batch_fetch_window(pavgpu_density_batch *b, int32_t win, uint64_t *kmer, int32_t *index,
                                                                                          int8_t *state_mer, int8_t *state, double *kern_fwd,
                                                                                          double *kern_fwdrev, double *kern_rev)
{
    if (!b || !b->ran || win < 0 || win >= b->n_win) { pav_set_error("density_batch_fetch_window: bad argument or batch not run"); return PAVGPU_ERR_ARG; }
    pavgpu_ctx *ctx = b->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t n = b->res[win].n_rows, r0 = b->res[win].row_off;
    if (n == 0) return PAVGPU_OK;
    cudaStream_t st = ctx->stream;
    // The caller's arrays are ordinary memory: seven device-to-pageable copies are seven synchronous staged transfers (0.29 ms per
    // window, r02 profile of call_inv_batch). One pinned block from the context's pool takes all columns with asynchronous copies and
    // one synchronisation; the host memcpy out of it is ~38 bytes per row.
    const size_t nn = (size_t)n;
    const size_t off_kmer = 0, off_k0 = off_kmer + 8 * nn, off_k1 = off_k0 + 8 * nn, off_k2 = off_k1 + 8 * nn, off_index = off_k2 + 8 * nn,
                 off_sm = off_index + 4 * nn, off_st = off_sm + nn, total = off_st + nn;
    char *h = nullptr;
    int prc = pav_pinned_take(ctx, total, reinterpret_cast<void **>(&h));
    if (prc) return prc;
    cudaError_t e = cudaSuccess;
    auto cp = [&](void *want, size_t off, const void *src, size_t bytes) {
        if (want && e == cudaSuccess) e = cudaMemcpyAsync(h + off, src, bytes, cudaMemcpyDeviceToHost, st);
    };
    cp(kmer, off_kmer, b->d_kmer + r0, 8 * nn);
    cp(kern_fwd, off_k0, b->d_k[0] + r0, 8 * nn);
    cp(kern_fwdrev, off_k1, b->d_k[1] + r0, 8 * nn);
    cp(kern_rev, off_k2, b->d_k[2] + r0, 8 * nn);
    cp(index, off_index, b->d_index + r0, 4 * nn);
    cp(state_mer, off_sm, b->d_state_mer + r0, nn);
    cp(state, off_st, b->d_state + r0, nn);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { pavgpu_free_host(h); pav_set_error("density_batch_fetch_window: %s", cudaGetErrorString(e)); return PAVGPU_ERR_CUDA; }
    if (kmer) memcpy(kmer, h + off_kmer, 8 * nn);
    if (kern_fwd) memcpy(kern_fwd, h + off_k0, 8 * nn);
    if (kern_fwdrev) memcpy(kern_fwdrev, h + off_k1, 8 * nn);
    if (kern_rev) memcpy(kern_rev, h + off_k2, 8 * nn);
    if (index) memcpy(index, h + off_index, 4 * nn);
    if (state_mer) memcpy(state_mer, h + off_sm, nn);
    if (state) memcpy(state, h + off_st, nn);
    pavgpu_free_host(h);
    return PAVGPU_OK;
}
