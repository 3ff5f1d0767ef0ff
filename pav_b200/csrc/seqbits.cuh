// seqbits.cuh -- bit-level primitives over the packed sequence planes: base codes and word packing, 32-base windows,
// the word-parallel homology scans, k-mers. Device code (every function is __device__ __forceinline__ under nvcc); the same
// text also compiles as plain C++ when PAV_DEV is predefined and the handful of CUDA intrinsics it uses (__ldg, __funnelshift_r,
// __byte_perm, __brev, __clz, __clzll, __ffs, __ffsll, min) are supplied by the includer -- tests/host_emul does that to run
// these exact functions against the oracle and the golden vectors on a machine without a GPU.
#pragma once
#include <cstdint>

#ifndef PAV_DEV
#define PAV_DEV __device__ __forceinline__
#endif

// ---- base codes and plane words ----------------------------------------------------------------
// A/a 0, C/c 1, G/g 2, T/t 3, everything else 4
PAV_DEV uint32_t base_code(uint32_t ch)
{
    ch &= 0xDFu;  // fold case
    return ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
}

// 32 ASCII bases (eight little-endian 32-bit loads) -> one 2-bit plane word (first base most significant) and one mask word
// (bit i set when base i is not ACGTacgt).
PAV_DEV void pack_word32(const uint32_t (&v)[8], uint64_t &word, uint32_t &mask)
{
    word = 0;
    mask = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t c = base_code((v[i] >> (8 * j)) & 0xFFu);
            int pos = i * 4 + j;
            word |= (uint64_t)(c & 3u) << (62 - 2 * pos);
            mask |= (c >> 2) << pos;
        }
    }
}

// ---- oriented sequences, 32-base windows, homology scans ---------------------------------------
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
    // Optional N summary of the mask plane (one bit per 8 mask words = 256 bases, set when any of them is non-zero; nullptr: none).
    // A window whose blocks are clear has an all-zero mask and can skip its two mask loads. Used by the k-mer kernels (-1.4 % on the
    // C5 k-mer part); the homology kernels leave it unset: measured on C2 (B200, r02) the dependent summary load in front of the mask
    // loads costs more than the two skipped sectors save (gather 0.122 against 0.094 ms, CTA queue 0.097 against 0.087 ms).
    const uint32_t *nsum;
};

// Mask words w and w+1 may hold set bits (always true without a summary).
PAV_DEV bool nsum_any2(const uint32_t *__restrict__ nsum, int64_t w)
{
#ifdef PAV_NO_NSUM   // A/B builds: mask loads never skipped
    return true;
#endif
    if (!nsum) return true;
    const int64_t b0 = w >> 3, b1 = (w + 1) >> 3;
    const uint32_t s0 = __ldg(nsum + (b0 >> 5));
    uint32_t any = (s0 >> (b0 & 31)) & 1u;
    if (b1 != b0) {
        const uint32_t s1 = (b1 >> 5) == (b0 >> 5) ? s0 : __ldg(nsum + (b1 >> 5));
        any |= (s1 >> (b1 & 31)) & 1u;
    }
    return any != 0;
}

// Upper-cased base as 0..3 (ACGT) or 4 (anything else, or out of range).
PAV_DEV int oseq_base(const OSeq &s, int64_t t)
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
PAV_DEV uint64_t revcomp32(uint64_t x)
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
PAV_DEV void fwd_window(const OSeq &s, int32_t f, uint64_t &bases, uint32_t &mask)
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
        if (nsum_any2(s.nsum, w)) { m0 = __ldg(s.nmask + w); m1 = __ldg(s.nmask + w + 1); } else { m0 = 0; m1 = 0; }
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
PAV_DEV void oseq_window(const OSeq &s, int32_t t, uint64_t &bases, uint32_t &mask)
{
    if (!s.rev) { fwd_window<TILED>(s, t, bases, mask); return; }
    uint64_t b; uint32_t m;
    fwd_window<TILED>(s, (int32_t)s.len - t - 32, b, m);
    bases = revcomp32(b);
    mask = __brev(m);
}

// HOM_BATCH_LOADS = 1 routes the first trips and the rests of score_indel2 through the prepared / loaded / finished windows below
// (all loads of a phase in flight together). Measured on C2 (B200, r02): slower than the plain windows at every register budget
// (queue kernel 0.092-0.104 ms against 0.0755 ms; 64 / 80 / 128 registers) -- the kernel is bound by DRAM row activations of
// scattered sectors, not by the number of dependent round trips per thread, and the selects cost instructions. Kept for A/B builds.
#ifndef HOM_BATCH_LOADS
#define HOM_BATCH_LOADS 0
#endif

// The same window in three steps, so that a caller can issue the loads of several windows before it looks at any of them:
// win_prep (all address arithmetic, no branch on data), win_load (four loads), win_finish (shifts, sequence edges, strand -- selects
// only). With the windows of a phase prepared, loaded and finished in that order the compiler keeps all their loads in flight
// together; through oseq_window every window waits for its own loads behind the branches of the one before (ncu, r02: a third of
// the homology kernel's warp time was spent on those serial round trips).
struct WinReq {
    int64_t w;              // first plane word
    int32_t sh, lead, over; // base offset inside the word; window positions before the sequence start / past its end
    uint32_t none, rev;     // no position of the window lies inside the sequence; reverse-complement view
};
struct WinRaw {
    uint64_t hi, lo;
    uint32_t m0, m1;
};

PAV_DEV WinReq win_prep(const OSeq &s, int32_t t)
{
    const int32_t len = (int32_t)s.len;
    const int32_t f = s.rev ? len - t - 32 : t;
    WinReq q;
    q.none = (f <= -32 || f >= len) ? 1u : 0u;
    q.lead = (f < 0 && !q.none) ? -f : 0;
    q.over = (!q.none && f + 32 > len) ? f + 32 - len : 0;
    const int64_t g = s.base + (q.none ? 0 : (int64_t)(f + q.lead));
    q.w = g >> 5;
    q.sh = (int32_t)(g & 31);
    q.rev = (uint32_t)s.rev;
    return q;
}

PAV_DEV WinRaw win_load(const OSeq &s, const WinReq &q)
{
    WinRaw r;
    r.hi = __ldg(s.pack2 + q.w); r.lo = __ldg(s.pack2 + q.w + 1);
    if (nsum_any2(s.nsum, q.w)) { r.m0 = __ldg(s.nmask + q.w); r.m1 = __ldg(s.nmask + q.w + 1); } else { r.m0 = 0; r.m1 = 0; }
    return r;
}

PAV_DEV void win_finish(const WinReq &q, const WinRaw &r, uint64_t &bases, uint32_t &mask)
{
    uint64_t b = q.sh ? ((r.hi << (2 * q.sh)) | (r.lo >> (64 - 2 * q.sh))) : r.hi;
    uint32_t m = __funnelshift_r(r.m0, r.m1, q.sh);
    if (q.lead) { b >>= 2 * q.lead; m = (m << q.lead) | ((1u << q.lead) - 1u); }
    if (q.over) m |= ~0u << (32 - q.over);
    if (q.none) { b = 0; m = 0xffffffffu; }
    if (q.rev) { b = revcomp32(b); m = __brev(m); }
    bases = b; mask = m;
}

// Matching bases of two finished windows from the scan's near end (left != 0: from base 31 down), 0..32.
PAV_DEV int window_stop(uint64_t wa, uint32_t ma, uint64_t wb, uint32_t mb, int left)
{
    const uint64_t x = wa ^ wb;
    const uint64_t d = (x | (x >> 1)) & 0x5555555555555555ull;      // one bit per differing base
    const uint32_t m = ma | mb;
    int stop_d, stop_m;
    if (left) {   // last base of the window = least significant group / highest mask bit
        stop_d = d ? ((__ffsll((long long)d) - 1) >> 1) : 32;
        stop_m = m ? __clz((int)m) : 32;
    } else {      // first base = most significant group / lowest mask bit
        stop_d = d ? (__clzll((long long)d) >> 1) : 32;
        stop_m = m ? (__ffs((int)m) - 1) : 32;
    }
    return min(stop_d, stop_m);
}

// common_extension with the windows of two steps (64 bases) in flight per iteration.
PAV_DEV int32_t common_extension2(const OSeq &A, int32_t a, const OSeq &B, int32_t b, int32_t limit, int left)
{
    int32_t h = 0;
    const int32_t step = left ? -32 : 32;
    int32_t pa = left ? a - 31 : a, pb = left ? b - 31 : b;
    while (h < limit) {
        const WinReq qa0 = win_prep(A, pa), qb0 = win_prep(B, pb), qa1 = win_prep(A, pa + step), qb1 = win_prep(B, pb + step);
        const WinRaw ra0 = win_load(A, qa0), rb0 = win_load(B, qb0), ra1 = win_load(A, qa1), rb1 = win_load(B, qb1);
        uint64_t wa, wb; uint32_t ma, mb;
        win_finish(qa0, ra0, wa, ma); win_finish(qb0, rb0, wb, mb);
        int stop = window_stop(wa, ma, wb, mb, left);
        if (stop < 32) { h += stop; return h < limit ? h : limit; }
        h += 32;
        if (h >= limit) return limit;
        win_finish(qa1, ra1, wa, ma); win_finish(qb1, rb1, wb, mb);
        stop = window_stop(wa, ma, wb, mb, left);
        if (stop < 32) { h += stop; return h < limit ? h : limit; }
        h += 32; pa += 2 * step; pb += 2 * step;
        if (h < 0) return limit;   // (cannot happen for sequences < 2^31; guards the 32-bit counter)
    }
    return limit;
}

// Longest common extension of A from a and B from b, 32 bases per step, capped at `limit`:
//   left == 0: common prefix of A[a..] and B[b..]        (window i covers a+32i .. a+32i+31)
//   left != 0: common suffix of A[..a] and B[..b]         (window i covers a-32i-31 .. a-32i)
// Stops at the first mismatch, non-ACGT base or sequence end on either side.
template <bool TILED = false>
PAV_DEV int32_t common_extension(const OSeq &A, int32_t a, const OSeq &B, int32_t b, int32_t limit, int left)
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
PAV_DEV int dev_homology_raw(const uint64_t *t_pack2, const uint32_t *t_nmask, int64_t t_base, int64_t t_len, int t_rev,
                                                    int64_t p, const uint64_t *v_pack2, const uint32_t *v_nmask, int64_t v_base, int64_t v_len,
                                                    int v_rev, int64_t v0, int n, int left)
{
    const OSeq T{t_pack2, t_nmask, t_base, t_len, t_rev, nullptr, nullptr, 0, 0, nullptr};
    const OSeq V{v_pack2, v_nmask, v_base, v_len, v_rev, nullptr, nullptr, 0, 0, nullptr};
    if (n <= 0 || p < 0 || p >= T.len) return 0;
    const int32_t p32 = (int32_t)p, v32 = (int32_t)v0;
    int32_t h = common_extension(T, p32, V, left ? v32 + n - 1 : v32, n, left);
    if (h < n) return h;
    // the flank is shorter than 2^31, so the self-comparison ends at a sequence edge long before the cap
    return n + common_extension(T, left ? p32 - n : p32 + n, T, p32, 0x7fffffff - n, left);
}

// The same scan over sequences that may carry a staged tile.
PAV_DEV int dev_homology_tiled(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n, int left)
{
    if (n <= 0 || p < 0 || p >= T.len) return 0;
    const int32_t p32 = (int32_t)p, v32 = (int32_t)v0;
    int32_t h = common_extension<true>(T, p32, V, left ? v32 + n - 1 : v32, n, left);
    if (h < n) return h;
    return n + common_extension<true>(T, left ? p32 - n : p32 + n, T, p32, 0x7fffffff - n, left);
}

PAV_DEV int dev_left_homology(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n)
{
    return dev_homology_raw(T.pack2, T.nmask, T.base, T.len, T.rev, p, V.pack2, V.nmask, V.base, V.len, V.rev, v0, n, 1);
}

PAV_DEV int dev_right_homology(const OSeq &T, int64_t p, const OSeq &V, int64_t v0, int n)
{
    return dev_homology_raw(T.pack2, T.nmask, T.base, T.len, T.rev, p, V.pack2, V.nmask, V.base, V.len, V.rev, v0, n, 0);
}

// One indel scored the way pavlib/cigarcall.py:137-266 does it: the left shift (only when the previous op was '=',
// :149,225), then the breakpoint homology on both sides in the reference and in the contig. INS: SV sequence =
// contig[sq : sq+n] (re-sliced after the shift); DEL: reference[pr : pr+n] (never re-sliced, POS/END/SEQ stay unshifted).
// The five scans run as one rolled loop so the scan code exists once in the kernel (with all five call sites inlined the
// kernel is instruction-fetch bound). TILED: the sequences carry staged words (dev_homology_tiled), else plain global loads.
struct IndelScore {
    int32_t pos, end, qry_pos, qry_end;
    int32_t ls, hom_rl, hom_rr, hom_tl;
    int32_t hom_tr, seq_start;
};

template <bool TILED>
PAV_DEV void score_indel(const OSeq &R, const OSeq &Q, int32_t svtype, int32_t n, int32_t pr, int32_t pq, int32_t eqb, IndelScore &o)
{
    const int32_t L = (int32_t)Q.len;
    const bool ins = (svtype == 0);
    int ls = 0, hom_rl = 0, hom_rr = 0, hom_tl = 0, hom_tr = 0;
    int32_t sp = pr, sq = pq;
#pragma unroll 1
    for (int sc = (eqb > 0 ? 0 : 1); sc < 5; sc++) {
        const bool on_ref = sc <= 2;            // scans 0..2 walk the reference, 3..4 the contig
        const int left = (sc == 0 || sc == 1 || sc == 3);
        int64_t p;
        if (sc == 0) p = (int64_t)pr - 1;
        else if (sc == 1) p = (int64_t)sp - 1;
        else if (sc == 2) p = ins ? (int64_t)sp : (int64_t)sp + n;
        else if (sc == 3) p = (int64_t)sq - 1;
        else p = ins ? (int64_t)sq + n : (int64_t)sq;
        const OSeq &T = on_ref ? R : Q;
        const OSeq &V = ins ? Q : R;
        const int64_t v0 = ins ? (int64_t)sq : (int64_t)pr;   // sq == pq while sc == 0
        int h;
        if (TILED) h = dev_homology_tiled(T, p, V, v0, n, left);
        else h = dev_homology_raw(T.pack2, T.nmask, T.base, T.len, T.rev, p, V.pack2, V.nmask, V.base, V.len, V.rev, v0, n, left);
        if (sc == 0) { ls = min(eqb, h); sp = pr - ls; sq = pq - ls; }
        else if (sc == 1) hom_rl = h;
        else if (sc == 2) hom_rr = h;
        else if (sc == 3) hom_tl = h;
        else hom_tr = h;
    }
    if (ins) {          // cigarcall.py:157-173
        o.pos = sp; o.end = sp + 1;
        if (Q.rev) { o.qry_end = L - sq; o.qry_pos = o.qry_end - n; } else { o.qry_pos = sq; o.qry_end = sq + n; }
        o.seq_start = sq;
    } else {            // cigarcall.py:233-266 (POS/END/SEQ stay unshifted)
        o.pos = pr; o.end = pr + n;
        o.qry_pos = Q.rev ? L - sq : sq;
        o.qry_end = o.qry_pos + 1;
        o.seq_start = pr;
    }
    o.ls = ls; o.hom_rl = hom_rl; o.hom_rr = hom_rr; o.hom_tl = hom_tl; o.hom_tr = hom_tr;
}

// ---- the same scoring with a convergent first trip --------------------------------------------------
// 85-90 % of the scans of a real batch stop within their first 32 bases, and most SV sequences are shorter than 32 bases.
// score_indel runs every scan as a data-dependent loop of its own, so a warp executes each trip with the few lanes that still
// need it. Here the first 32 steps of every scan are ONE straight-line window compare for all lanes: the reference reads the
// SV sequence circularly (call.py:582 sv[-((h + 1) % n)], :637 sv[h % n]), so for n < 32 the 32-base pattern the flank is
// compared with is the periodic expansion of the SV's n bases, built with log2(32 / n) shift-ors. Only scans that match all 32
// bases continue, in common_extension: for n <= 32 a whole copy of the SV has matched by then, so the scan goes on as the
// flank compared with itself shifted by n (see dev_homology_raw), resumed at step 32.
PAV_DEV uint64_t shl64(uint64_t x, int s) { return s < 64 ? x << s : 0ull; }
PAV_DEV uint64_t shr64(uint64_t x, int s) { return s < 64 ? x >> s : 0ull; }

// left != 0: bases 32-n..31 of the window (the SV's n bases, last base least significant) repeated towards base 0;
// left == 0: bases 0..n-1 repeated towards base 31. The mask word alike. n >= 32: unchanged.
PAV_DEV void periodic32(uint64_t &b, uint32_t &m, int n, int left)
{
    if (n >= 32) return;
    uint64_t x;
    uint32_t y;
    // the period doubles with every step (n, 2n, 4n, ...): five steps cover n = 1; shifts past the word are zero, so all lanes run
    // the same five steps whatever their n
    if (left) {
        x = b & (shl64(1ull, 2 * n) - 1ull);
        y = (m >> (32 - n)) << (32 - n);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const int s = n << j;
            x |= shl64(x, 2 * s);
            y |= s < 32 ? y >> s : 0u;
        }
    } else {
        x = (b >> (64 - 2 * n)) << (64 - 2 * n);
        y = m & ((1u << n) - 1u);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const int s = n << j;
            x |= shr64(x, 2 * s);
            y |= s < 32 ? y << s : 0u;
        }
    }
    b = x; m = y;
}

// First 32 steps of a scan: the flank window at p (left != 0: bases p-31..p, else p..p+31) against the pattern. Returns 0..32.
template <bool TILED>
PAV_DEV int first_trip(const OSeq &T, int32_t p, uint64_t pat, uint32_t patm, int left)
{
    if (p < 0 || p >= (int32_t)T.len) return 0;
    uint64_t wt; uint32_t mt;
    oseq_window<TILED>(T, left ? p - 31 : p, wt, mt);
    const uint64_t x = wt ^ pat;
    const uint64_t d = (x | (x >> 1)) & 0x5555555555555555ull;
    const uint32_t m = mt | patm;
    int stop_d, stop_m;
    if (left) {
        stop_d = d ? ((__ffsll((long long)d) - 1) >> 1) : 32;
        stop_m = m ? __clz((int)m) : 32;
    } else {
        stop_d = d ? (__clzll((long long)d) >> 1) : 32;
        stop_m = m ? (__ffs((int)m) - 1) : 32;
    }
    return min(stop_d, stop_m);
}

// The rest of a scan whose first 32 steps all matched. V = SV sequence at v0, n bases. One common_extension site serves both
// stages: (n > 32 only) the flank against the rest of the SV, then the flank against itself shifted by n.
template <bool TILED>
PAV_DEV int scan_rest(const OSeq &T, int32_t p, const OSeq &V, int32_t v0, int n, int left)
{
    int32_t h = 32;
#pragma unroll 1
    for (int stage = (n > 32 ? 0 : 1); stage < 2; stage++) {
        const OSeq B = stage == 0 ? V : T;      // (a copy: field-wise selects; a reference would force the structs into local memory)
        const int32_t a = left ? p - h : p + h;
        const int32_t b = stage == 0 ? (left ? v0 + n - 1 - 32 : v0 + 32) : (left ? a + n : a - n);
        const int32_t lim = stage == 0 ? n - 32 : 0x7fffffff - 64 - n;
        const int32_t e = (TILED || !HOM_BATCH_LOADS) ? common_extension<TILED>(T, a, B, b, lim, left) : common_extension2(T, a, B, b, lim, left);
        h += e;
        if (stage == 0 && e < lim) break;
    }
    return h;
}

// ---- cooperative rests -----------------------------------------------------------------------------
// A scan that goes on after its first 32 bases is a chain of dependent window reads when one thread follows it (tandem repeats:
// hundreds of bases). Here G consecutive lanes share one scan: lane g compares window g, g + G, ... of it, all G windows of a
// round are in flight together, and the first lane (in scan order) whose window stops decides. Whole warps call these functions
// together (every lane the same number of times); a group with nothing to do passes limit 0. Device only (warp votes).
#ifdef __CUDACC__
template <int G>
__device__ __forceinline__ int32_t coop_extension(const OSeq &A, int32_t a, const OSeq &B, int32_t b, int32_t limit, int left, int g)
{
    static_assert(G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group size");
    const unsigned FULLW = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gbase = lane & ~(G - 1);
    const unsigned gmask = (G == 32) ? FULLW : (((1u << G) - 1u) << gbase);
    const int32_t a0 = left ? a - 31 : a, b0 = left ? b - 31 : b, step = left ? -32 : 32;
    int32_t h = 0;          // bases matched so far (multiple of 32 while the group is still scanning)
    bool done = limit <= 0;
    int32_t result = 0;
    while (__any_sync(FULLW, !done)) {
        int stop = 32;
        if (!done) {
            const int32_t t = h / 32 + g;                 // my window of this round
            if ((int64_t)t * 32 < (int64_t)limit) {
                uint64_t wa, wb; uint32_t ma, mb;
                oseq_window<false>(A, a0 + step * t, wa, ma);
                oseq_window<false>(B, b0 + step * t, wb, mb);
                stop = window_stop(wa, ma, wb, mb, left);
            } else stop = 0;                              // past the cap: the scan ends here at the latest
        }
        const unsigned bal = __ballot_sync(FULLW, !done && stop < 32) & gmask;
        const int f = bal ? __ffs((int)bal) - 1 : lane;   // first stopping lane of my group
        const int sf = __shfl_sync(FULLW, stop, f);       // (every lane of the warp executes the vote and the shuffle)
        if (!done) {
            if (bal) {
                result = h + 32 * (f - gbase) + sf;
                done = true;
            } else {
                h += 32 * G;
                if (h < 0 || h >= limit) { result = limit; done = true; }
            }
        }
    }
    return result < limit ? result : limit;
}

// scan_rest by a group of G lanes (all lanes of the group pass the same arguments; active == false: nothing to do, returns 32).
template <int G>
__device__ __forceinline__ int scan_rest_coop(const OSeq &T, int32_t p, const OSeq &V, int32_t v0, int n, int left, bool active, int g)
{
    const int32_t lim0 = (active && n > 32) ? n - 32 : 0;
    const int32_t e0 = coop_extension<G>(T, left ? p - 32 : p + 32, V, left ? v0 + n - 1 - 32 : v0 + 32, lim0, left, g);
    int32_t h = 32 + e0;
    const bool cont = active && e0 >= lim0;              // n <= 32: straight to the self-comparison; n > 32: only after a whole copy
    const int32_t a = left ? p - h : p + h;
    const int32_t b = left ? a + n : a - n;
    const int32_t e1 = coop_extension<G>(T, a, T, b, cont ? 0x7fffffff - 64 - n : 0, left, g);
    return h + e1;
}

template <int G>
__device__ __forceinline__ int indel_rest_coop(const OSeq &R, const OSeq &Q, bool ins, int32_t n, int32_t pr, int32_t pq, int32_t ls, int sc, bool active, int g);
#endif

// The pieces of one indel's scoring, in the order the reference runs them (cigarcall.py:137-266). score_indel2 strings them
// together for one thread; homology_queue_kernel runs the same pieces with the rests of a whole CTA's indels pooled in between.
//   sc 0 = left-shift scan (ref, leftwards from pr - 1, SV sequence at the unshifted position); sc 1..4 = hom_ref_l, hom_ref_r,
//   hom_tig_l, hom_tig_r at the shifted position.
PAV_DEV void scan_geometry(int sc, bool ins, int32_t n, int32_t pr, int32_t pq, int32_t ls, int32_t &p, int32_t &v0, bool &on_ref, int &left)
{
    const int32_t sp = pr - ls, sq = pq - ls;
    on_ref = sc <= 2;
    left = (sc == 0 || sc == 1 || sc == 3);
    if (sc == 0) { p = pr - 1; v0 = ins ? pq : pr; return; }
    v0 = ins ? sq : pr;                      // INS: the SV sequence is re-sliced at the shifted position; DEL: never (:233-266)
    if (sc == 1) p = sp - 1;
    else if (sc == 2) p = ins ? sp : sp + n;
    else if (sc == 3) p = sq - 1;
    else p = ins ? sq + n : sq;
}

// First 32 steps of a scan from finished windows: flank window (wt, mt) of T at p against the pattern.
PAV_DEV int first_trip_w(const OSeq &T, int32_t p, uint64_t wt, uint32_t mt, uint64_t pat, uint32_t patm, int left)
{
    const int stop = window_stop(wt, mt, pat, patm, left);
    return (p < 0 || p >= (int32_t)T.len) ? 0 : stop;
}

// First trip of the left-shift scan (cigarcall.py:149-155 / :225-231). 0..32; 32 = all matched, the rest is scan 0.
template <bool TILED>
PAV_DEV int indel_phase0(const OSeq &R, const OSeq &Q, bool ins, int32_t n, int32_t pr, int32_t pq)
{
    if (!TILED && HOM_BATCH_LOADS) {   // both windows requested before either is used
        const OSeq V = ins ? Q : R;
        const WinReq qv = win_prep(V, (ins ? pq : pr) + n - 32), qt = win_prep(R, pr - 32);
        const WinRaw rv = win_load(V, qv), rt = win_load(R, qt);
        uint64_t patl, wt; uint32_t patlm, mt;
        win_finish(qv, rv, patl, patlm);
        win_finish(qt, rt, wt, mt);
        periodic32(patl, patlm, n, 1);
        return first_trip_w(R, pr - 1, wt, mt, patl, patlm, 1);
    }
    const OSeq &V = ins ? Q : R;
    uint64_t patl; uint32_t patlm;
    oseq_window<TILED>(V, (ins ? pq : pr) + n - 32, patl, patlm);
    periodic32(patl, patlm, n, 1);
    return first_trip<TILED>(R, pr - 1, patl, patlm, 1);
}

// First trips of the four breakpoint homologies at the shifted position (:178-182 / :247-251): hom[0..3] = ref left, ref right,
// contig left, contig right (32 = all matched, rest pending).
template <bool TILED>
PAV_DEV void indel_phase1(const OSeq &R, const OSeq &Q, bool ins, int32_t n, int32_t pr, int32_t pq, int32_t ls, int (&hom)[4])
{
    const int32_t sp = pr - ls, sq = pq - ls, v0 = ins ? sq : pr;
    uint64_t patl, patr; uint32_t patlm, patrm;
    if (!TILED && HOM_BATCH_LOADS) {   // all six windows requested before any is used
        const OSeq V = ins ? Q : R;
        const int32_t p_rl = sp - 1, p_rr = ins ? sp : sp + n, p_tl = sq - 1, p_tr = ins ? sq + n : sq;
        const WinReq qvl = win_prep(V, v0 + n - 32), qvr = win_prep(V, v0);
        const WinReq q0 = win_prep(R, p_rl - 31), q1 = win_prep(R, p_rr), q2 = win_prep(Q, p_tl - 31), q3 = win_prep(Q, p_tr);
        const WinRaw rvl = win_load(V, qvl), rvr = win_load(V, qvr);
        const WinRaw r0 = win_load(R, q0), r1 = win_load(R, q1), r2 = win_load(Q, q2), r3 = win_load(Q, q3);
        win_finish(qvl, rvl, patl, patlm);
        win_finish(qvr, rvr, patr, patrm);
        periodic32(patl, patlm, n, 1);
        periodic32(patr, patrm, n, 0);
        uint64_t w; uint32_t m;
        win_finish(q0, r0, w, m); hom[0] = first_trip_w(R, p_rl, w, m, patl, patlm, 1);
        win_finish(q1, r1, w, m); hom[1] = first_trip_w(R, p_rr, w, m, patr, patrm, 0);
        win_finish(q2, r2, w, m); hom[2] = first_trip_w(Q, p_tl, w, m, patl, patlm, 1);
        win_finish(q3, r3, w, m); hom[3] = first_trip_w(Q, p_tr, w, m, patr, patrm, 0);
        return;
    }
    const OSeq &V = ins ? Q : R;
    oseq_window<TILED>(V, v0 + n - 32, patl, patlm);
    periodic32(patl, patlm, n, 1);
    oseq_window<TILED>(V, v0, patr, patrm);
    periodic32(patr, patrm, n, 0);
    hom[0] = first_trip<TILED>(R, sp - 1, patl, patlm, 1);
    hom[1] = first_trip<TILED>(R, ins ? sp : sp + n, patr, patrm, 0);
    hom[2] = first_trip<TILED>(Q, sq - 1, patl, patlm, 1);
    hom[3] = first_trip<TILED>(Q, ins ? sq + n : sq, patr, patrm, 0);
}

// The rest of scan sc (its first 32 steps all matched).
template <bool TILED>
PAV_DEV int indel_rest(const OSeq &R, const OSeq &Q, bool ins, int32_t n, int32_t pr, int32_t pq, int32_t ls, int sc)
{
    int32_t p, v0; bool on_ref; int left;
    scan_geometry(sc, ins, n, pr, pq, ls, p, v0, on_ref, left);
    const OSeq T = on_ref ? R : Q;
    const OSeq V = ins ? Q : R;
    return scan_rest<TILED>(T, p, V, v0, n, left);
}

#ifdef __CUDACC__
template <int G>
__device__ __forceinline__ int indel_rest_coop(const OSeq &R, const OSeq &Q, bool ins, int32_t n, int32_t pr, int32_t pq, int32_t ls, int sc, bool active, int g)
{
    int32_t p, v0; bool on_ref; int left;
    scan_geometry(sc, ins, n, pr, pq, ls, p, v0, on_ref, left);
    const OSeq T = on_ref ? R : Q;
    const OSeq V = ins ? Q : R;
    return scan_rest_coop<G>(T, p, V, v0, n, left, active, g);
}
#endif

PAV_DEV void indel_finish(const OSeq &Q, bool ins, int32_t n, int32_t pr, int32_t pq, int32_t ls, const int (&hom)[4], IndelScore &o)
{
    const int32_t L = (int32_t)Q.len, sp = pr - ls, sq = pq - ls;
    if (ins) {          // cigarcall.py:157-173
        o.pos = sp; o.end = sp + 1;
        if (Q.rev) { o.qry_end = L - sq; o.qry_pos = o.qry_end - n; } else { o.qry_pos = sq; o.qry_end = sq + n; }
        o.seq_start = sq;
    } else {            // cigarcall.py:233-266 (POS/END/SEQ stay unshifted)
        o.pos = pr; o.end = pr + n;
        o.qry_pos = Q.rev ? L - sq : sq;
        o.qry_end = o.qry_pos + 1;
        o.seq_start = pr;
    }
    o.ls = ls; o.hom_rl = hom[0]; o.hom_rr = hom[1]; o.hom_tl = hom[2]; o.hom_tr = hom[3];
}

template <bool TILED>
PAV_DEV void score_indel2(const OSeq &R, const OSeq &Q, int32_t svtype, int32_t n, int32_t pr, int32_t pq, int32_t eqb, IndelScore &o)
{
    const bool ins = (svtype == 0);
    int h0 = indel_phase0<TILED>(R, Q, ins, n, pr, pq);
    if (eqb <= 0) h0 = 0;                       // the shift counts only when the previous op was '='
    int ls = 0;
    int hom[4] = {0, 0, 0, 0};
    // sc 0 = rest of the left-shift scan (only when it can still grow the shift); entering sc 1 everything moves to the shifted
    // position and the four homologies take their first trips; sc 1..4 = their rests. One rolled loop: one copy of the rest code.
#pragma unroll 1
    for (int sc = (h0 == 32 && eqb > 32) ? 0 : 1; sc < 5; sc++) {
        if (sc == 1) {
            ls = min(eqb, h0);
            indel_phase1<TILED>(R, Q, ins, n, pr, pq, ls, hom);
        }
        const int hcur = sc == 0 ? 32 : sc == 1 ? hom[0] : sc == 2 ? hom[1] : sc == 3 ? hom[2] : hom[3];
        if (hcur != 32) continue;
        const int h = indel_rest<TILED>(R, Q, ins, n, pr, pq, ls, sc);
        if (sc == 0) h0 = h;
        else if (sc == 1) hom[0] = h;
        else if (sc == 2) hom[1] = h;
        else if (sc == 3) hom[2] = h;
        else hom[3] = h;
    }
    indel_finish(Q, ins, n, pr, pq, ls, hom, o);
}

// ---- staged words: which part of a plane the opt-in homology kernels copy on chip ---------------
#ifndef HOM_TILE_WORDS_N
#define HOM_TILE_WORDS_N 768
#endif
constexpr int HOM_TILE_WORDS = HOM_TILE_WORDS_N;   // homology_tiled_kernel: 32-base words per sequence and warp (768 words = 24.5 kbp)
constexpr int HOM_TILE_MARGIN = 160;               //   bases staged beyond the outermost breakpoints
constexpr int NBR_WORDS = 8;                       // homology_nbr_kernel: words per sequence and indel (256 bases)

// Word range [w0, w0 + nw) of a plane covering breakpoints lo_g..hi_g (plane base coordinates) plus the margin, 4-word aligned
// and clamped to the plane; nw = 0 when it does not fit the tile.
PAV_DEV void tile_range(long long lo_g, long long hi_g, int64_t plane_words, int64_t &w0, int32_t &nw)
{
    long long a = (lo_g - HOM_TILE_MARGIN) >> 5, b = ((hi_g + HOM_TILE_MARGIN) >> 5) + 2;
    a = max(a, 0ll) & ~3ll;
    b = min((b + 3) & ~3ll, (long long)plane_words);
    w0 = a;
    nw = (b > a && b - a <= HOM_TILE_WORDS) ? (int32_t)(b - a) : 0;
}

// First word of the 8-word neighbourhood around plane coordinate c (sector-aligned, inside the plane); -1 if the plane is too small.
PAV_DEV int64_t nbr_first_word(long long c, int64_t plane_words)
{
    if (plane_words < NBR_WORDS) return -1;
    long long w = ((c - 64) >> 5) & ~3ll;
    w = max(w, 0ll);
    return (int64_t)min(w, (long long)plane_words - NBR_WORDS);
}

// ---- k-mers ---------------------------------------------------------------------------------------
PAV_DEV uint64_t hash_kmer(uint64_t k, int log2cap)
{
    return (k * 0x9E3779B97F4A7C15ull) >> (64 - log2cap);
}

// kmer.py:118-133 as bit tricks: complement, then reverse the 2-bit groups of the 64-bit word.
PAV_DEV uint64_t kmer_revcomp(uint64_t kmer, int k)
{
    uint64_t x = ~kmer;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | (uint64_t)__byte_perm((uint32_t)(x >> 32), 0, 0x0123);
    return x >> (64 - 2 * k);
}

// k-mer starting at global base g (first base most significant); returns false if any of its k bases
// is not ACGT (stream() would not have emitted it, kmer.py:206-221).
PAV_DEV bool kmer_at(const uint64_t *__restrict__ pack2, const uint32_t *__restrict__ nmask, int64_t g, int k,
                                        uint64_t &kmer, const uint32_t *__restrict__ nsum = nullptr)
{
    int64_t w = g >> 5;
    int s = (int)(g & 31);
    if (nsum_any2(nsum, w)) {
        uint64_t m = (uint64_t)__ldg(nmask + w) | ((uint64_t)__ldg(nmask + w + 1) << 32);
        m >>= s;
        uint64_t kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1);
        if (m & kmask) return false;
    }
    uint64_t hi = __ldg(pack2 + w), lo = __ldg(pack2 + w + 1);
    uint64_t x = s ? ((hi << (2 * s)) | (lo >> (64 - 2 * s))) : hi;
    kmer = x >> (64 - 2 * k);
    return true;
}
