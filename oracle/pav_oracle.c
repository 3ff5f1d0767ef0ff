/*
 * pav_oracle.c -- TEST INFRASTRUCTURE ONLY. CPU restatement (plain C, scalar, ASCII strings) of the
 * reference algorithms on PAV's variant-calling hot path. It is the checker for the CUDA path and
 * the "port" CPU baseline of bench.py; it is never linked, imported or called by the product
 * (pav_b200/), which has no CPU fallback.
 *
 * Parity status: PINNED against golden vectors produced by running the unmodified reference
 * (PAV 2.4.6.0) in the build container -- tests/golden/ (generator: tests/golden/make_golden.py),
 * checked by tests/test_oracle_golden.py. The reference itself ships no tests or golden vectors.
 *
 * Reference files restated here (paths relative to /root/reference):
 *   pavlib/align/align.py:286-322            cigar_str_to_tuples       -> next_cigar_op()
 *   pavlib/call.py:542-592 / 595-647         left_/right_homology      -> orc_left_homology()/orc_right_homology()
 *   pavlib/cigarcall.py:50-311               per-record CIGAR walk     -> orc_walk_record()
 *   dep/svpop/dep/kanapy/util/kmer.py:61-69,118-133,186-221  k-mer append / rev_complement / stream
 *                                                                      -> kmer_stream(), kmer_rc()
 *   pavlib/seq.py:305-325                    ref_kmers (Counter)       -> sorted multiset in orc_density()
 *   scripts/density.py:154-342,508-545       get_smoothed_density + __main__ checks -> orc_density()
 *   scipy.stats.gaussian_kde (third party, scipy 1.18.1 in this image; version unpinned by the
 *   reference's Dockerfile:57-70): covariance = var(ddof=1) * factor^2, estimate[j] = sum_i w_i *
 *   exp(-((x_i - x_j)/L)^2 / 2) * (2 pi)^-1/2 / L, accumulated in data order        -> kde_eval()
 *   numpy.interp (linear, slope * (x - x0) + y0)                                     -> orc_density()
 *
 * The data structures are deliberately different from the CUDA path (ASCII bytes instead of 2-bit
 * planes, qsort + bsearch instead of a hash table, sequential sums instead of parallel reductions)
 * so that agreement is evidence, not tautology.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_ILLEGAL_OP 1    /* cigarcall.py:289-307 */
#define ORC_ERR_MISSING_LEN 2   /* align.py:310-313 */
#define ORC_ERR_UNKNOWN_OP 3    /* align.py:315-318 */
#define ORC_ERR_INDEX 4         /* align.py:308 running off the string (IndexError) */
#define ORC_ERR_NOMEM 9
#define ORC_INV_FAIL 125        /* pavlib/constants.py:55 */

/* ------------------------------------------------------------------ homology ------------------ */

static int is_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

/* call.py:542-592. seq / sv are upper-case. Circular read of sv from its end: sv[-((h+1) % n)],
 * where index -0 is index 0. */
int64_t orc_left_homology(int64_t pos, const char *seq, int64_t seq_len, const char *sv, int64_t svlen)
{
    (void)seq_len;
    if (seq == NULL || sv == NULL) return 0;
    int64_t h = 0;
    while (h <= pos) {
        char b = seq[pos - h];
        if (!is_acgt(b)) break;
        int64_t m = (h + 1) % svlen;
        char s = (m == 0) ? sv[0] : sv[svlen - m];
        if (s != b) break;
        h++;
    }
    return h;
}

/* call.py:595-647 */
int64_t orc_right_homology(int64_t pos, const char *seq, int64_t seq_len, const char *sv, int64_t svlen)
{
    if (seq == NULL || sv == NULL) return 0;
    int64_t h = 0, limit = seq_len - pos;
    while (h < limit) {
        char b = seq[pos + h];
        if (!is_acgt(b)) break;
        if (sv[h % svlen] != b) break;
        h++;
    }
    return h;
}

/* ------------------------------------------------------------------ Path A --------------------- */

typedef struct {
    int64_t pos_ref;   /* POS */
    int64_t qry_pos;   /* 0-based position on the forward contig (QRY_REGION = qry_pos+1) */
    int32_t rec;       /* row number in df_align */
    uint8_t ref_base;  /* original case */
    uint8_t alt_base;  /* original case (contig already reverse-complemented when REV) */
    uint8_t pad[2];
} orc_snv_t;

typedef struct {
    int64_t pos, end, svlen;
    int64_t qry_pos, qry_end;       /* 0-based half-open on the forward contig: QRY_REGION = qry_pos+1 .. qry_end */
    int64_t left_shift;
    int64_t hom_ref_l, hom_ref_r, hom_tig_l, hom_tig_r;
    int64_t seq_start;              /* INS: offset of SEQ in the oriented contig; DEL: offset in the reference */
    int32_t rec;
    int32_t svtype;                 /* 0 INS, 1 DEL */
} orc_indel_t;

typedef struct {
    orc_snv_t *snv; int64_t n_snv, cap_snv;
    orc_indel_t *indel; int64_t n_indel, cap_indel;
    /* error detail */
    int64_t err_cigar_index; int err_op; int64_t err_pos_ref, err_pos_tig, err_text_pos; int err_char;
} orc_walk_t;

orc_walk_t *orc_walk_new(void) { return (orc_walk_t *)calloc(1, sizeof(orc_walk_t)); }
void orc_walk_free(orc_walk_t *w) { if (w) { free(w->snv); free(w->indel); free(w); } }
int64_t orc_walk_n_snv(const orc_walk_t *w) { return w->n_snv; }
int64_t orc_walk_n_indel(const orc_walk_t *w) { return w->n_indel; }
const orc_snv_t *orc_walk_snv(const orc_walk_t *w) { return w->snv; }
const orc_indel_t *orc_walk_indel(const orc_walk_t *w) { return w->indel; }
void orc_walk_error(const orc_walk_t *w, int64_t *cigar_index, int *op, int64_t *pos_ref, int64_t *pos_tig,
                    int64_t *text_pos, int *ch)
{
    *cigar_index = w->err_cigar_index; *op = w->err_op; *pos_ref = w->err_pos_ref; *pos_tig = w->err_pos_tig;
    *text_pos = w->err_text_pos; *ch = w->err_char;
}

static int push_snv(orc_walk_t *w, orc_snv_t r)
{
    if (w->n_snv == w->cap_snv) {
        int64_t c = w->cap_snv ? w->cap_snv * 2 : 1024;
        orc_snv_t *p = (orc_snv_t *)realloc(w->snv, (size_t)c * sizeof(orc_snv_t));
        if (!p) return ORC_ERR_NOMEM;
        w->snv = p; w->cap_snv = c;
    }
    w->snv[w->n_snv++] = r;
    return 0;
}

static int push_indel(orc_walk_t *w, orc_indel_t r)
{
    if (w->n_indel == w->cap_indel) {
        int64_t c = w->cap_indel ? w->cap_indel * 2 : 256;
        orc_indel_t *p = (orc_indel_t *)realloc(w->indel, (size_t)c * sizeof(orc_indel_t));
        if (!p) return ORC_ERR_NOMEM;
        w->indel = p; w->cap_indel = c;
    }
    w->indel[w->n_indel++] = r;
    return 0;
}

/* align.py:286-322 -- one token; returns 0 ok, or an ORC_ERR_* code. */
static int next_cigar_op(const char *cigar, int64_t n, int64_t *pos, int64_t *oplen, char *op,
                         int64_t *err_text_pos, int *err_char)
{
    int64_t p = *pos, q = p;
    while (q < n && cigar[q] >= '0' && cigar[q] <= '9') q++;
    if (q >= n) { *err_text_pos = q; return ORC_ERR_INDEX; }           /* cigar[len_pos] past the end */
    if (q == p) { *err_text_pos = p; return ORC_ERR_MISSING_LEN; }
    if (strchr("MIDNSHP=X", cigar[q]) == NULL) { *err_text_pos = p; *err_char = (unsigned char)cigar[p]; return ORC_ERR_UNKNOWN_OP; }
    int64_t v = 0;
    for (int64_t i = p; i < q; i++) v = v * 10 + (cigar[i] - '0');
    *oplen = v; *op = cigar[q]; *pos = q + 1;
    return 0;
}

/*
 * cigarcall.py:50-311 for one alignment record.
 *   ref / ref_up : whole chromosome, original case / upper-cased
 *   qry / qry_up : whole contig ALREADY in reference orientation (reverse-complemented iff is_rev)
 */
int orc_walk_record(orc_walk_t *w, const char *cigar, int64_t cigar_len, int64_t pos_ref0,
                    const char *ref, const char *ref_up, int64_t ref_len,
                    const char *qry, const char *qry_up, int64_t qry_len,
                    int is_rev, int32_t rec)
{
    int64_t pos_ref = pos_ref0, pos_tig = 0, tpos = 0, cigar_index = 0;
    char last_op = 0; int64_t last_oplen = 0;
    (void)ref;

    while (tpos < cigar_len) {
        int64_t oplen = 0; char op = 0; int ech = 0; int64_t etp = 0;
        int rc = next_cigar_op(cigar, cigar_len, &tpos, &oplen, &op, &etp, &ech);
        if (rc) { w->err_text_pos = etp; w->err_char = ech; return rc; }
        cigar_index++;

        if (op == '=') {
            pos_ref += oplen; pos_tig += oplen;
        } else if (op == 'X') {
            for (int64_t i = 0; i < oplen; i++) {
                orc_snv_t r; memset(&r, 0, sizeof r);
                int64_t pr = pos_ref + i, pt = pos_tig + i;
                r.pos_ref = pr;
                r.ref_base = (uint8_t)ref[pr];
                r.alt_base = (uint8_t)qry[pt];
                r.qry_pos = is_rev ? (qry_len - pt - 1) : pt;
                r.rec = rec;
                if (push_snv(w, r)) return ORC_ERR_NOMEM;
            }
            pos_ref += oplen; pos_tig += oplen;
        } else if (op == 'I') {
            const char *sv = qry_up + pos_tig;          /* seq_upper */
            int64_t ls = 0;
            if (last_op == '=') {
                ls = orc_left_homology(pos_ref - 1, ref_up, ref_len, sv, oplen);
                if (last_oplen < ls) ls = last_oplen;
            }
            int64_t sv_pos_ref = pos_ref - ls, sv_pos_tig = pos_tig - ls, sv_end_tig = sv_pos_tig + oplen;
            sv = qry_up + sv_pos_tig;                   /* re-sliced after the shift */
            orc_indel_t r; memset(&r, 0, sizeof r);
            r.svtype = 0; r.rec = rec; r.svlen = oplen; r.left_shift = ls;
            r.pos = sv_pos_ref; r.end = sv_pos_ref + 1;
            if (is_rev) { r.qry_end = qry_len - sv_pos_tig; r.qry_pos = r.qry_end - oplen; }
            else { r.qry_pos = sv_pos_tig; r.qry_end = sv_pos_tig + oplen; }
            r.hom_ref_l = orc_left_homology(sv_pos_ref - 1, ref_up, ref_len, sv, oplen);
            r.hom_ref_r = orc_right_homology(sv_pos_ref, ref_up, ref_len, sv, oplen);
            r.hom_tig_l = orc_left_homology(sv_pos_tig - 1, qry_up, qry_len, sv, oplen);
            r.hom_tig_r = orc_right_homology(sv_end_tig, qry_up, qry_len, sv, oplen);
            r.seq_start = sv_pos_tig;
            if (push_indel(w, r)) return ORC_ERR_NOMEM;
            pos_tig += oplen;
        } else if (op == 'D') {
            const char *sv = ref_up + pos_ref;          /* NOT re-sliced after the shift (reference quirk) */
            int64_t ls = 0;
            if (last_op == '=') {
                ls = orc_left_homology(pos_ref - 1, ref_up, ref_len, sv, oplen);
                if (last_oplen < ls) ls = last_oplen;
            }
            int64_t sv_pos_ref = pos_ref - ls, sv_end_ref = sv_pos_ref + oplen, sv_pos_tig = pos_tig - ls;
            orc_indel_t r; memset(&r, 0, sizeof r);
            r.svtype = 1; r.rec = rec; r.svlen = oplen; r.left_shift = ls;
            r.pos = pos_ref; r.end = pos_ref + oplen;   /* unshifted (cigarcall.py:254-258) */
            r.qry_pos = is_rev ? (qry_len - sv_pos_tig) : sv_pos_tig;
            r.qry_end = r.qry_pos + 1;                  /* QRY_REGION = qry_pos+1 .. qry_pos+1 */
            r.hom_ref_l = orc_left_homology(sv_pos_ref - 1, ref_up, ref_len, sv, oplen);
            r.hom_ref_r = orc_right_homology(sv_end_ref, ref_up, ref_len, sv, oplen);
            r.hom_tig_l = orc_left_homology(sv_pos_tig - 1, qry_up, qry_len, sv, oplen);
            r.hom_tig_r = orc_right_homology(sv_pos_tig, qry_up, qry_len, sv, oplen);
            r.seq_start = pos_ref;
            if (push_indel(w, r)) return ORC_ERR_NOMEM;
            pos_ref += oplen;
        } else if (op == 'S' || op == 'H') {
            pos_tig += oplen;
        } else {
            w->err_cigar_index = cigar_index; w->err_op = op; w->err_pos_ref = pos_ref; w->err_pos_tig = pos_tig;
            return ORC_ERR_ILLEGAL_OP;
        }
        last_op = op; last_oplen = oplen;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ Path B --------------------- */

static int base_code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

/* kmer.py:118-133 */
uint64_t orc_kmer_rc(uint64_t kmer, int k)
{
    uint64_t rev = 0;
    for (int i = 0; i < k; i++) { rev = (rev << 2) | ((kmer & 3u) ^ 3u); kmer >>= 2; }
    return rev;
}

/* kmer.py:186-221: returns number of k-mers written (kmers/index may be NULL to count only). */
int64_t orc_kmer_stream(const char *seq, int64_t n, int k, uint64_t *kmers, int32_t *index)
{
    uint64_t mask = (k >= 32) ? ~(uint64_t)0 : (((uint64_t)1 << (2 * k)) - 1);
    uint64_t kmer = 0; int load = 1; int64_t kmer_index = -(int64_t)k, out = 0;
    for (int64_t i = 0; i < n; i++) {
        kmer_index++;
        int c = base_code(seq[i]);
        if (c >= 0) {
            kmer = ((kmer << 2) | (uint64_t)c) & mask;
            if (load == k) {
                if (kmers) { kmers[out] = kmer; index[out] = (int32_t)kmer_index; }
                out++;
            } else load++;
        } else load = 1;
    }
    return out;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

static int in_set(const uint64_t *set, int64_t n, uint64_t key)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (set[mid] < key) lo = mid + 1; else hi = mid; }
    return lo < n && set[lo] == key;
}

typedef struct {
    int64_t n_rows; int smoothed;
    uint64_t *kmer; int32_t *index; int8_t *state_mer; int8_t *state;
    double *kern[3];
    int64_t n_eval;          /* number of full KDE evaluations (lattice points), for flop accounting */
    int64_t max_ref_count;
} orc_density_t;

void orc_density_free(orc_density_t *d)
{
    if (!d) return;
    free(d->kmer); free(d->index); free(d->state_mer); free(d->state);
    for (int s = 0; s < 3; s++) free(d->kern[s]);
    free(d);
}
int64_t orc_density_rows(const orc_density_t *d) { return d->n_rows; }
int orc_density_smoothed(const orc_density_t *d) { return d->smoothed; }
int64_t orc_density_n_eval(const orc_density_t *d) { return d->n_eval; }
const uint64_t *orc_density_kmer(const orc_density_t *d) { return d->kmer; }
const int32_t *orc_density_index(const orc_density_t *d) { return d->index; }
const int8_t *orc_density_state_mer(const orc_density_t *d) { return d->state_mer; }
const int8_t *orc_density_state(const orc_density_t *d) { return d->state; }
const double *orc_density_kern(const orc_density_t *d, int s) { return d->kern[s]; }

typedef struct { int64_t n; double *x_scaled; double norm; double w; } kde_t;

/* scipy gaussian_kde.evaluate for one state at lattice point j, times n_s (density.py:106-115). */
static double kde_eval(const kde_t *kd, double L, int64_t j)
{
    if (kd->n == 0) return 0.0;
    double xj = (double)j / L, est = 0.0;
    for (int64_t i = 0; i < kd->n; i++) {
        double r = kd->x_scaled[i] - xj;
        double arg = r * r;
        est += kd->w * (exp(-arg / 2.0) * kd->norm);
    }
    return est * (double)kd->n;
}

static int argmax3(double a, double b, double c)
{
    int m = 0; double v = a;
    if (b > v) { m = 1; v = b; }
    if (c > v) { m = 2; }
    return m;
}

/*
 * scripts/density.py __main__ + get_smoothed_density().
 * ref_seq: reference window, forward strand (seq.py:316). tig_seq: contig window, forward contig
 * coordinates (never reverse-complemented: density.py:502,543). rev: "-r true" (density.py:538-539).
 * Returns 0 with *out set, or 125 (soft failure), or 9 (no memory).
 */
int orc_density(const char *ref_seq, int64_t ref_len, const char *tig_seq, int64_t tig_len, int k, int rev,
                int min_inf, double smooth, int min_state, int srs, double delta, orc_density_t **out)
{
    *out = NULL;
    /* reference k-mer multiset (seq.py:305-325) */
    int64_t nr = orc_kmer_stream(ref_seq, ref_len, k, NULL, NULL);
    if (nr == 0) return ORC_INV_FAIL;                                   /* density.py:510-513 */
    uint64_t *rk = (uint64_t *)malloc((size_t)nr * 8); int32_t *ri = (int32_t *)malloc((size_t)nr * 4);
    if (!rk || !ri) return ORC_ERR_NOMEM;
    orc_kmer_stream(ref_seq, ref_len, k, rk, ri);
    free(ri);
    qsort(rk, (size_t)nr, 8, cmp_u64);
    int64_t nu = 0, run = 0, max_run = 0;
    for (int64_t i = 0; i < nr; i++) {
        if (i > 0 && rk[i] == rk[i - 1]) run++; else { run = 1; rk[nu++] = rk[i]; }
        if (run > max_run) max_run = run;
    }
    if (max_run > 100) { free(rk); return ORC_INV_FAIL; }               /* density.py:516-527 */
    if (rev) {                                                          /* density.py:538-539 */
        for (int64_t i = 0; i < nu; i++) rk[i] = orc_kmer_rc(rk[i], k);
        qsort(rk, (size_t)nu, 8, cmp_u64);
    }

    /* contig k-mers + orientation state (density.py:543-545,170-175) */
    int64_t nt = orc_kmer_stream(tig_seq, tig_len, k, NULL, NULL);
    uint64_t *tk = (uint64_t *)malloc((size_t)(nt ? nt : 1) * 8); int32_t *ti = (int32_t *)malloc((size_t)(nt ? nt : 1) * 4);
    int8_t *sm = (int8_t *)malloc((size_t)(nt ? nt : 1));
    if (!tk || !ti || !sm) return ORC_ERR_NOMEM;
    orc_kmer_stream(tig_seq, tig_len, k, tk, ti);
    static const int8_t M[2][2] = {{-1, 2}, {0, 1}};
    int64_t cnt[3] = {0, 0, 0};
    for (int64_t i = 0; i < nt; i++) {
        int f = in_set(rk, nu, tk[i]), r = in_set(rk, nu, orc_kmer_rc(tk[i], k));
        sm[i] = M[f][r];
        if (sm[i] >= 0) cnt[sm[i]]++;
    }
    free(rk);
    /* drop -1, drop low-count states (density.py:178-190) */
    int keep_state[3];
    for (int s = 0; s < 3; s++) keep_state[s] = cnt[s] >= min_state;
    int64_t N = 0;
    for (int64_t i = 0; i < nt; i++)
        if (sm[i] >= 0 && keep_state[sm[i]]) { tk[N] = tk[i]; ti[N] = ti[i]; sm[N] = sm[i]; N++; }

    orc_density_t *d = (orc_density_t *)calloc(1, sizeof(orc_density_t));
    if (!d) return ORC_ERR_NOMEM;
    d->n_rows = N; d->kmer = tk; d->index = ti; d->state_mer = sm; d->max_ref_count = max_run;
    d->state = (int8_t *)malloc((size_t)(N ? N : 1));
    for (int64_t i = 0; i < N; i++) d->state[i] = -1;
    *out = d;
    if (N < min_inf) { d->smoothed = 0; return ORC_OK; }               /* density.py:193-194 */
    d->smoothed = 1;

    double bw = pow((double)N, -1.0 / 5.0) * smooth;                    /* density.py:198 */
    kde_t kd[3]; double L[3];
    for (int s = 0; s < 3; s++) {
        kd[s].n = 0; kd[s].x_scaled = NULL; L[s] = 1.0;
        d->kern[s] = (double *)malloc((size_t)N * 8);
        if (!d->kern[s]) return ORC_ERR_NOMEM;
    }
    for (int s = 0; s < 3; s++) {
        int64_t n = 0; double mean = 0.0;
        for (int64_t i = 0; i < N; i++) if (sm[i] == s) { n++; mean += (double)i; }
        kd[s].n = n;
        if (n == 0) continue;
        mean /= (double)n;
        double ss = 0.0;
        for (int64_t i = 0; i < N; i++) if (sm[i] == s) { double dv = (double)i - mean; ss += dv * dv; }
        double var = ss / (double)(n - 1);                              /* np.cov(bias=False) */
        L[s] = sqrt(var) * bw;                                          /* cho_cov = chol(cov) * factor */
        kd[s].norm = pow(2.0 * M_PI, -0.5) / L[s];
        kd[s].w = 1.0 / (double)n;
        kd[s].x_scaled = (double *)malloc((size_t)n * 8);
        int64_t q = 0;
        for (int64_t i = 0; i < N; i++) if (sm[i] == s) kd[s].x_scaled[q++] = (double)i / L[s];
    }

    uint8_t *have = (uint8_t *)calloc((size_t)N, 1);
    int64_t n_eval = 0;
    /* sampled lattice (density.py:211-247) */
    int64_t n_samp = 0; int64_t *samp = (int64_t *)malloc((size_t)(N / srs + 3) * 8);
    for (int64_t j = 0; j < N; j += srs) samp[n_samp++] = j;
    if (samp[n_samp - 1] != N - 1) samp[n_samp++] = N - 1;
    for (int64_t a = 0; a < n_samp; a++) {
        int64_t j = samp[a];
        for (int s = 0; s < 3; s++) d->kern[s][j] = kde_eval(&kd[s], L[s], j);
        d->state[j] = (int8_t)argmax3(d->kern[0][j], d->kern[1][j], d->kern[2][j]);
        have[j] = 1; n_eval++;
    }
    /* gaps: full KDE or interpolation (density.py:260-323) */
    for (int64_t a = 0; a + 1 < n_samp; a++) {
        int64_t lo = samp[a], hi = samp[a + 1];
        if (hi == lo + 1) continue;
        int change = d->state[lo] != d->state[hi];
        for (int64_t j = lo + 1; j <= hi && !change; j++) if (sm[j] != sm[lo]) change = 1;
        double dmax = 0.0;
        for (int s = 0; s < 3; s++) { double dv = fabs(d->kern[s][lo] - d->kern[s][hi]); if (dv > dmax) dmax = dv; }
        if (change || dmax > delta) {
            for (int64_t j = lo + 1; j < hi; j++) {
                for (int s = 0; s < 3; s++) d->kern[s][j] = kde_eval(&kd[s], L[s], j);
                n_eval++;
            }
        } else {
            for (int s = 0; s < 3; s++) {
                double slope = (d->kern[s][hi] - d->kern[s][lo]) / ((double)hi - (double)lo);
                for (int64_t j = lo + 1; j < hi; j++) d->kern[s][j] = slope * ((double)j - (double)lo) + d->kern[s][lo];
            }
        }
    }
    /* spikes and final state (density.py:330-338; pandas aligns the RHS frame on the column name) */
    for (int64_t j = 0; j < N; j++) {
        for (int s = 0; s < 3; s++) if (d->kern[s][j] > 1.0) d->kern[s][j] = 1.0 / d->kern[s][j];
        d->state[j] = (int8_t)argmax3(d->kern[0][j], d->kern[1][j], d->kern[2][j]);
    }
    d->n_eval = n_eval;
    free(have); free(samp);
    for (int s = 0; s < 3; s++) free(kd[s].x_scaled);
    return ORC_OK;
}
