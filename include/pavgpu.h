/*
 * pavgpu.h -- C ABI of libpavgpu.so: the B200 (sm_100a) implementation of PAV's variant-calling
 * hot path. Plain pointers and sizes only; no Python, torch or C++ types cross this boundary.
 *
 * The reference (EichlerLab/pav 2.4.6.0) is pure Python and has no FFI, so each entry point cites
 * the reference *Python* interface a binding would replace (paths relative to the PAV tree). The
 * ctypes binding that PAV's pavlib would load is shown in INTEGRATION.md and shipped as
 * pav_b200/_capi.py.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on failure; pavgpu_last_error() gives the message
 *     (thread-local). CUDA failures never abort the process (the Python layer raises RuntimeError,
 *     mirroring how pavlib reports errors: pavlib/cigarcall.py:289-307, pavlib/inv.py:268-281).
 *   - "host" pointers are caller-owned unless the name ends in _out; *_out buffers are allocated
 *     by the library and released with pavgpu_free_host().
 *   - coordinates are 0-based; sequence ids index the sequence store they refer to.
 */
#ifndef PAVGPU_H
#define PAVGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAVGPU_OK 0
#define PAVGPU_ERR_CUDA (-1)
#define PAVGPU_ERR_ARG (-2)
#define PAVGPU_ERR_NOMEM (-3)
#define PAVGPU_ERR_CIGAR_SYNTAX (-4) /* pavlib/align/align.py:308-318 */
#define PAVGPU_ERR_NCCL (-5)
#define PAVGPU_INV_FAIL 125          /* pavlib/constants.py:55 ERR_INV_FAIL (per-window status, not a return code) */

/* BAM op codes packed as (len << 4) | code ("packed 4-bit CIGAR ops"); pavlib/align/align.py:12-20 */
#define PAVGPU_OP_M 0
#define PAVGPU_OP_I 1
#define PAVGPU_OP_D 2
#define PAVGPU_OP_N 3
#define PAVGPU_OP_S 4
#define PAVGPU_OP_H 5
#define PAVGPU_OP_P 6
#define PAVGPU_OP_EQ 7
#define PAVGPU_OP_X 8

typedef struct pavgpu_ctx pavgpu_ctx;           /* one CUDA device + stream */
typedef struct pavgpu_seqstore pavgpu_seqstore; /* 2-bit + N-mask planes resident in HBM */
typedef struct pavgpu_cigar_batch pavgpu_cigar_batch;
typedef struct pavgpu_density_batch pavgpu_density_batch;

/* ---------------------------------------------------------------- context ------------------- */
const char *pavgpu_last_error(void);
int pavgpu_device_count(void);                       /* number of CUDA devices, or <0 */
int pavgpu_ctx_create(int device, pavgpu_ctx **ctx_out);
void pavgpu_ctx_destroy(pavgpu_ctx *ctx);
int pavgpu_ctx_device(const pavgpu_ctx *ctx);
void pavgpu_free_host(void *p);
/* Pinned host buffer from the context's pool (released with pavgpu_free_host): staging for sequences a caller is about to hand
 * to pavgpu_seqstore_create -- copies out of pinned memory run at PCIe speed instead of through the driver's bounce buffers.
 * Replaces nothing in the reference (pysam hands out Python strings, pavlib/cigarcall.py:59-66). */
int pavgpu_host_alloc(pavgpu_ctx *ctx, size_t bytes, void **buf_out);
/* Evict L2 between timed iterations: overwrites a scratch buffer of `bytes` (> 126 MB L2) on the context stream. */
int pavgpu_l2_flush(pavgpu_ctx *ctx, size_t bytes);

/* ---------------------------------------------------------------- sequence store ------------ */
/*
 * Replaces pysam.FastaFile(...).fetch() + Bio reverse_complement + str.upper() on whole
 * chromosomes/contigs (pavlib/cigarcall.py:58-75, pavlib/seq.py:328-360): sequences are uploaded
 * once as ASCII, packed on the GPU to a 2-bit plane (A0 C1 G2 T3, case-insensitive; first base in
 * the most significant bits of each 64-bit word -- kanapy/util/kmer.py:50-69 order) and a 1-bit
 * "not ACGTacgt" plane. Reverse-complement access is index arithmetic on the device.
 */
int pavgpu_seqstore_create(pavgpu_ctx *ctx, int32_t n_seq, const uint8_t *const *seq_ascii,
                           const int64_t *seq_len, pavgpu_seqstore **store_out);
/* Build from already packed planes (a packed-reference sidecar, pav_b200/sidecar.py). The byte sizes of the caller's buffers are
 * checked against the layout this library derives from seq_len (PAVGPU_ERR_ARG on a mismatch: planes written under another
 * alignment / guard layout, or a truncated file). */
int pavgpu_seqstore_create_packed(pavgpu_ctx *ctx, int32_t n_seq, const int64_t *seq_len,
                                  const uint64_t *pack2_host, size_t pack2_bytes, const uint32_t *nmask_host, size_t nmask_bytes,
                                  pavgpu_seqstore **store_out);
/* Allocate planes for n_seq sequences without filling them (receiver side of a device broadcast). */
int pavgpu_seqstore_create_empty(pavgpu_ctx *ctx, int32_t n_seq, const int64_t *seq_len,
                                 pavgpu_seqstore **store_out);
void pavgpu_seqstore_free(pavgpu_seqstore *store);
int32_t pavgpu_seqstore_n_seq(const pavgpu_seqstore *store);
int64_t pavgpu_seqstore_total_bases(const pavgpu_seqstore *store); /* incl. alignment padding */
/* Device pointers + sizes in bytes of the two planes (for ncclBroadcast by the caller's communicator). */
int pavgpu_seqstore_planes(const pavgpu_seqstore *store, void **d_pack2, size_t *pack2_bytes,
                           void **d_nmask, size_t *nmask_bytes);
/* Copy the planes back to the host (tests, packed sidecar files). Sizes from pavgpu_seqstore_planes(). */
int pavgpu_seqstore_export(const pavgpu_seqstore *store, uint64_t *pack2_host, uint32_t *nmask_host);
/* Checksums of the two planes computed on the device (sums_out[0] 2-bit plane, [1] mask plane; order-independent 64-bit sums of
 * position-salted words): every rank of a multi-GPU run compares what pavgpu_seqstore_broadcast left in its HBM with rank 0's. */
int pavgpu_seqstore_checksum(const pavgpu_seqstore *store, uint64_t sums_out[2]);
/* Base offset (in bases) of sequence i inside the planes. */
int64_t pavgpu_seqstore_offset(const pavgpu_seqstore *store, int32_t seq_id);

/* ---------------------------------------------------------------- Path A: CIGAR walk -------- */
/* Host tokenizer: replaces pavlib.align.cigar_str_to_tuples (pavlib/align/align.py:286-322). */
typedef struct {
    int32_t code;      /* 0 none, 2 missing length, 3 unknown operation, 4 ran off the string (IndexError) */
    int32_t rec;       /* record whose CIGAR is malformed */
    int64_t op_index;  /* number of well-formed ops before the error in that record */
    int64_t text_pos;  /* position in the CIGAR string the reference reports */
    int32_t ch;        /* character the reference prints for "unknown operation" */
} pavgpu_parse_err;

/* cigar_text: concatenated CIGAR strings; text_off[n_rec+1] byte offsets. On success *ops_out holds
 * all packed ops, op_off_out[n_rec+1] their per-record offsets. A syntax error in record r truncates
 * r's ops at the error and is reported in *err (the walk must still process everything before it:
 * the reference raises lazily, in record order). */
int pavgpu_cigar_parse(const char *cigar_text, const int64_t *text_off, int32_t n_rec,
                       uint32_t **ops_out, int64_t *op_off_out, pavgpu_parse_err *err);

/* Per-record CIGAR summary for the SAM -> alignment-table step (pavlib/align/align.py:666-794, what get_align_bed derives from
 * pysam's cigartuples record by record): computed on the device, one warp per record, from the packed ops of pavgpu_cigar_parse. */
typedef struct { /* 48 B */
    int64_t ref_bp;       /* bases of the reference covered by the core: sum of = X D N lengths (:735-741) */
    int64_t qry_bp;       /* bases of the query covered by the core: sum of = X I lengths */
    int64_t lead;         /* summed lengths of the clip ops (S / H) before the first non-clip op (:706-716) */
    int64_t trail;        /* ... after the last non-clip op */
    int32_t first_body;   /* op number of the first / last non-clip op inside the record; -1: the record has only clips */
    int32_t last_body;
    int32_t clip_h_first; /* length of op 0 when it is H, else 0 */
    int32_t lead_s;       /* length of the first op that is not H when that op is S, else 0 (pysam's query_alignment_start) */
    int32_t flags;        /* bit 0: an M op is present (rejected, :700-704); bit 1: a clip op lies between first_body and last_body */
    int32_t n_ops;
} pavgpu_cigar_rec_stats;
int pavgpu_cigar_record_stats(pavgpu_ctx *ctx, const uint32_t *ops, const int64_t *op_off, int32_t n_rec,
                              pavgpu_cigar_rec_stats *stats_out);

typedef struct { /* 16 B: one row per mismatched base (pavlib/cigarcall.py:98-135) */
    int32_t pos_ref;  /* POS (END = POS + 1) */
    int32_t qry_pos;  /* 0-based position on the forward contig; QRY_REGION = qry_pos+1 .. qry_pos+1 */
    int32_t rec;      /* row number in the alignment table */
    int32_t op_idx;   /* CIGAR op number inside the record (emission-order key) */
} pavgpu_snv_row;

typedef struct { /* 64 B: one row per I / D op (pavlib/cigarcall.py:141-282) */
    int32_t rec, op_idx;
    int32_t svtype;   /* 0 INS, 1 DEL */
    int32_t svlen;
    int32_t pos, end; /* INS: left-shifted; DEL: NOT shifted (reference quirk, cigarcall.py:254-258) */
    int32_t qry_pos, qry_end; /* forward-contig, 0-based half-open; QRY_REGION = qry_pos+1 .. qry_end */
    int32_t left_shift;
    int32_t hom_ref_l, hom_ref_r, hom_tig_l, hom_tig_r;
    int32_t seq_start; /* INS: offset of SEQ in the reference-oriented contig; DEL: offset in the chromosome */
    int32_t pad[2];
} pavgpu_indel_row;

typedef struct {
    int32_t code;     /* 0 none, 1 illegal op (cigarcall.py:289-307) */
    int32_t rec;
    int64_t op_index; /* 0-based op number in the record (reference prints op_index + 1) */
    int32_t opcode;   /* BAM code of the offending op */
    int32_t pos_ref, pos_qry; /* walk position when the op was reached */
} pavgpu_cigar_err;

typedef struct {
    float ms_h2d, ms_kernels, ms_d2h;       /* CUDA-event times on the context stream */
    float ms_scan, ms_emit, ms_homology;    /* per-kernel breakdown */
    int64_t n_ops, n_snv, n_indel, n_chunks;
    int32_t kernel_launches;
    int32_t homology_tiled;                 /* homology kernel used: 0 gathers, 1 per-warp shared-memory tiles, 2 per-indel neighbourhoods (cp.async),
                                               3 per-indel neighbourhoods (bulk copies + mbarrier), 4 gathers with CTA-pooled scan rests */
    float ms_count;                         /* per-record row counts + record scan on the device (single-pass walk) */
    int32_t walk_passes;                    /* 1 single-pass walk (default), 3 multi-pass walk, 0 empty batch */
    int32_t graph;                          /* 1 when the step was replayed as one CUDA graph */
    int32_t pad0;
} pavgpu_cigar_stats;

/* Upload a batch of alignment records (SoA). All arrays have n_rec entries except op_off (n_rec+1).
 * ref_seq_id / qry_seq_id index ref_store / qry_store; rev != 0 means the contig is aligned as its
 * reverse complement (pavlib/cigarcall.py:63-72). */
int pavgpu_cigar_batch_create(pavgpu_ctx *ctx, int32_t n_rec, const int32_t *ref_seq_id,
                              const int32_t *qry_seq_id, const int32_t *pos, const uint8_t *rev,
                              const uint32_t *ops, const int64_t *op_off, pavgpu_cigar_batch **batch_out);
void pavgpu_cigar_batch_free(pavgpu_cigar_batch *batch);
/* Run the walk on the device: inputs and outputs stay in HBM (this is what bench.py times as `value`). */
int pavgpu_cigar_batch_run(pavgpu_cigar_batch *batch, const pavgpu_seqstore *ref_store,
                           const pavgpu_seqstore *qry_store, pavgpu_cigar_stats *stats);
/* Copy the rows of the last run to the host (emission order: record, op, base). */
int pavgpu_cigar_batch_fetch(pavgpu_cigar_batch *batch, pavgpu_snv_row **snv_out, int64_t *n_snv,
                             pavgpu_indel_row **indel_out, int64_t *n_indel, pavgpu_cigar_err *err);
/* One-shot host-buffer call = create + run + fetch + free. Replaces the loop body of
 * pavlib.cigarcall.make_insdel_snv_calls (pavlib/cigarcall.py:50-311). */
int pavgpu_cigar_call(pavgpu_ctx *ctx, const pavgpu_seqstore *ref_store, const pavgpu_seqstore *qry_store,
                      int32_t n_rec, const int32_t *ref_seq_id, const int32_t *qry_seq_id, const int32_t *pos,
                      const uint8_t *rev, const uint32_t *ops, const int64_t *op_off,
                      pavgpu_snv_row **snv_out, int64_t *n_snv, pavgpu_indel_row **indel_out, int64_t *n_indel,
                      pavgpu_cigar_err *err, pavgpu_cigar_stats *stats);

/* Single-call parity helpers for pavlib.call.left_homology / right_homology (pavlib/call.py:542-647)
 * on upper-cased ASCII strings (runs the same device routine the walk uses, on one thread). */
int pavgpu_homology(pavgpu_ctx *ctx, int32_t n, const uint8_t *seq, int64_t seq_len, const uint8_t *sv,
                    int64_t sv_len, const int64_t *pos, int32_t *left_out, int32_t *right_out);

/* ---------------------------------------------------------------- Path B: k-mer density ----- */
/* One window = one run of scripts/density.py (scripts/density.py:423-571). */
typedef struct {
    int32_t ref_seq_id; int32_t tig_seq_id;
    int32_t ref_pos, ref_end;   /* --refregion, 0-based half-open, forward strand (pavlib/seq.py:316) */
    int32_t tig_pos, tig_end;   /* --tigregion, forward contig coordinates (never reverse-complemented) */
    int32_t rev;                /* -r true: reverse-complement the reference k-mer set (density.py:538-539) */
    int32_t srs;                /* --staterunsmooth */
} pavgpu_density_window;

typedef struct {
    int32_t k;                  /* -k: 1 <= k <= 31 on the GPU path (exact 62-bit keys, all-ones is the empty-slot sentinel); larger k fails with PAVGPU_ERR_ARG */
    int32_t min_informative;    /* --mininf 2000 */
    int32_t min_state_count;    /* --minstatecount 20 */
    int32_t max_ref_kmer_count; /* MAX_REF_KMER_COUNT 100 (density.py:47) */
    double smooth;              /* --densmooth 1 */
    double delta;               /* --staterundelta 0.005 */
} pavgpu_density_params;

typedef struct {
    int32_t status;     /* 0 ok, 125 soft failure (no reference k-mers / k-mer count > max) */
    int32_t smoothed;   /* 0: fewer than min_informative rows => STATE = -1 and no KERN_* (density.py:193-194) */
    int64_t row_off;    /* first row of this window in the column arrays */
    int64_t n_rows;
    int64_t n_eval;     /* lattice points with a full KDE evaluation (sampled + filled) */
} pavgpu_density_result;

typedef struct {
    float ms_h2d, ms_kernels, ms_d2h;
    float ms_kmer, ms_kde, ms_fill;
    int64_t bases, rows, kde_pairs;
    int32_t kernel_launches;
    int32_t kmer_tables_on_chip;   /* windows whose reference k-mer table was built and probed in shared memory (kmer_window_kernel); the rest used tables in HBM */
} pavgpu_density_stats;

void pavgpu_density_default_params(pavgpu_density_params *p);
/* Batched scan: columns are concatenated over windows in window order (row_off / n_rows per window):
 * KMER (uint64), INDEX (int32), STATE_MER (int8), STATE (int8), KERN_FWD/FWDREV/REV (float64; NaN rows
 * for un-smoothed windows). */
int pavgpu_density_batch_create(pavgpu_ctx *ctx, int32_t n_win, const pavgpu_density_window *win,
                                const pavgpu_density_params *params, pavgpu_density_batch **batch_out);
void pavgpu_density_batch_free(pavgpu_density_batch *batch);
int pavgpu_density_batch_run(pavgpu_density_batch *batch, const pavgpu_seqstore *ref_store,
                             const pavgpu_seqstore *tig_store, pavgpu_density_stats *stats);
int pavgpu_density_batch_fetch(pavgpu_density_batch *batch, pavgpu_density_result *res /* n_win */,
                               uint64_t **kmer_out, int32_t **index_out, int8_t **state_mer_out,
                               int8_t **state_out, double **kern_fwd_out, double **kern_fwdrev_out,
                               double **kern_rev_out, int64_t *n_rows_total);

/* One run of equal STATE in row order: the tuple pavlib.density.rl_encoder yields (pavlib/density.py:330-361). */
typedef struct {
    int32_t state;        /* -1 for a window returned un-smoothed (scripts/density.py:193-194) */
    int32_t count;        /* rows in the run */
    int32_t first_index;  /* INDEX of its first row */
    int32_t last_index;   /* INDEX of its last row */
} pavgpu_state_run;

/* Run lengths of STATE for every window of the batch: what pavlib.inv.scan_for_inv decides from after every expansion
 * (pavlib/inv.py:294-342) -- 16 bytes per run instead of 38 bytes per row. res[n_win], run_off[n_win + 1] caller-owned;
 * *runs_out library-allocated (pavgpu_free_host), runs of window w at [run_off[w], run_off[w + 1]). */
int pavgpu_density_batch_fetch_runs(pavgpu_density_batch *batch, pavgpu_density_result *res, pavgpu_state_run **runs_out,
                                    int64_t *run_off, int64_t *n_runs_total);
/* All columns of ONE window into caller-owned arrays of res[win].n_rows entries each (NULL = skip the column): the window that
 * becomes a call and whose table the rule writes out (pavlib/inv.py:440-454, rules/call_inv.snakefile:279-282). */
int pavgpu_density_batch_fetch_window(pavgpu_density_batch *batch, int32_t win, uint64_t *kmer, int32_t *index,
                                      int8_t *state_mer, int8_t *state, double *kern_fwd, double *kern_fwdrev, double *kern_rev);

/* ---------------------------------------------------------------- coordinate lifts ---------- */
/* Reference <-> contig lifts through alignment records (pavlib/align/lift.py:177-331 lift_to_sub, :380-476 lift_to_qry), batched:
 * the index holds, for every packed CIGAR op of every record, its first reference and first contig coordinate (segmented prefix
 * sums on the device, one CTA per record); a lift is a binary search inside the record plus the reference's block rules (one-base
 * blocks map to the end of their image, reverse records flip the contig coordinate, contig -> reference accepts the exact end of a
 * block). pos / rev / qry_len: POS, REV and contig length of each record. *bad_rec_out: first record with an op the lift does not
 * handle (anything but M I D S H = X), or -1. status[i]: 0 = lifted, 1 = no block (the reference raises RuntimeError). */
typedef struct pavgpu_lift_index pavgpu_lift_index;
int pavgpu_lift_index_create(pavgpu_ctx *ctx, const uint32_t *ops, const int64_t *op_off, int32_t n_rec, const int64_t *pos,
                             const uint8_t *rev, const int64_t *qry_len, pavgpu_lift_index **index_out, int32_t *bad_rec_out);
void pavgpu_lift_index_free(pavgpu_lift_index *index);
int pavgpu_lift_points(pavgpu_lift_index *index, int32_t n, const int32_t *rec, const int64_t *coord, int32_t to_qry,
                       int64_t *out, int32_t *status);

/* ---------------------------------------------------------------- multi-GPU ----------------- */
/* Reference broadcast over NVLink: rank 0 owns a filled store, the other ranks an empty
 * one of the same shape. The caller moves the 128-byte id from rank 0 to all ranks (any side
 * channel); the library dlopen()s libnccl.so.2 on first use. */
int pavgpu_nccl_unique_id(uint8_t id_out[128]);
/* The communicator made from `id` is kept for the life of the process, one per (device, rank, n_ranks): with id == NULL the
 * broadcast reuses it (ncclCommInitRank costs seconds, the broadcast milliseconds). Every rank must make the same choice;
 * pavgpu_nccl_comm_cached() tells whether this rank has one, pavgpu_nccl_comm_release_all() destroys them. */
int pavgpu_seqstore_broadcast(pavgpu_ctx *ctx, pavgpu_seqstore *store, const uint8_t id[128],
                              int32_t rank, int32_t n_ranks, float *ms_out);
int pavgpu_nccl_comm_cached(pavgpu_ctx *ctx, int32_t rank, int32_t n_ranks);
void pavgpu_nccl_comm_release_all(void);

#ifdef __cplusplus
}
#endif
#endif /* PAVGPU_H */
