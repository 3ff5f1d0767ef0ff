// lift.cu -- reference <-> contig coordinate lifts through alignment records, batched on the device.
//
// Reference semantics: pavlib/align/lift.py:177-331 (lift_to_sub), :380-476 (lift_to_qry): the reference keeps two interval trees per
// record (one node per aligned / inserted / deleted block) and answers one point per call. Here the blocks ARE the packed CIGAR ops:
// a segmented exclusive prefix sum of the per-op advances (one CTA per record) gives every op its first reference and its first
// contig coordinate, and a lift is a binary search over those inside the record plus the reference's block rules:
//   * an aligned block (M = X) longer than one base maps linearly; a one-base aligned block, an insertion (to the reference) or a
//     deletion (to the contig) maps to the END of its one-base image (lift.py:243-249 / :436-441 -- `if stop - start > 1`);
//   * contig -> reference on a reverse record flips the coordinate first (len - pos), reference -> contig flips it afterwards;
//   * contig -> reference: a position that is no block's but is exactly the end of one (pos - 1 inside it) uses that block (:226-238);
//     anything else is "no match" (status 1; the host raises the reference's RuntimeError).
// Which record covers a position (one candidate, none, several) stays on the host: it is a dictionary look-up per chromosome.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "liftcore.cuh"

namespace {

constexpr int LP_THREADS = 256;
constexpr int LP_ITEMS = 4;

// One CTA per record: exclusive prefix sums of the reference / contig advances of its ops, POS added to the reference one.
__global__ void __launch_bounds__(LP_THREADS)
lift_prefix_kernel(const uint32_t *__restrict__ ops, const int64_t *__restrict__ op_off, const int64_t *__restrict__ pos, int64_t *__restrict__ ref_start,
                   int64_t *__restrict__ qry_start, int32_t *__restrict__ bad_rec)
{
    const int32_t r = blockIdx.x;
    const int64_t o0 = op_off[r], n = op_off[r + 1] - o0;
    __shared__ long long s_r[LP_THREADS / 32], s_q[LP_THREADS / 32];
    __shared__ long long s_carry_r, s_carry_q;
    if (threadIdx.x == 0) { s_carry_r = pos[r]; s_carry_q = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned bad = 0;
    for (int64_t base = 0; base < n; base += LP_THREADS * LP_ITEMS) {
        const int64_t i0 = base + (int64_t)threadIdx.x * LP_ITEMS;
        long long ra[LP_ITEMS], qa[LP_ITEMS], tr = 0, tq = 0;
#pragma unroll
        for (int k = 0; k < LP_ITEMS; k++) {
            ra[k] = qa[k] = 0;
            if (i0 + k < n) lift_op_advance(ops[o0 + i0 + k], ra[k], qa[k], bad);
            tr += ra[k]; tq += qa[k];
        }
        long long ir = tr, iq = tq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long ur = __shfl_up_sync(0xffffffffu, ir, d), uq = __shfl_up_sync(0xffffffffu, iq, d);
            if (lane >= d) { ir += ur; iq += uq; }
        }
        if (lane == 31) { s_r[wid] = ir; s_q[wid] = iq; }
        __syncthreads();
        long long br = s_carry_r, bq = s_carry_q;
        for (int w = 0; w < wid; w++) { br += s_r[w]; bq += s_q[w]; }
        long long er = br + ir - tr, eq = bq + iq - tq;      // exclusive prefix of this thread's first op
#pragma unroll
        for (int k = 0; k < LP_ITEMS; k++) {
            if (i0 + k < n) { ref_start[o0 + i0 + k] = er; qry_start[o0 + i0 + k] = eq; }
            er += ra[k]; eq += qa[k];
        }
        __syncthreads();
        if (threadIdx.x == LP_THREADS - 1) { s_carry_r = er; s_carry_q = eq; }   // the last thread holds the chunk's inclusive total
        __syncthreads();
    }
    if (__syncthreads_or((int)bad) && threadIdx.x == 0) atomicMin(bad_rec, r);
}

__global__ void __launch_bounds__(128)
lift_points_kernel(const uint32_t *__restrict__ ops, const int64_t *__restrict__ op_off, const int64_t *__restrict__ ref_start,
                   const int64_t *__restrict__ qry_start, const uint8_t *__restrict__ rev, const int64_t *__restrict__ qry_len, int32_t n,
                   const int32_t *__restrict__ rec, const int64_t *__restrict__ coord, int to_qry, int64_t *__restrict__ out, int32_t *__restrict__ status)
{
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int32_t r = rec[t];
    int64_t v = 0;
    status[t] = lift_point(ops, ref_start, qry_start, op_off[r], op_off[r + 1], rev[r], qry_len[r], to_qry, coord[t], v);
    out[t] = v;
}

}  // namespace

struct pavgpu_lift_index {
    pavgpu_ctx *ctx;
    int32_t n_rec;
    int64_t n_ops;
    void *arena;
    uint32_t *d_ops;
    int64_t *d_op_off, *d_ref_start, *d_qry_start, *d_qry_len;
    uint8_t *d_rev;
};

extern "C" __attribute__((visibility("default"))) void pavgpu_lift_index_free(pavgpu_lift_index *idx)
{
    if (!idx) return;
    if (idx->arena) { cudaSetDevice(idx->ctx->device); pav_dev_free(idx->ctx, idx->arena); }
    delete idx;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_lift_index_create(pavgpu_ctx *ctx, const uint32_t *ops, const int64_t *op_off, int32_t n_rec,
                                                                                 const int64_t *pos, const uint8_t *rev, const int64_t *qry_len,
                                                                                 pavgpu_lift_index **index_out, int32_t *bad_rec_out)
{
    if (!ctx || !index_out || n_rec < 0 || (n_rec > 0 && (!op_off || !pos || !rev || !qry_len))) { pav_set_error("lift_index_create: bad argument"); return PAVGPU_ERR_ARG; }
    *index_out = nullptr;
    if (bad_rec_out) *bad_rec_out = -1;
    const int64_t n_ops = n_rec > 0 ? op_off[n_rec] : 0;
    if (n_rec > 0 && (op_off[0] != 0 || n_ops < 0 || (n_ops > 0 && !ops))) { pav_set_error("lift_index_create: op_off must start at 0 and ascend"); return PAVGPU_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t no = (size_t)std::max<int64_t>(n_ops, 1), nr = (size_t)std::max<int32_t>(n_rec, 1);
    const size_t o_ops = 0, o_off = o_ops + up(4 * no), o_rs = o_off + up(8 * (nr + 1)), o_qs = o_rs + up(8 * no), o_ql = o_qs + up(8 * no),
                 o_pos = o_ql + up(8 * nr), o_rev = o_pos + up(8 * nr), o_bad = o_rev + up(nr), total = o_bad + 256;
    pavgpu_lift_index *idx = new pavgpu_lift_index();
    idx->ctx = ctx; idx->n_rec = n_rec; idx->n_ops = n_ops; idx->arena = nullptr;
    int rc = [&]() -> int {
        CUDA_TRY(pav_dev_alloc(ctx, total, &idx->arena));
        char *base = static_cast<char *>(idx->arena);
        idx->d_ops = reinterpret_cast<uint32_t *>(base + o_ops); idx->d_op_off = reinterpret_cast<int64_t *>(base + o_off);
        idx->d_ref_start = reinterpret_cast<int64_t *>(base + o_rs); idx->d_qry_start = reinterpret_cast<int64_t *>(base + o_qs);
        idx->d_qry_len = reinterpret_cast<int64_t *>(base + o_ql); idx->d_rev = reinterpret_cast<uint8_t *>(base + o_rev);
        int64_t *d_pos = reinterpret_cast<int64_t *>(base + o_pos);
        int32_t *d_bad = reinterpret_cast<int32_t *>(base + o_bad);
        if (n_rec == 0) return PAVGPU_OK;
        if (n_ops > 0) CUDA_TRY(cudaMemcpyAsync(idx->d_ops, ops, 4 * (size_t)n_ops, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(idx->d_op_off, op_off, 8 * ((size_t)n_rec + 1), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_pos, pos, 8 * (size_t)n_rec, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(idx->d_qry_len, qry_len, 8 * (size_t)n_rec, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(idx->d_rev, rev, (size_t)n_rec, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemsetAsync(d_bad, 0x7f, 4, st));
        lift_prefix_kernel<<<(unsigned)n_rec, LP_THREADS, 0, st>>>(idx->d_ops, idx->d_op_off, d_pos, idx->d_ref_start, idx->d_qry_start, d_bad);
        CUDA_TRY(cudaGetLastError());
        int32_t bad = 0;
        CUDA_TRY(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (bad_rec_out && bad >= 0 && bad < n_rec) *bad_rec_out = bad;     // a record with an op the lift does not handle (N, P, ...): lift.py:165-168
        return PAVGPU_OK;
    }();
    if (rc != PAVGPU_OK) { pavgpu_lift_index_free(idx); return rc; }
    *index_out = idx;
    return PAVGPU_OK;
}

extern "C" __attribute__((visibility("default"))) int pavgpu_lift_points(pavgpu_lift_index *idx, int32_t n, const int32_t *rec, const int64_t *coord, int32_t to_qry,
                                                                           int64_t *out, int32_t *status)
{
    if (!idx || n < 0 || (n > 0 && (!rec || !coord || !out || !status))) { pav_set_error("lift_points: bad argument"); return PAVGPU_ERR_ARG; }
    if (n == 0) return PAVGPU_OK;
    for (int32_t i = 0; i < n; i++)
        if (rec[i] < 0 || rec[i] >= idx->n_rec) { pav_set_error("lift_points: record %d of point %d is outside the index", rec[i], i); return PAVGPU_ERR_ARG; }
    pavgpu_ctx *ctx = idx->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_rec = 0, o_co = up(4 * (size_t)n), o_out = o_co + up(8 * (size_t)n), o_st = o_out + up(8 * (size_t)n), total = o_st + up(4 * (size_t)n);
    void *arena = nullptr;
    CUDA_TRY(pav_dev_alloc(ctx, total, &arena));
    char *base = static_cast<char *>(arena);
    int rc = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(base + o_rec, rec, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(base + o_co, coord, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
        lift_points_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(idx->d_ops, idx->d_op_off, idx->d_ref_start, idx->d_qry_start, idx->d_rev, idx->d_qry_len, n,
                                                                         reinterpret_cast<int32_t *>(base + o_rec), reinterpret_cast<int64_t *>(base + o_co), to_qry ? 1 : 0,
                                                                         reinterpret_cast<int64_t *>(base + o_out), reinterpret_cast<int32_t *>(base + o_st));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out, base + o_out, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(status, base + o_st, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return PAVGPU_OK;
    }();
    pav_dev_free(ctx, arena);
    return rc;
}
