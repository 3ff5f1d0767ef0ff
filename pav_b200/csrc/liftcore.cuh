// liftcore.cuh -- the arithmetic of a coordinate lift over the packed CIGAR ops of one record: per-op advances, the search for the
// block that holds a position, and the block rules of the reference (pavlib/align/lift.py:177-331 lift_to_sub, :380-476
// lift_to_qry). Device code (every function is __device__ __forceinline__ under nvcc, used by lift.cu); the same text compiles as
// plain C++ when PAV_DEV is predefined -- tests/host_emul/lift_host.cpp runs these exact functions against the stored answers of the
// reference's AlignLift on a machine without a GPU.
#pragma once
#include <cstdint>

#include "pavgpu.h"

#ifndef PAV_DEV
#define PAV_DEV __device__ __forceinline__
#endif

// Reference / contig advance of one packed op (M = X D advance the reference, M = X I S H the contig: lift.py:155-176 builds its
// trees from exactly these ops); `bad` is set for anything else (N, P, ...: "Unhandled CIGAR operation", lift.py:165-168).
PAV_DEV void lift_op_advance(uint32_t op, long long &ra, long long &qa, unsigned &bad)
{
    const uint32_t code = op & 15u;
    const long long len = op >> 4;
    const bool m = code == PAVGPU_OP_M || code == PAVGPU_OP_EQ || code == PAVGPU_OP_X;
    ra = (m || code == PAVGPU_OP_D) ? len : 0;
    qa = (m || code == PAVGPU_OP_I || code == PAVGPU_OP_S || code == PAVGPU_OP_H) ? len : 0;
    if (!(m || code == PAVGPU_OP_I || code == PAVGPU_OP_D || code == PAVGPU_OP_S || code == PAVGPU_OP_H)) bad = 1u;
}

// Last op k of [lo, hi) with start[k] <= p, or lo - 1.
PAV_DEV int64_t lift_last_le(const int64_t *start, int64_t lo, int64_t hi, int64_t p)
{
    int64_t a = lo, b = hi;
    while (a < b) {
        const int64_t m = (a + b) >> 1;
        if (start[m] <= p) a = m + 1; else b = m;
    }
    return a - 1;
}

// The block containing q: the last op starting at or before q that advances in the source coordinate, if q is inside it and it is
// a block of the lift (reference -> contig: M = X D; contig -> reference: M = X I -- clips advance the contig but lift nothing).
PAV_DEV bool lift_find(const uint32_t *ops, const int64_t *start, int64_t lo, int64_t hi, int to_qry, int64_t q, int64_t &k, int64_t &len, bool &aligned)
{
    k = lift_last_le(start, lo, hi, q);
    while (k >= lo) {
        const uint32_t op = ops[k], code = op & 15u;
        len = op >> 4;
        aligned = code == PAVGPU_OP_M || code == PAVGPU_OP_EQ || code == PAVGPU_OP_X;
        const bool adv = to_qry ? (aligned || code == PAVGPU_OP_D) : (aligned || code == PAVGPU_OP_I || code == PAVGPU_OP_S || code == PAVGPU_OP_H);
        if (adv && len > 0) {
            const bool in_lift = to_qry ? true : (aligned || code == PAVGPU_OP_I);
            return in_lift && q < start[k] + len;
        }
        k--;     // ops that do not advance here share their start with the next one: step over them
    }
    return false;
}

// One lift inside record [lo, hi). ref_start / qry_start: first reference / contig coordinate of every op (exclusive prefix sums of
// the advances, POS added to the reference one). Returns 0 and the lifted coordinate, or 1 when no block holds the position.
PAV_DEV int lift_point(const uint32_t *ops, const int64_t *ref_start, const int64_t *qry_start, int64_t lo, int64_t hi, int rev, int64_t qry_len,
                       int to_qry, int64_t p, int64_t &out)
{
    const int64_t *start = to_qry ? ref_start : qry_start;
    const int64_t *image = to_qry ? qry_start : ref_start;
    if (!to_qry && rev) p = qry_len - p;
    int64_t k, len; bool aligned;
    bool ok = lift_find(ops, start, lo, hi, to_qry, p, k, len, aligned);
    if (!ok && !to_qry) ok = lift_find(ops, start, lo, hi, to_qry, p - 1, k, len, aligned) && start[k] + len == p;   // exactly the end of a block (lift.py:226-238)
    if (!ok) { out = 0; return 1; }
    // image interval of the block: aligned -> [image, image + len), else one base [image, image + 1)
    const int64_t d0 = image[k], d1 = aligned ? d0 + len : d0 + 1;
    int64_t v = (d1 - d0 > 1) ? d0 + (p - start[k]) : d1;
    if (to_qry && rev) v = qry_len - v;
    out = v;
    return 0;
}
