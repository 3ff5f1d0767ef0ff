// Host build of pav_b200/csrc/liftcore.cuh: the functions lift.cu's kernels call (per-op advances, block search, block rules),
// compiled as plain C++. Test infrastructure (tests/test_device_logic_cpu.py): the same source text the kernels inline is run against
// the stored answers of the reference's AlignLift on a machine without a GPU. What it cannot cover: the block-wide prefix scan of
// lift_prefix_kernel (restated here as a sequential exclusive sum over the same lift_op_advance) and the launch geometry.
#include <cstdint>

#define PAV_DEV static inline
#include "liftcore.cuh"

extern "C" {

// ref_start / qry_start of every op (what lift_prefix_kernel writes); returns the first record with an unhandled op, or -1.
int32_t emu_lift_prefix(const uint32_t *ops, const int64_t *op_off, int32_t n_rec, const int64_t *pos, int64_t *ref_start, int64_t *qry_start)
{
    int32_t bad_rec = -1;
    for (int32_t r = 0; r < n_rec; r++) {
        long long er = pos[r], eq = 0;
        unsigned bad = 0;
        for (int64_t i = op_off[r]; i < op_off[r + 1]; i++) {
            long long ra, qa;
            lift_op_advance(ops[i], ra, qa, bad);
            ref_start[i] = er; qry_start[i] = eq;
            er += ra; eq += qa;
        }
        if (bad && bad_rec < 0) bad_rec = r;
    }
    return bad_rec;
}

void emu_lift_points(const uint32_t *ops, const int64_t *op_off, const int64_t *ref_start, const int64_t *qry_start, const uint8_t *rev,
                     const int64_t *qry_len, int32_t n, const int32_t *rec, const int64_t *coord, int32_t to_qry, int64_t *out, int32_t *status)
{
    for (int32_t t = 0; t < n; t++) {
        const int32_t r = rec[t];
        int64_t v = 0;
        status[t] = lift_point(ops, ref_start, qry_start, op_off[r], op_off[r + 1], rev[r], qry_len[r], to_qry, coord[t], v);
        out[t] = v;
    }
}

}  // extern "C"
