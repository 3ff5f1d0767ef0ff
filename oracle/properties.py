"""Size-independent properties of the hot path's outputs, for inputs too large for the CPU oracle (BASELINE configs[2..4]).

Path A: the rows are a lossless encoding of the alignment -- decoding them against the reference (apply SNVs, delete DEL
spans, insert INS sequences at their *left-shifted* positions) must give back the aligned contig; row counts must equal
what the CIGARs say; rows must come in emission order. Path B: INDEX strictly increasing inside the window, states in
range, the dominant state is the planted orientation.

TEST INFRASTRUCTURE (like the rest of oracle/): host-side numpy, used by tests/ and the full-size drivers under profiles/.
Nothing under pav_b200/ imports it.
"""
import numpy as np

from pav_b200 import fasta

OP_I, OP_D, OP_X, OP_H, OP_S = 1, 2, 8, 5, 4


def cigar_counts(ops):
    code = ops & 15
    return int((ops[code == OP_X] >> 4).sum()), int(((code == OP_I) | (code == OP_D)).sum())


def check_emission_order(snv, indel):
    """Rows sorted by (record, op); SNV rows of one op by position."""
    for rows in (snv, indel):
        if len(rows) > 1:
            key = rows['rec'].astype(np.int64) << 32 | rows['op_idx'].astype(np.int64)
            assert (np.diff(key) >= 0).all(), 'rows out of emission order'
    if len(snv) > 1:
        same = (snv['rec'][1:] == snv['rec'][:-1]) & (snv['op_idx'][1:] == snv['op_idx'][:-1])
        assert (np.diff(snv['pos_ref'].astype(np.int64))[same] == 1).all(), 'SNV rows of one X run are not consecutive'


def check_snv_bases(snv, rec_ref_arr, rec_tig_arr, rec_rev):
    """Every SNV row points at a reference base and a contig base that differ (case-insensitively). Returns #rows checked."""
    n = 0
    bounds = np.searchsorted(snv['rec'], np.arange(len(rec_ref_arr) + 1))
    for r in range(len(rec_ref_arr)):
        a, b = bounds[r], bounds[r + 1]
        if a == b:
            continue
        rb = fasta.UPPER[rec_ref_arr[r][snv['pos_ref'][a:b]]]
        tb = rec_tig_arr[r][snv['qry_pos'][a:b]]
        tb = fasta.UPPER[fasta.COMPLEMENT[tb] if rec_rev[r] else tb]
        assert (rb != tb).all(), f'record {r}: SNV row whose REF and ALT bases agree'
        n += b - a
    return n


def decode_record(ref_arr, pos0, end0, rev, tig_arr, snv, indel):
    """Apply the rows of one record to reference[pos0:end0] -> (decoded query in reference orientation, N-taint flags)."""
    seg = fasta.UPPER[ref_arr[pos0:end0]]
    taint = seg == ord('N')     # the contig carries real bases where the reference has N: those positions cannot be compared
    if len(snv):
        alt = tig_arr[snv['qry_pos']]
        alt = fasta.UPPER[fasta.COMPLEMENT[alt] if rev else alt]
        seg[snv['pos_ref'] - pos0] = alt
        taint[snv['pos_ref'] - pos0] = False
    keep = np.ones(len(seg), dtype=bool)
    dels = indel[indel['svtype'] == 1]
    if len(dels):
        ln = dels['svlen'].astype(np.int64)
        idx = np.repeat(dels['pos'].astype(np.int64) - pos0, ln) + (np.arange(int(ln.sum())) - np.repeat(np.cumsum(ln) - ln, ln))
        keep[idx] = False
    ins = indel[indel['svtype'] == 0]
    if len(ins):
        ln = ins['svlen'].astype(np.int64)
        within = np.arange(int(ln.sum())) - np.repeat(np.cumsum(ln) - ln, ln)
        if rev:   # SEQ = reverse complement of forward contig[qry_pos : qry_end]
            src = np.repeat(ins['qry_end'].astype(np.int64) - 1, ln) - within
            vals = fasta.UPPER[fasta.COMPLEMENT[tig_arr[src]]]
        else:
            src = np.repeat(ins['qry_pos'].astype(np.int64), ln) + within
            vals = fasta.UPPER[tig_arr[src]]
        at = np.repeat(ins['pos'].astype(np.int64) - pos0, ln)   # inserted before reference position POS (already left-shifted)
        seg = np.insert(seg, at, vals)
        keep = np.insert(keep, at, True)
        taint = np.insert(taint, at, False)
    return seg[keep], taint[keep]


def check_roundtrip(rec, ref_arr, pos0, end0, rev, tig_arr, clip_l, clip_r, snv, indel):
    """decode(reference, rows of the record) == aligned part of the contig (reference orientation)."""
    got, taint = decode_record(ref_arr, pos0, end0, rev, tig_arr, snv, indel)
    q = fasta.UPPER[fasta.COMPLEMENT[tig_arr[::-1]] if rev else tig_arr]
    q = q[clip_l:len(q) - clip_r]
    assert len(got) == len(q), f'record {rec}: decoded length {len(got)} != aligned contig length {len(q)}'
    bad = (got != q) & ~taint
    assert not bad.any(), f'record {rec}: decoded sequence differs from the contig at {int(bad.sum())} positions (first {int(np.flatnonzero(bad)[0])})'
    return int(len(q))


def clips_of(ops_rec):
    """(leading, trailing) clipped query bases of one record's packed ops."""
    cl = cr = 0
    for op in ops_rec[:2]:
        if (op & 15) in (OP_H, OP_S):
            cl += int(op >> 4)
        else:
            break
    for op in ops_rec[::-1][:2]:
        if (op & 15) in (OP_H, OP_S):
            cr += int(op >> 4)
        else:
            break
    return cl, cr


def check_walk(table_rows, ops, op_off, snv, indel, rec_ref_arr, rec_tig_arr, roundtrip_every=1):
    """All Path-A properties for a batch. ``table_rows`` = list of (POS, END, REV) per record. Returns a summary dict."""
    n_x, n_id = cigar_counts(ops)
    assert len(snv) == n_x and len(indel) == n_id, f'row counts ({len(snv)}, {len(indel)}) != CIGAR counts ({n_x}, {n_id})'
    check_emission_order(snv, indel)
    rec_rev = [bool(t[2]) for t in table_rows]
    n_snv_checked = check_snv_bases(snv, rec_ref_arr, rec_tig_arr, rec_rev)
    assert (indel['left_shift'] >= 0).all() and (indel[['hom_ref_l', 'hom_ref_r', 'hom_tig_l', 'hom_tig_r']].view(np.int32) >= 0).all()
    is_del = indel['svtype'] == 1
    assert (indel['end'][is_del] - indel['pos'][is_del] == indel['svlen'][is_del]).all()
    assert (indel['end'][~is_del] - indel['pos'][~is_del] == 1).all()
    assert (indel['qry_end'][~is_del] - indel['qry_pos'][~is_del] == indel['svlen'][~is_del]).all()
    sb = np.searchsorted(snv['rec'], np.arange(len(table_rows) + 1))
    ib = np.searchsorted(indel['rec'], np.arange(len(table_rows) + 1))
    bases = n_rt = 0
    for r in range(0, len(table_rows), max(int(roundtrip_every), 1)):
        pos0, end0, rev = table_rows[r]
        cl, cr = clips_of(ops[op_off[r]:op_off[r + 1]])
        bases += check_roundtrip(r, rec_ref_arr[r], int(pos0), int(end0), bool(rev), rec_tig_arr[r], cl, cr, snv[sb[r]:sb[r + 1]], indel[ib[r]:ib[r + 1]])
        n_rt += 1
    return {'snv_rows': int(len(snv)), 'indel_rows': int(len(indel)), 'snv_base_checks': int(n_snv_checked),
            'roundtrip_records': n_rt, 'roundtrip_bases': int(bases)}


def check_density_window(res, win_len, k, expect_state=None, min_frac=0.5):
    """One window of a density batch (dict from pavlib.density._split)."""
    if res['status'] != 0:
        return False
    idx = res['INDEX']
    assert len(idx) <= win_len - k + 1
    if len(idx):
        assert idx[0] >= 0 and idx[-1] <= win_len - k and (np.diff(idx.astype(np.int64)) > 0).all(), 'INDEX not strictly increasing'
        assert np.isin(res['STATE_MER'], (0, 1, 2)).all()
        if res['smoothed']:
            assert np.isin(res['STATE'], (0, 1, 2)).all()
            for c in ('KERN_FWD', 'KERN_FWDREV', 'KERN_REV'):
                v = res[c]
                # NaN = state absent / below --minstatecount (density.py:181-190); numpy.interp's slope*(x-x0)+y0 may leave a negative denormal
                assert not (v < -1e-12).any() and not (v > 1.0 + 1e-9).any(), c
            if expect_state is not None:
                assert (res['STATE'] == expect_state).mean() >= min_frac, 'dominant state is not the planted orientation'
    return True
