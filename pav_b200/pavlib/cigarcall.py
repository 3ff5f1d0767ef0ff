"""Call SNVs and INS/DEL variants from CIGAR strings on the GPU.

Drop-in for the reference's ``pavlib.cigarcall`` (pavlib/cigarcall.py): same function name,
arguments, return value ``(df_snv, df_insdel)``, column order, dtypes (all ``object``), row order
(stable sort on ``#CHROM, POS, END, ID``), index values and exception messages. The per-record
Python loop of the reference (:50-311) is replaced by one ``pavgpu_cigar_call``; this module only
moves bytes in (FASTA -> HBM, CIGAR text -> packed ops) and formats the rows that come back.
"""
import numpy as np
import pandas as pd

from .. import device, fasta
from . import variant

CALL_SOURCE = 'CIGAR'          # pavlib/cigarcall.py:19
CALL_CIGAR_BATCH_COUNT = 10    # pavlib/cigarcall.py:21 (read by rules/align.snakefile:163)

SNV_COLUMNS = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'REF', 'ALT', 'HAP', 'QRY_REGION', 'QRY_STRAND', 'CI',
               'ALIGN_INDEX', 'CALL_SOURCE']
INSDEL_COLUMNS = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'HAP', 'QRY_REGION', 'QRY_STRAND', 'CI', 'ALIGN_INDEX',
                  'LEFT_SHIFT', 'HOM_REF', 'HOM_TIG', 'CALL_SOURCE', 'SEQ']

_S = np.dtypes.StringDType()
_CHR = np.array([chr(i) for i in range(256)], dtype=object)
_OP_CHAR = 'MIDNSHP=X'

# statistics of the last call (device timings etc.), for bench.py
last_stats = None


def _first_seen(values):
    seen = {}
    for v in values:
        if v not in seen:
            seen[v] = len(seen)
    return seen


def _join(*parts):
    """Element-wise string concatenation of StringDType arrays / scalars -> object array of str."""
    out = parts[0]
    for p in parts[1:]:
        out = np.strings.add(out, p)
    return out.astype(object)


def _istr(a):
    return np.asarray(a).astype(_S)


def _sort_order(chrom_codes, pos, end, ids):
    """Permutation equal to pandas' stable ``sort_values(['#CHROM','POS','END','ID'])``."""
    order = np.lexsort((end, pos, chrom_codes))
    if len(order) > 1:
        c, p, e = chrom_codes[order], pos[order], end[order]
        tie = (c[1:] == c[:-1]) & (p[1:] == p[:-1]) & (e[1:] == e[:-1])
        if tie.any():
            # groups of equal (chrom, pos, end): order them by ID, stable
            starts = np.flatnonzero(tie & ~np.concatenate(([False], tie[:-1])))
            for s in starts.tolist():
                t = s + 1
                while t < len(tie) and tie[t]:
                    t += 1
                grp = order[s:t + 1]
                grp_ids = [ids[i] for i in grp.tolist()]
                order[s:t + 1] = grp[np.array(sorted(range(len(grp)), key=grp_ids.__getitem__), dtype=np.int64)]
    return order


def _empty(columns):
    return pd.DataFrame([], columns=columns)


def _frame(cols, columns, order):
    data = {}
    for name in columns:
        v = cols[name]
        if isinstance(v, np.ndarray):
            data[name] = v[order]
        else:  # scalar column
            a = np.empty(len(order), dtype=object)
            a[:] = v
            data[name] = a
    return pd.DataFrame(data, columns=columns, index=pd.Index(order, dtype=np.int64), dtype=object)


def _raise_illegal(err, chrom, qry, align_index):
    rec = err.rec
    op = _OP_CHAR[err.opcode] if err.opcode < len(_OP_CHAR) else '?'
    if op == 'M':
        raise RuntimeError((
            'Illegal operation code in CIGAR string at operation {}: '
            'Alignments must be generated with =/X (not M): '
            'opcode={}, subject={}:{}, query={}:{}, align-index={}'
        ).format(err.op_index + 1, op, chrom[rec], err.pos_ref, qry[rec], err.pos_qry, align_index[rec]))
    raise RuntimeError((
        'Illegal operation code in CIGAR string at operation {}: '
        'opcode={}, subject={}:{} , query={}:{}, align-index={}'
    ).format(err.op_index + 1, op, chrom[rec], err.pos_ref, qry[rec], err.pos_qry, align_index[rec]))


def _raise_parse(perr, chrom, qry, pos):
    r = perr.rec
    if perr.code == 2:
        raise RuntimeError('Missing length in CIGAR string for contig {} alignment starting at {}:{}: CIGAR index {}'.format(
            qry[r], chrom[r], pos[r], perr.text_pos))
    if perr.code == 3:
        raise RuntimeError('Unknown CIGAR operation for contig {} alignment starting at {}:{}: CIGAR operation {}'.format(
            qry[r], chrom[r], pos[r], chr(perr.ch)))
    raise IndexError('string index out of range')


def make_insdel_snv_calls(df_align, ref_fa_name, tig_fa_name, hap, version_id=True):
    """
    Parse variants from CIGAR strings.

    :param df_align: Post-cut BED of read alignments (needs ``#CHROM POS INDEX QRY_ID REV CIGAR``).
    :param ref_fa_name: Reference FASTA file name.
    :param tig_fa_name: Contig FASTA file name.
    :param hap: String identifying the haplotype ("h1", "h2").
    :param version_id: Version duplicate variant IDs if `True`.

    :return: ``(df_snv, df_insdel)`` -- the order the reference actually returns (pavlib/cigarcall.py:362).
    """
    global last_stats
    n_rec = df_align.shape[0]
    if n_rec == 0:
        return _empty(SNV_COLUMNS), _empty(INSDEL_COLUMNS)

    chrom = df_align['#CHROM'].to_numpy(dtype=object)
    qry = df_align['QRY_ID'].to_numpy(dtype=object)
    rev = np.array([bool(x) for x in df_align['REV'].tolist()], dtype=bool)
    pos0 = np.array([int(x) for x in df_align['POS'].tolist()], dtype=np.int64)
    align_index = df_align['INDEX'].to_numpy(dtype=object)
    cigars = df_align['CIGAR'].tolist()

    ctx = device.get_context()

    # Sequences referenced by this table -> HBM (packed on the device)
    ref_names = _first_seen(str(c) for c in chrom.tolist())
    tig_names = _first_seen(str(q) for q in qry.tolist())
    ref_fa = fasta.open_fasta(ref_fa_name)
    tig_fa = fasta.open_fasta(tig_fa_name)
    ref_arr = [ref_fa.fetch_array(nm) for nm in ref_names]
    tig_arr = [tig_fa.fetch_array(nm) for nm in tig_names]
    ref_store = device.SeqStore(ctx, list(ref_names), ref_arr)
    tig_store = device.SeqStore(ctx, list(tig_names), tig_arr)
    try:
        ref_id = np.array([ref_names[str(c)] for c in chrom.tolist()], dtype=np.int32)
        qry_id = np.array([tig_names[str(q)] for q in qry.tolist()], dtype=np.int32)
        ops, op_off, perr = device.parse_cigars(cigars)
        snv, indel, cerr, stats = device.cigar_call(ctx, ref_store, tig_store, ref_id, qry_id, pos0.astype(np.int32),
                                                    rev.astype(np.uint8), ops, op_off)
    finally:
        ref_store.close()
        tig_store.close()
    last_stats = stats.as_dict()

    # Errors surface in walk order (the reference raises lazily while iterating records and ops)
    if cerr.code == 1 and (perr.code == 0 or (cerr.rec, cerr.op_index) < (perr.rec, perr.op_index)):
        _raise_illegal(cerr, chrom, qry, align_index)
    if perr.code != 0:
        _raise_parse(perr, chrom, qry, pos0)

    return build_frames(snv, indel, chrom, qry, rev, align_index, ref_arr, tig_arr, ref_id, qry_id, hap, version_id)


def build_frames(snv, indel, chrom, qry, rev, align_index, ref_arr, tig_arr, ref_id, qry_id, hap, version_id):
    """Rows from the device (``pavgpu_snv_row`` / ``pavgpu_indel_row`` arrays in emission order) -> the two
    DataFrames of the reference. Host-side string formatting only; every coordinate comes from the GPU."""
    n_rec = len(chrom)
    chrom_s = np.array([f'{c}' for c in chrom.tolist()], dtype=object)
    chrom_rank = {c: i for i, c in enumerate(sorted(set(chrom_s.tolist())))}
    chrom_code_rec = np.array([chrom_rank[c] for c in chrom_s.tolist()], dtype=np.int64)
    qry_s = np.array([f'{q}' for q in qry.tolist()], dtype=object)
    strand_rec = np.where(rev, '-', '+').astype(object)

    # ------------------------------------------------------------------ SNV rows (cigarcall.py:98-135)
    if len(snv):
        rec = snv['rec'].astype(np.int64)
        pos = snv['pos_ref'].astype(np.int64)
        qp = snv['qry_pos'].astype(np.int64)
        ref_b = np.empty(len(snv), dtype=np.uint8)
        alt_b = np.empty(len(snv), dtype=np.uint8)
        bounds = np.searchsorted(rec, np.arange(n_rec + 1))
        for r in np.flatnonzero(np.diff(bounds)).tolist():
            a, b = bounds[r], bounds[r + 1]
            ref_b[a:b] = ref_arr[ref_id[r]][pos[a:b]]
            t = tig_arr[qry_id[r]][qp[a:b]]
            alt_b[a:b] = fasta.COMPLEMENT[t] if rev[r] else t
        chrom_row = chrom_s[rec]
        ids = _join(chrom_row.astype(_S), '-', _istr(pos + 1), '-SNV-', _CHR[fasta.UPPER[ref_b]].astype(_S),
                    _CHR[fasta.UPPER[alt_b]].astype(_S))
        qp1 = _istr(qp + 1)
        cols = {
            '#CHROM': chrom[rec], 'POS': pos.astype(object), 'END': (pos + 1).astype(object), 'ID': ids,
            'SVTYPE': 'SNV', 'SVLEN': 1, 'REF': _CHR[ref_b], 'ALT': _CHR[alt_b], 'HAP': hap,
            'QRY_REGION': _join(qry_s[rec].astype(_S), ':', qp1, '-', qp1), 'QRY_STRAND': strand_rec[rec],
            'CI': 0, 'ALIGN_INDEX': align_index[rec], 'CALL_SOURCE': CALL_SOURCE,
        }
        if version_id:
            cols['ID'] = variant.version_id(pd.Series(cols['ID'], dtype=object)).to_numpy(dtype=object)
        order = _sort_order(chrom_code_rec[rec], pos, pos + 1, cols['ID'])
        df_snv = _frame(cols, SNV_COLUMNS, order)
    else:
        df_snv = _empty(SNV_COLUMNS)

    # ------------------------------------------------------------------ INS / DEL rows (cigarcall.py:141-282)
    if len(indel):
        rec = indel['rec'].astype(np.int64)
        pos = indel['pos'].astype(np.int64)
        end = indel['end'].astype(np.int64)
        svlen = indel['svlen'].astype(np.int64)
        is_del = indel['svtype'] == 1
        qp = indel['qry_pos'].astype(np.int64)
        qe = indel['qry_end'].astype(np.int64)
        svtype = np.where(is_del, 'DEL', 'INS').astype(object)
        seq = np.empty(len(indel), dtype=object)
        for i, (r, d, p, n, a, b) in enumerate(zip(rec.tolist(), is_del.tolist(), pos.tolist(), svlen.tolist(),
                                                   qp.tolist(), qe.tolist())):
            if d:
                seq[i] = ref_arr[ref_id[r]][p:p + n].tobytes().decode('ascii')
            else:
                s = tig_arr[qry_id[r]][a:b]
                seq[i] = (fasta.reverse_complement(s) if rev[r] else s).tobytes().decode('ascii')
        chrom_row = chrom_s[rec]
        ids = _join(chrom_row.astype(_S), '-', _istr(pos + 1), '-', svtype.astype(_S), '-', _istr(svlen))
        qry_region = _join(qry_s[rec].astype(_S), ':', _istr(qp + 1), '-', _istr(np.where(is_del, qp + 1, qe)))
        cols = {
            '#CHROM': chrom[rec], 'POS': pos.astype(object), 'END': end.astype(object), 'ID': ids, 'SVTYPE': svtype,
            'SVLEN': svlen.astype(object), 'HAP': hap, 'QRY_REGION': qry_region, 'QRY_STRAND': strand_rec[rec], 'CI': 0,
            'ALIGN_INDEX': align_index[rec], 'LEFT_SHIFT': indel['left_shift'].astype(np.int64).astype(object),
            'HOM_REF': _join(_istr(indel['hom_ref_l']), ',', _istr(indel['hom_ref_r'])),
            'HOM_TIG': _join(_istr(indel['hom_tig_l']), ',', _istr(indel['hom_tig_r'])),
            'CALL_SOURCE': CALL_SOURCE, 'SEQ': seq,
        }
        if version_id:
            cols['ID'] = variant.version_id(pd.Series(cols['ID'], dtype=object)).to_numpy(dtype=object)
        order = _sort_order(chrom_code_rec[rec], pos, end, cols['ID'])
        df_insdel = _frame(cols, INSDEL_COLUMNS, order)
    else:
        df_insdel = _empty(INSDEL_COLUMNS)

    return df_snv, df_insdel
