"""Alignment-table helpers with the reference's interface and error text (pavlib/align/align.py).

* ``cigar_str_to_tuples`` (:286-322): the hot path does not call this generator (records are tokenised in bulk
  by ``pavgpu_cigar_parse``); it exists because other PAV code imports it from ``pavlib.align``.
* ``get_align_bed`` (:666-794), ``count_cigar`` (:534-663), ``check_record`` (:364-508), ``clip_soft_to_hard``
  (:797-831): SAM text -> the alignment table that feeds ``make_insdel_snv_calls``, without
  pysam/htslib. Per-record span counting reuses the bulk tokenizer of libpavgpu (host C) and numpy reductions.
"""
import gzip
import re

import numpy as np
import pandas as pd

from .lift import AlignLift  # noqa: F401  (reference: pavlib/align/__init__.py re-exports lift)

_DIGITS = frozenset('0123456789')
_OPS = frozenset('MIDNSHP=X')


def cigar_str_to_tuples(record):
    """Yield ``(length, op)`` for an alignment record (Series with CIGAR, QRY_ID, #CHROM, POS) or a CIGAR string."""
    cigar = record['CIGAR'] if isinstance(record, pd.Series) else record
    pos, n = 0, len(cigar)
    while pos < n:
        end = pos
        while cigar[end] in _DIGITS:  # IndexError past the end, like the reference
            end += 1
        if end == pos:
            raise RuntimeError('Missing length in CIGAR string for contig {} alignment starting at {}:{}: CIGAR index {}'.format(
                record['QRY_ID'], record['#CHROM'], record['POS'], pos))
        if cigar[end] not in _OPS:
            raise RuntimeError('Unknown CIGAR operation for contig {} alignment starting at {}:{}: CIGAR operation {}'.format(
                record['QRY_ID'], record['#CHROM'], record['POS'], cigar[pos]))
        yield int(cigar[pos:end]), cigar[end]
        pos = end + 1


CIGAR_M, CIGAR_I, CIGAR_D, CIGAR_N, CIGAR_S, CIGAR_H, CIGAR_P, CIGAR_EQ, CIGAR_X = range(9)
CIGAR_CODE_TO_CHAR = dict(enumerate('MIDNSHP=X'))
ALIGN_COLUMNS = ['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END', 'QRY_LEN', 'RG', 'AO', 'MAPQ', 'REV', 'FLAGS',
                 'HAP', 'CIGAR']


def _count_cigar_fast(cigar, allow_m):
    """``count_cigar`` for a well-formed record: the C tokenizer of libpavgpu + numpy sums. Returns ``None`` for anything irregular
    (syntax error, op outside =XID[M], clips not of the form [H][S] ... [S][H]); the caller then takes the op-by-op path, which raises
    the reference's exceptions."""
    from .. import device
    try:
        ops, _, perr = device.parse_cigars([cigar])
    except Exception:  # noqa: BLE001  (e.g. an op length beyond the packed format)
        return None
    if perr.code != 0 or len(ops) == 0:
        return None
    code, ln = (ops & 15).astype(np.int64), (ops >> 4).astype(np.int64)
    is_clip = (code == CIGAR_S) | (code == CIGAR_H)
    body = np.flatnonzero(~is_clip)
    if len(body) == 0:
        return None
    a, b = int(body[0]), int(body[-1]) + 1
    if a > 2 or len(code) - b > 2 or is_clip[a:b].any():
        return None
    lead, tail = code[:a].tolist(), code[b:].tolist()
    if lead not in ([], [CIGAR_H], [CIGAR_S], [CIGAR_H, CIGAR_S]) or tail not in ([], [CIGAR_S], [CIGAR_H], [CIGAR_S, CIGAR_H]):
        return None
    allowed = (CIGAR_EQ, CIGAR_X, CIGAR_I, CIGAR_D) + ((CIGAR_M,) if allow_m else ())
    cb, lb = code[a:b], ln[a:b]
    if not np.isin(cb, allowed).all():
        return None
    if (ln[:a] == 0).any() or (ln[b:] == 0).any():
        return None     # zero-length clips make the op-by-op checks ("duplicate", "before") behave differently: leave them to it
    ref_bp = int(lb[cb != CIGAR_I].sum())
    tig_bp = int(lb[cb != CIGAR_D].sum())
    clip = {CIGAR_H: 0, CIGAR_S: 0}
    clip_l, clip_r = dict(clip), dict(clip)
    for c, n_ in zip(lead, ln[:a].tolist()):
        clip_l[c] = n_
    for c, n_ in zip(tail, ln[b:].tolist()):
        clip_r[c] = n_
    return ref_bp, tig_bp, clip_l[CIGAR_H], clip_l[CIGAR_S], clip_r[CIGAR_H], clip_r[CIGAR_S]


def count_cigar(row, allow_m=False):
    """``(ref_bp, tig_bp, clip_h_l, clip_s_l, clip_h_r, clip_s_r)`` of an alignment record; same structural checks and
    messages as the reference (clips only at the ends, H outside S, no M unless ``allow_m``)."""
    fast = _count_cigar_fast(row['CIGAR'] if isinstance(row, pd.Series) else row, allow_m)
    if fast is not None:
        return fast
    ref_bp = tig_bp = clip_s_l = clip_h_l = clip_s_r = clip_h_r = 0
    ops = list(cigar_str_to_tuples(row))
    n, i = len(ops), 0
    while i < n and ops[i][1] in {'S', 'H'}:
        ln, op = ops[i]
        if op == 'S':
            if clip_s_l > 0:
                raise RuntimeError('Duplicate S records (left) at index {}'.format(i))
            clip_s_l = ln
        if op == 'H':
            if clip_h_l > 0:
                raise RuntimeError('Duplicate H records (left) at index {}'.format(i))
            if clip_s_l > 0:
                raise RuntimeError('S record before H (left) at index {}'.format(i))
            clip_h_l = ln
        i += 1
    while i < n:
        ln, op = ops[i]
        if op in {'=', 'X', 'I', 'D'} or op == 'M':
            if op == 'M' and not allow_m:
                raise RuntimeError('CIGAR op "M" is not allowed')
            if clip_s_r > 0 or clip_h_r > 0:
                raise RuntimeError('Found clipped bases before last non-clipped CIGAR operation at operation {} ({}{})'.format(i, ln, op))
            if op != 'I':
                ref_bp += ln
            if op != 'D':
                tig_bp += ln
        elif op == 'S':
            if clip_s_r > 0:
                raise RuntimeError('Duplicate S records (right) at operation {}'.format(i))
            if clip_h_r > 0:
                raise RuntimeError('H record before S record (right) at operation {}'.format(i))
            clip_s_r = ln
        elif op == 'H':
            if clip_h_r > 0:
                raise RuntimeError('Duplicate H records (right) at operation {}'.format(i))
            clip_h_r = ln
        else:
            raise RuntimeError('Bad CIGAR op: ' + op)
        i += 1
    return ref_bp, tig_bp, clip_h_l, clip_s_l, clip_h_r, clip_s_r


def _where(row):
    return '(INDEX={}, QRY={}:{}-{}, REF={}:{}-{})'.format(row['INDEX'], row['QRY_ID'], row['QRY_POS'], row['QRY_END'], row['#CHROM'],
                                                            row['POS'], row['END'])


def check_record(row, df_tig_fai):
    """Sanity checks of one alignment-table record (reference: pavlib/align/align.py:364-508); raises ``RuntimeError``."""
    try:
        ref_bp, tig_bp, _, _, _, _ = count_cigar(row)
    except Exception as ex:
        raise RuntimeError('CIGAR parsing error: {} {}'.format(ex, _where(row)))
    tig_len = df_tig_fai[row['QRY_ID']]
    if row['QRY_LEN'] != tig_len:
        raise RuntimeError('QRY_LEN != length from FAI ({} != {}) {}'.format(row['QRY_LEN'], tig_len, _where(row)))
    if row['QRY_POS'] >= row['QRY_END']:
        raise RuntimeError('QRY_POS >= QRY_END ({} >= {}) {}'.format(row['QRY_POS'], row['QRY_END'], _where(row)))
    if row['POS'] >= row['END']:
        raise RuntimeError('POS >= END ({} >= {}) {}'.format(row['POS'], row['END'], _where(row)))
    if row['POS'] < 0:
        raise RuntimeError('POS ({}) < 0 {}'.format(row['POS'], _where(row)))
    if row['QRY_POS'] < 0:
        raise RuntimeError('QRY_POS ({}) < 0 {}'.format(row['QRY_POS'], _where(row)))
    if row['POS'] + ref_bp != row['END']:
        raise RuntimeError('END mismatch: POS + ref_bp != END ({} != {}) {}'.format(row['POS'] + ref_bp, row['END'], _where(row)))
    if row['QRY_POS'] + tig_bp != row['QRY_END']:
        raise RuntimeError('QRY_POS + tig_bp != QRY_END: {} != {} {}'.format(row['QRY_POS'] + tig_bp, row['QRY_END'], _where(row)))
    if row['QRY_END'] > tig_len:
        raise RuntimeError('QRY_END > tig_len ({} > {}) {}'.format(row['QRY_END'], tig_len, _where(row)))


def clip_soft_to_hard(cigar_tuples):
    """Merge leading / trailing S and H ops of ``[(op code, length)]`` into single H ops (reference: :797-831)."""
    front_n = 0
    while len(cigar_tuples) > 0 and cigar_tuples[0][0] in {CIGAR_H, CIGAR_S}:
        front_n += cigar_tuples[0][1]
        cigar_tuples = cigar_tuples[1:]
    back_n = 0
    while len(cigar_tuples) > 0 and cigar_tuples[-1][0] in {CIGAR_H, CIGAR_S}:
        back_n += cigar_tuples[-1][1]
        cigar_tuples = cigar_tuples[:-1]
    if len(cigar_tuples) == 0:
        if front_n + back_n == 0:
            raise RuntimeError('Cannot convert soft clipping to hard: No CIGAR records')
        return [(front_n + back_n, CIGAR_H)]   # (sic) the reference swaps the tuple order in this corner
    if front_n > 0:
        cigar_tuples = [(CIGAR_H, front_n)] + cigar_tuples
    if back_n > 0:
        cigar_tuples = cigar_tuples + [(CIGAR_H, back_n)]
    return cigar_tuples


_LEAD_CLIPS = re.compile(r'^(?:\d+[SH])*')


def _n_digits(n):
    """Decimal digits of non-negative int64 lengths (< 2^28)."""
    return 1 + (n >= 10).astype(np.int64) + (n >= 100) + (n >= 1000) + (n >= 10_000) + (n >= 100_000) + (n >= 1_000_000) + (n >= 10_000_000) + (n >= 100_000_000)


def _core_text(cigar, core_c, core_n, all_n):
    """CIGAR text of the ops between the clips, as ``str(length) + op`` per op. That is a substring of the SAM field unless a length
    was written with leading zeros, so the common case costs two regex matches instead of one f-string per op."""
    if int(_n_digits(all_n).sum()) + len(all_n) == len(cigar):    # every length is written without leading zeros
        end = len(cigar)
        while end > 0 and cigar[end - 1] in 'SH':      # strip trailing clip ops from the back (a regex anchored at '$' rescans the field)
            k = end - 1
            while k > 0 and cigar[k - 1] in _DIGITS:
                k -= 1
            if k == end - 1:
                break
            end = k
        return cigar[_LEAD_CLIPS.match(cigar).end():end]
    return ''.join(f'{a}{CIGAR_CODE_TO_CHAR[b]}' for b, a in zip(core_c.tolist(), core_n.tolist()))


def _record_stats_host(code, ln, op_off):
    """The per-record sums of ``pavgpu_cigar_record_stats`` with numpy on the host (same fields)."""
    from .. import _capi
    out = np.zeros(len(op_off) - 1, dtype=_capi.CIGAR_REC_STATS)
    for i in range(len(op_off) - 1):
        c, n = code[op_off[i]:op_off[i + 1]], ln[op_off[i]:op_off[i + 1]]
        is_clip = (c == CIGAR_S) | (c == CIGAR_H)
        body = np.flatnonzero(~is_clip)
        s = out[i]
        s['n_ops'] = len(c)
        s['flags'] = int((c == CIGAR_M).any())
        if len(body) == 0:
            s['lead'], s['first_body'], s['last_body'] = int(n.sum()), -1, -1
        else:
            a, b = int(body[0]), int(body[-1])
            s['lead'], s['trail'], s['first_body'], s['last_body'] = int(n[:a].sum()), int(n[b + 1:].sum()), a, b
            cc, cn = c[a:b + 1], n[a:b + 1]
            eqx = (cc == CIGAR_EQ) | (cc == CIGAR_X)
            s['ref_bp'] = int(cn[eqx | (cc == CIGAR_D) | (cc == CIGAR_N)].sum())
            s['qry_bp'] = int(cn[eqx | (cc == CIGAR_I)].sum())
            if is_clip[a:b + 1].any():
                s['flags'] |= 2
        if len(c):
            s['clip_h_first'] = int(n[0]) if c[0] == CIGAR_H else 0
            noth = np.flatnonzero(c != CIGAR_H)
            if len(noth) and c[noth[0]] == CIGAR_S:
                s['lead_s'] = int(n[noth[0]])
    return out


def get_align_bed(align_file, df_tig_fai, hap, min_mapq=0, device_stats=False):
    """
    Read a SAM text file (plain or gzip) as the alignment table PAV processes (reference: pavlib/align/align.py:666-794,
    which reads SAM/BAM/CRAM through pysam). Unmapped records, records below ``min_mapq`` and records without a CIGAR are
    dropped; soft clips become hard clips; ``M`` operations are rejected; the table is sorted by
    ``#CHROM, POS, END (descending), QRY_ID`` and every record is sanity-checked.

    ``device_stats=True``: the per-record CIGAR sums (aligned spans, clips, M test) come from ``pavgpu_cigar_record_stats``
    (one warp per record on the GPU) instead of the numpy pass on the host; same table.
    """
    from .. import device
    recs = []
    opener = gzip.open if str(align_file).endswith('.gz') else open
    align_index = -1
    with opener(align_file, 'rt') as fh:
        for line in fh:
            if line.startswith('@') or not line.strip():
                continue
            align_index += 1
            tok = line.rstrip('\n').split('\t')
            flag, mapq, cigar = int(tok[1]), int(tok[4]), tok[5]
            if (flag & 0x4) or mapq < min_mapq or cigar == '*' or cigar == '':
                continue
            tags = {}
            for t in tok[11:]:
                k, ty, v = t.split(':', 2)
                if k in ('RG', 'AO'):
                    tags[k] = int(v) if ty == 'i' else v
            recs.append((align_index, tok[0], flag, tok[2], int(tok[3]) - 1, mapq, cigar, tags))
    if not recs:
        return pd.DataFrame([], columns=ALIGN_COLUMNS)

    # bulk tokenise (host C) and reduce per record
    ops, op_off, perr = device.parse_cigars([r[6] for r in recs])
    if perr.code != 0:
        raise RuntimeError('Malformed CIGAR in SAM record {} ({})'.format(recs[perr.rec][0], recs[perr.rec][1]))
    code = (ops & 15).astype(np.int64)
    ln = (ops >> 4).astype(np.int64)
    stats = device.cigar_record_stats(ops, op_off) if device_stats else _record_stats_host(code, ln, op_off)
    rows = []
    for i, (idx, qname, flag, rname, pos, mapq, _, tags) in enumerate(recs):
        c, n = code[op_off[i]:op_off[i + 1]], ln[op_off[i]:op_off[i + 1]]
        st = stats[i]
        if st['flags'] & 1:
            raise RuntimeError(('Found alignment match CIGAR operation (M) for record {} (Start = {}:{}): '
                                'Alignment requires CIGAR base-level match/mismatch (=X)').format(qname, rname, pos))
        lead, trail, ref_bp, qry_bp = int(st['lead']), int(st['trail']), int(st['ref_bp']), int(st['qry_bp'])
        if st['first_body'] < 0:
            core_c, core_n = c[:0], n[:0]
        else:
            core_c, core_n = c[st['first_body']:st['last_body'] + 1], n[st['first_body']:st['last_body'] + 1]
        # pysam: query_alignment_start counts leading soft clips only; hard clips are added back by the reference
        clip_h, lead_s = int(st['clip_h_first']), int(st['lead_s'])
        tig_map_pos = lead if lead > 0 else 0
        if lead_s + clip_h != tig_map_pos:
            raise RuntimeError(f'First aligned based from pysam ({lead_s}) does not match clipping ({tig_map_pos}) at alignment record {idx}')
        tig_map_end = tig_map_pos + qry_bp
        parts = ([f'{lead}H'] if lead > 0 else []) + [_core_text(recs[i][6], core_c, core_n, n)] + ([f'{trail}H'] if trail > 0 else [])
        tig_len = df_tig_fai[qname]
        rev = bool(flag & 0x10)
        rows.append((rname, pos, pos + ref_bp, idx, qname, tig_len - tig_map_end if rev else tig_map_pos,
                     tig_len - tig_map_pos if rev else tig_map_end, tig_len, tags.get('RG', 'NA'), tags.get('AO', 'NA'), mapq, rev,
                     f'0x{flag:04x}', hap, ''.join(parts)))
    df = pd.DataFrame({col: pd.Series([r[j] for r in rows], dtype=object) for j, col in enumerate(ALIGN_COLUMNS)}, columns=ALIGN_COLUMNS)
    df.sort_values(['#CHROM', 'POS', 'END', 'QRY_ID'], ascending=[True, True, False, True], inplace=True)
    df.apply(check_record, df_tig_fai=df_tig_fai, axis=1)
    return df
