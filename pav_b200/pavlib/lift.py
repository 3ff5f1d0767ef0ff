"""Reference <-> contig coordinate lift through alignment records.

Same interface and corner-case behaviour as the reference's ``pavlib.align.AlignLift``
(pavlib/align/lift.py:12-487) -- including its quirks (a 1-bp aligned block lifts to the block END,
reverse records lift ``len - pos``) -- but each record's CIGAR becomes two sorted block tables
(prefix sums over the packed ops + ``numpy.searchsorted``) instead of two Python interval trees.
SURVEY 8(f) rank 2: needed by ``scan_for_inv`` on boxes where the reference package is absent.
"""
import numpy as np

from .. import device
from . import seq as pavseq

_MATCH_CODES = (0, 7, 8)  # M = X


class _RecordMap:
    __slots__ = ('r_begin', 'r_end', 'r_d0', 'r_d1', 'q_begin', 'q_end', 'q_d0', 'q_d1')


class AlignLift:
    def __init__(self, df, df_fai, cache_align=10):
        """:param df: alignment table (#CHROM POS END QRY_ID QRY_POS QRY_END REV INDEX CIGAR), unique index.
        :param df_fai: Series of contig lengths keyed by contig name. :param cache_align: records to keep mapped."""
        self.df = df
        self.df_fai = df_fai
        self.cache_align = cache_align
        if len(set(df.index)) != df.shape[0]:
            raise RuntimeError('Cannot create AlignLift object with duplicate index values')
        self._index = np.asarray(df.index)
        self._chrom = df['#CHROM'].to_numpy(dtype=object)
        self._pos = df['POS'].to_numpy(dtype=np.int64)
        self._end = df['END'].to_numpy(dtype=np.int64)
        self._qid = df['QRY_ID'].to_numpy(dtype=object)
        self._qpos = df['QRY_POS'].to_numpy(dtype=np.int64)
        self._qend = df['QRY_END'].to_numpy(dtype=np.int64)
        # per-record values handed out with every lift, read once (a ``df.iloc[i]`` per lookup costs more than the lift itself)
        self._rev = df['REV'].to_numpy()
        self._aln_index = df['INDEX'].to_numpy()
        self._cigar = df['CIGAR'].to_numpy(dtype=object)
        self._by_chrom, self._by_qid = {}, {}
        for i, (c, q) in enumerate(zip(self._chrom.tolist(), self._qid.tolist())):
            self._by_chrom.setdefault(c, []).append(i)
            self._by_qid.setdefault(q, []).append(i)
        self._by_chrom = {c: np.array(v, dtype=np.int64) for c, v in self._by_chrom.items()}
        self._by_qid = {q: np.array(v, dtype=np.int64) for q, v in self._by_qid.items()}
        # scalar copies for the common case of one candidate record (a numpy mask over a 1-element array costs ~10 us per lookup)
        self._pos_l, self._end_l = self._pos.tolist(), self._end.tolist()
        self._qpos_l, self._qend_l = self._qpos.tolist(), self._qend.tolist()
        # Block tables of the records looked at so far. The reference keeps the interval trees of the last ``cache_align`` (10)
        # records; a batch of hundreds of loci on as many records rebuilt its tables on most lookups that way (a fifth of the host
        # time of call_inv_batch). The tables are a cache either way, so they are kept until they add up to PAVGPU_LIFT_CACHE_MB
        # (default 256), oldest out first, and never fewer than ``cache_align`` of them.
        import collections
        import os
        self._cache = collections.OrderedDict()
        self._cache_bytes = 0
        self._cache_cap = int(os.environ.get('PAVGPU_LIFT_CACHE_MB', '256')) << 20
        self._dev_index = None      # device.LiftIndex, built by the first batched lift

    # ------------------------------------------------------------------ per-record block tables
    def _record_map(self, i):
        if i in self._cache:
            self._cache.move_to_end(i)
            return self._cache[i]
        row = {'#CHROM': self._chrom[i], 'POS': self._pos[i], 'QRY_ID': self._qid[i]}
        ops, _, perr = device.parse_cigars([self._cigar[i]])
        if perr.code != 0:
            raise RuntimeError('Malformed CIGAR for alignment {}:{} ({})'.format(row['#CHROM'], row['POS'], row['QRY_ID']))
        code = (ops & 15).astype(np.int64)
        ln = (ops >> 4).astype(np.int64)
        is_m = (code == 0) | (code == 7) | (code == 8)
        bad = ~(is_m | (code == 1) | (code == 2) | (code == 4) | (code == 5))
        if bad.any():
            raise RuntimeError('Unhandled CIGAR operation: {}: Alignment {}:{} ({})'.format(
                'MIDNSHP=X'[int(code[bad][0])], row['#CHROM'], row['POS'], row['QRY_ID']))
        ref_adv = np.where(is_m | (code == 2), ln, 0)
        qry_adv = np.where(is_m | (code == 1) | (code == 4) | (code == 5), ln, 0)
        sub = int(row['POS']) + np.cumsum(ref_adv) - ref_adv
        qry = np.cumsum(qry_adv) - qry_adv
        m = _RecordMap()
        rsel = is_m | (code == 2)
        m.r_begin, m.r_end = sub[rsel], sub[rsel] + ln[rsel]
        m.r_d0 = qry[rsel]
        m.r_d1 = np.where(is_m[rsel], qry[rsel] + ln[rsel], qry[rsel] + 1)
        qsel = is_m | (code == 1)
        m.q_begin, m.q_end = qry[qsel], qry[qsel] + ln[qsel]
        m.q_d0 = sub[qsel]
        m.q_d1 = np.where(is_m[qsel], sub[qsel] + ln[qsel], sub[qsel] + 1)
        m_bytes = sum(getattr(m, f).nbytes for f in _RecordMap.__slots__)
        while len(self._cache) >= max(int(self.cache_align), 1) and self._cache_bytes + m_bytes > self._cache_cap:
            _, old = self._cache.popitem(last=False)
            self._cache_bytes -= sum(getattr(old, f).nbytes for f in _RecordMap.__slots__)
        self._cache[i] = m
        self._cache_bytes += m_bytes
        return m

    @staticmethod
    def _block(begin, end, p):
        """Index of the block with begin <= p < end, or -1."""
        k = int(np.searchsorted(begin, p, side='right')) - 1
        # zero-length blocks share a begin; walk back to a block that really contains p
        while k >= 0 and not (begin[k] <= p < end[k]):
            if begin[k] < p and end[k] <= p:
                return -1
            k -= 1
        return k

    # ------------------------------------------------------------------ batched point lifts on the device
    def _device_index(self):
        """``device.LiftIndex`` over every record of the table, built on first use; ``None`` when a record has an op the lift does not
        handle (the per-point path then raises for that record when -- and only when -- it is looked at, like the reference)."""
        if self._dev_index is None:
            self._dev_index = False
            try:
                ops, op_off, perr = device.parse_cigars(self._cigar.tolist())
                if perr.code == 0:
                    qlen = np.array([int(self.df_fai[q]) for q in self._qid.tolist()], dtype=np.int64)
                    idx = device.LiftIndex(device.get_context(), ops, op_off, self._pos, np.asarray(self._rev, dtype=bool).astype(np.uint8), qlen)
                    if idx.bad_rec < 0:
                        self._dev_index = idx
            except Exception:  # noqa: BLE001  (a contig missing from the .fai, a malformed CIGAR, no device: the per-point path reports what is
                pass           #  wrong for the record that is actually looked at, like the reference; the lift is host logic either way)
        return self._dev_index or None

    def lift_points(self, ids, coords, to_qry, gap=False):
        """Many point lifts at once: ``ids[i]`` = chromosome (``to_qry``) or contig name, ``coords[i]`` = position. Same results, in
        order, as ``lift_to_qry(id, coord)`` / ``lift_to_sub(id, coord, gap)`` one at a time -- the tuples, ``None`` where no single
        record covers the position, ``RuntimeError`` for the first position inside a record that no block covers -- with the block
        search of all positions in one launch (``pavgpu_lift_points``)."""
        idx = self._device_index()
        one = self.lift_to_qry if to_qry else (lambda i, c: self.lift_to_sub(i, c, gap))
        if idx is None:
            return [one(i, c) for i, c in zip(ids, coords)]
        by = self._by_chrom if to_qry else self._by_qid
        lo_l, hi_l = (self._pos_l, self._end_l) if to_qry else (self._qpos_l, self._qend_l)
        recs, sel = [], []
        out = [None] * len(ids)
        for n, (name, pos) in enumerate(zip(ids, coords)):
            cand = by.get(name)
            hit = [] if cand is None else [i for i in cand.tolist() if lo_l[i] <= pos < hi_l[i]]
            if len(hit) == 1:
                recs.append(hit[0])
                sel.append(n)
            elif len(hit) == 0 and gap and not to_qry:
                out[n] = self._get_subject_gap(name, pos)
        lifted, status = idx.lift(recs, [coords[n] for n in sel], to_qry)
        for n, i, v, st in zip(sel, recs, lifted.tolist(), status.tolist()):
            if st != 0:
                one(ids[n], coords[n])      # raises the reference's RuntimeError with its text
                raise RuntimeError('lift: device and host disagree on {}:{}'.format(ids[n], coords[n]))
            out[n] = ((self._qid[i] if to_qry else self._chrom[i]), v, self._rev[i], v, v, (self._aln_index[i],))
        return out

    def lift_regions_to_qry(self, regions):
        """``lift_region_to_qry`` for a list of regions (both ends of all of them in one device call)."""
        ids = [r.chrom for r in regions for _ in (0, 1)]
        coords = [c for r in regions for c in (r.pos, r.end)]
        pts = self.lift_points(ids, coords, True)
        out = []
        for k, _ in enumerate(regions):
            q_pos, q_end = pts[2 * k], pts[2 * k + 1]
            if q_pos is None or q_end is None or q_pos[0] != q_end[0] or q_pos[2] != q_end[2]:
                out.append(None)
                continue
            out.append(pavseq.Region(q_pos[0], q_pos[1], q_end[1], is_rev=q_pos[2], pos_min=q_pos[3], pos_max=q_pos[4],
                                     end_min=q_end[3], end_max=q_end[4], pos_aln_index=(q_pos[5],), end_aln_index=(q_end[5],)))
        return out

    # ------------------------------------------------------------------ point lifts
    def lift_to_qry(self, subject_id, coord):
        """Reference position(s) -> ``(QRY_ID, pos, is_rev, min, max, (INDEX,))`` or ``None``."""
        ret_list = issubclass(coord.__class__, (list, tuple))
        if not ret_list:
            coord = (coord,)
        out = []
        for pos in coord:
            cand = self._by_chrom.get(subject_id)
            if cand is None:
                hit = ()
            elif len(cand) == 1:
                i0 = int(cand[0])
                hit = (i0,) if self._pos_l[i0] <= pos < self._end_l[i0] else ()
            else:
                hit = cand[(self._pos[cand] <= pos) & (self._end[cand] > pos)]
            if len(hit) != 1:
                out.append(None)
                continue
            i = int(hit[0])
            m = self._record_map(i)
            row = {'#CHROM': self._chrom[i], 'QRY_ID': self._qid[i], 'REV': self._rev[i], 'INDEX': self._aln_index[i]}
            k = self._block(m.r_begin, m.r_end, pos)
            if k < 0:
                raise RuntimeError(('Program bug: Found no matches in a lift-tree for a record withing a '
                                    'global to-query tree: {}:{} (index={})').format(subject_id, pos, self._index[i]))
            if m.r_d1[k] - m.r_d0[k] > 1:
                qry_pos = int(m.r_d0[k] + (pos - m.r_begin[k]))
            else:
                qry_pos = int(m.r_d1[k])
            if row['REV']:
                qry_pos = self.df_fai[row['QRY_ID']] - qry_pos
            out.append((row['QRY_ID'], qry_pos, row['REV'], qry_pos, qry_pos, (row['INDEX'],)))
        return out if ret_list else out[0]

    def lift_to_sub(self, query_id, coord, gap=False):
        """Contig position(s) -> ``(#CHROM, pos, is_rev, min, max, (INDEX,))`` or ``None``."""
        ret_list = issubclass(coord.__class__, (list, tuple))
        if not ret_list:
            coord = (coord,)
        out = []
        for pos in coord:
            pos_org = pos
            cand = self._by_qid.get(query_id)
            if cand is None:
                hit = ()
            elif len(cand) == 1:
                i0 = int(cand[0])
                hit = (i0,) if self._qpos_l[i0] <= pos < self._qend_l[i0] else ()
            else:
                hit = cand[(self._qpos[cand] <= pos) & (self._qend[cand] > pos)]
            if len(hit) == 0 and gap:
                out.append(self._get_subject_gap(query_id, pos))
                continue
            if len(hit) != 1:
                out.append(None)
                continue
            i = int(hit[0])
            m = self._record_map(i)
            row = {'#CHROM': self._chrom[i], 'QRY_ID': self._qid[i], 'REV': self._rev[i], 'INDEX': self._aln_index[i]}
            if row['REV']:
                pos = self.df_fai[query_id] - pos
            k = self._block(m.q_begin, m.q_end, pos)
            if k < 0:
                k = self._block(m.q_begin, m.q_end, pos - 1)
                if k < 0 or m.q_end[k] != pos:
                    raise RuntimeError(('Found no matches in a lift-tree for a record within a '
                                        'global to-subject tree: {}:{} (index={}, gap={})').format(query_id, pos_org, self._index[i], gap))
            if m.q_d1[k] - m.q_d0[k] > 1:
                lift_pos = int(m.q_d0[k] + (pos - m.q_begin[k]))
            else:
                lift_pos = int(m.q_d1[k])
            out.append((row['#CHROM'], lift_pos, row['REV'], lift_pos, lift_pos, (row['INDEX'],)))
        return out if ret_list else out[0]

    def _get_subject_gap(self, query_id, pos):
        """Midpoint between the two alignment records flanking an unaligned contig position."""
        if pos is None:
            return None
        sub = self.df.loc[self.df['QRY_ID'] == query_id]
        left = sub.loc[sub['QRY_END'] < pos]
        right = sub.loc[sub['QRY_POS'] > pos]
        if left.shape[0] == 0 or right.shape[0] == 0:
            return None
        row_l = sub.loc[left['QRY_END'].sort_values().index[-1]]
        row_r = sub.loc[right['QRY_POS'].sort_values().index[0]]
        if row_l['#CHROM'] != row_r['#CHROM']:
            return None
        return (row_l['#CHROM'], int((row_l['QRY_END'] + row_r['QRY_POS']) / 2),
                row_l['REV'] if row_l['REV'] == row_r['REV'] else None, row_l['QRY_END'], row_r['QRY_POS'],
                (row_l['INDEX'], row_r['INDEX']))

    # ------------------------------------------------------------------ region lifts
    def lift_region_to_sub(self, region, gap=False):
        sub_pos, sub_end = self.lift_to_sub(region.chrom, (region.pos, region.end), gap)
        if sub_pos is None or sub_end is None:
            return None
        if sub_pos[0] != sub_end[0] or (sub_pos[2] is not None and sub_end[2] is not None and sub_pos[2] != sub_end[2]):
            return None
        return pavseq.Region(sub_pos[0], sub_pos[1], sub_end[1], is_rev=False, pos_min=sub_pos[3], pos_max=sub_pos[4],
                             end_min=sub_end[3], end_max=sub_end[4], pos_aln_index=(sub_pos[5],), end_aln_index=(sub_end[5],))

    def lift_region_to_qry(self, region):
        q_pos, q_end = self.lift_to_qry(region.chrom, (region.pos, region.end))
        if q_pos is None or q_end is None:
            return None
        if q_pos[0] != q_end[0] or q_pos[2] != q_end[2]:
            return None
        return pavseq.Region(q_pos[0], q_pos[1], q_end[1], is_rev=q_pos[2], pos_min=q_pos[3], pos_max=q_pos[4],
                             end_min=q_end[3], end_max=q_end[4], pos_aln_index=(q_pos[5],), end_aln_index=(q_end[5],))
