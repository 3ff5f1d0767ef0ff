"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/pav_oracle.c).

Nothing under pav_b200/ imports this module. It is used by tests/ (as the checker), by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.

Parity status: PINNED. tests/test_oracle_golden.py checks every function here against the golden
vectors in tests/golden/, which were produced by the unmodified reference (PAV 2.4.6.0) via
tests/golden/make_golden.py.

Restated reference functions (file:line under /root/reference):
  make_insdel_snv_calls   pavlib/cigarcall.py:24-362 (row assembly :98-135,:185-210,:254-279; sort :320,:343)
  version_id              dep/svpop/svpoplib/variant.py:664-752
  left_/right_homology    pavlib/call.py:542-647
  density_table           scripts/density.py:423-571 (+ get_smoothed_density :154-342)
  rl_encoder              pavlib/density.py:330-361
"""
import ctypes
import gzip
import os
import re
import subprocess

import numpy as np
import pandas as pd

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, '_build')
_LIB_PATH = os.path.join(_BUILD, 'libpavoracle.so')
_SRC = os.path.join(_HERE, 'pav_oracle.c')
_lib = None


def build(force=False):
    """Compile the C restatement with gcc (no GPU, no reference sources involved)."""
    os.makedirs(_BUILD, exist_ok=True)
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC)):
        return _LIB_PATH
    tmp = _LIB_PATH + f'.{os.getpid()}.tmp'
    subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-std=gnu11', '-ffp-contract=off', '-o', tmp, _SRC, '-lm'])
    os.replace(tmp, _LIB_PATH)
    return _LIB_PATH


class _Snv(ctypes.Structure):
    _fields_ = [('pos_ref', ctypes.c_int64), ('qry_pos', ctypes.c_int64), ('rec', ctypes.c_int32),
                ('ref_base', ctypes.c_uint8), ('alt_base', ctypes.c_uint8), ('pad', ctypes.c_uint8 * 2)]


SNV_DTYPE = np.dtype([('pos_ref', '<i8'), ('qry_pos', '<i8'), ('rec', '<i4'), ('ref_base', 'u1'),
                      ('alt_base', 'u1'), ('pad', 'u1', (2,))])
INDEL_DTYPE = np.dtype([('pos', '<i8'), ('end', '<i8'), ('svlen', '<i8'), ('qry_pos', '<i8'), ('qry_end', '<i8'),
                        ('left_shift', '<i8'), ('hom_ref_l', '<i8'), ('hom_ref_r', '<i8'), ('hom_tig_l', '<i8'),
                        ('hom_tig_r', '<i8'), ('seq_start', '<i8'), ('rec', '<i4'), ('svtype', '<i4')])


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, i32, vp, cp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_char_p
        L.orc_left_homology.restype = i64
        L.orc_left_homology.argtypes = [i64, cp, i64, cp, i64]
        L.orc_right_homology.restype = i64
        L.orc_right_homology.argtypes = [i64, cp, i64, cp, i64]
        L.orc_walk_new.restype = vp
        L.orc_walk_free.argtypes = [vp]
        L.orc_walk_n_snv.restype = i64
        L.orc_walk_n_snv.argtypes = [vp]
        L.orc_walk_n_indel.restype = i64
        L.orc_walk_n_indel.argtypes = [vp]
        L.orc_walk_snv.restype = vp
        L.orc_walk_snv.argtypes = [vp]
        L.orc_walk_indel.restype = vp
        L.orc_walk_indel.argtypes = [vp]
        L.orc_walk_error.argtypes = [vp] + [ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(i64),
                                            ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_int)]
        L.orc_walk_record.restype = ctypes.c_int
        L.orc_walk_record.argtypes = [vp, cp, i64, i64, cp, cp, i64, cp, cp, i64, ctypes.c_int, i32]
        L.orc_kmer_rc.restype = ctypes.c_uint64
        L.orc_kmer_rc.argtypes = [ctypes.c_uint64, ctypes.c_int]
        L.orc_kmer_stream.restype = i64
        L.orc_kmer_stream.argtypes = [cp, i64, ctypes.c_int, vp, vp]
        L.orc_density.restype = ctypes.c_int
        L.orc_density.argtypes = [cp, i64, cp, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.POINTER(vp)]
        L.orc_density_free.argtypes = [vp]
        for nm, rt in [('rows', i64), ('smoothed', ctypes.c_int), ('n_eval', i64), ('kmer', vp), ('index', vp),
                       ('state_mer', vp), ('state', vp)]:
            f = getattr(L, 'orc_density_' + nm)
            f.restype = rt
            f.argtypes = [vp]
        L.orc_density_kern.restype = vp
        L.orc_density_kern.argtypes = [vp, ctypes.c_int]
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------
# FASTA (independent of the product's reader)
# ---------------------------------------------------------------------------------------------

_FA_CACHE = {}


class _FastaDict:
    """``{name: bytes}`` view of a FASTA. With a samtools ``.fai`` next to a plain file only the requested
    records are read (like an indexed pysam fetch); otherwise the whole file is parsed once."""

    def __init__(self, path):
        self.path = path
        self.index = None
        self.cache = {}
        if not path.endswith('.gz') and os.path.exists(path + '.fai'):
            self.index = {}
            with open(path + '.fai') as fh:
                for line in fh:
                    t = line.rstrip('\n').split('\t')
                    self.index[t[0]] = (int(t[1]), int(t[2]), int(t[3]), int(t[4]))
        else:
            opener = gzip.open if path.endswith('.gz') else open
            with opener(path, 'rb') as fh:
                data = fh.read()
            for rec in data.split(b'>')[1:]:
                head, _, body = rec.partition(b'\n')
                self.cache[head.split()[0].decode()] = body.replace(b'\n', b'').replace(b'\r', b'')

    def __getitem__(self, name):
        if name in self.cache:
            return self.cache[name]
        if self.index is None or name not in self.index:
            raise KeyError(name)
        length, offset, linebases, linewidth = self.index[name]
        n_lines = (length + linebases - 1) // linebases if length else 0
        with open(self.path, 'rb') as fh:
            fh.seek(offset)
            raw = fh.read(n_lines * linewidth)
        seq = raw.replace(b'\n', b'').replace(b'\r', b'')[:length]
        if len(self.cache) > 8:
            self.cache.clear()
        self.cache[name] = seq
        return seq

    def __iter__(self):
        return iter(self.index if self.index is not None else self.cache)

    def keys(self):
        return list(iter(self))


def read_fasta(path):
    """FASTA as a ``{name: bytes}``-like object (original case). Plain (+ optional .fai) or gzip."""
    key = (path, os.path.getmtime(path))
    if key in _FA_CACHE:
        return _FA_CACHE[key]
    fa = _FastaDict(path)
    if len(_FA_CACHE) > 4:
        _FA_CACHE.clear()
    _FA_CACHE[key] = fa
    return fa


_COMP = bytes.maketrans(b'ACGTMRWSYKVHDBNacgtmrwsykvhdbn', b'TGCAKYWSRMBDHVNtgcakywsrmbdhvn')


def revcomp_bytes(b):
    return b.translate(_COMP)[::-1]


# ---------------------------------------------------------------------------------------------
# Path A
# ---------------------------------------------------------------------------------------------

def left_homology(pos_tig, seq_tig, seq_sv):
    if seq_sv is None or seq_tig is None:
        return 0
    a, b = seq_tig.encode(), seq_sv.encode()
    return int(lib().orc_left_homology(pos_tig, a, len(a), b, len(b)))


def right_homology(pos_tig, seq_tig, seq_sv):
    if seq_sv is None or seq_tig is None:
        return 0
    a, b = seq_tig.encode(), seq_sv.encode()
    return int(lib().orc_right_homology(pos_tig, a, len(a), b, len(b)))


def version_id(ids):
    """svpoplib/variant.py:664-752 on a list of str."""
    from collections import Counter
    cnt = Counter(ids)
    dup = {k for k, c in cnt.items() if c > 1}
    if not dup:
        return list(ids)
    used = set(ids) - dup
    out = list(ids)
    for i, name in enumerate(out):
        if name in dup:
            new = name
            if new in used:
                if re.match(r'.*\.\d+$', name):
                    stem, ver = name.rsplit('.', 1)
                    ver = int(ver) + 1
                else:
                    stem, ver = name, 1
                new = f'{stem}.{ver}'
                while new in used:
                    ver += 1
                    new = f'{stem}.{ver}'
            out[i] = new
            used.add(new)
    return out


SNV_COLS = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'REF', 'ALT', 'HAP', 'QRY_REGION', 'QRY_STRAND', 'CI',
            'ALIGN_INDEX', 'CALL_SOURCE']
INSDEL_COLS = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'HAP', 'QRY_REGION', 'QRY_STRAND', 'CI', 'ALIGN_INDEX',
               'LEFT_SHIFT', 'HOM_REF', 'HOM_TIG', 'CALL_SOURCE', 'SEQ']


class CigarError(RuntimeError):
    pass


def walk_rows(df_align, ref_fa_name, tig_fa_name):
    """C walk over all records -> (snv structured array, indel structured array, per-record context).

    Per-record context = (chrom, qry_id, is_rev, align_index, oriented contig bytes, reference bytes).
    Raises the reference's exceptions (cigarcall.py:289-307, align.py:308-318).
    """
    L = lib()
    ref_fa = read_fasta(ref_fa_name)
    tig_fa = read_fasta(tig_fa_name) if tig_fa_name != ref_fa_name else ref_fa
    w = L.orc_walk_new()
    ctx = []
    try:
        cur_ref = cur_ref_name = cur_ref_up = None
        cur_tig = cur_tig_name = cur_tig_rev = cur_tig_up = None
        for rec, (_, row) in enumerate(df_align.iterrows()):
            is_rev = bool(row['REV'])
            if cur_ref_name is None or row['#CHROM'] != cur_ref_name:
                cur_ref_name = row['#CHROM']
                cur_ref = ref_fa[str(cur_ref_name)]
                cur_ref_up = cur_ref.upper()
            if cur_tig_name is None or row['QRY_ID'] != cur_tig_name or is_rev != cur_tig_rev:
                cur_tig_name = row['QRY_ID']
                cur_tig = tig_fa[str(cur_tig_name)]
                if is_rev:
                    cur_tig = revcomp_bytes(cur_tig)
                cur_tig_rev = is_rev
                cur_tig_up = cur_tig.upper()
            cigar = row['CIGAR'].encode()
            rc = L.orc_walk_record(w, cigar, len(cigar), int(row['POS']), cur_ref, cur_ref_up, len(cur_ref),
                                   cur_tig, cur_tig_up, len(cur_tig), int(is_rev), rec)
            ctx.append((cur_ref_name, cur_tig_name, is_rev, row['INDEX'], cur_tig, cur_ref))
            if rc != 0:
                ci, op, pr, pt, tp, ch = (ctypes.c_int64(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64(),
                                          ctypes.c_int64(), ctypes.c_int())
                L.orc_walk_error(w, ci, op, pr, pt, tp, ch)
                if rc == 1:
                    if chr(op.value) == 'M':
                        raise CigarError((
                            'Illegal operation code in CIGAR string at operation {}: '
                            'Alignments must be generated with =/X (not M): '
                            'opcode={}, subject={}:{}, query={}:{}, align-index={}'
                        ).format(ci.value, chr(op.value), cur_ref_name, pr.value, cur_tig_name, pt.value, row['INDEX']))
                    raise CigarError((
                        'Illegal operation code in CIGAR string at operation {}: '
                        'opcode={}, subject={}:{} , query={}:{}, align-index={}'
                    ).format(ci.value, chr(op.value), cur_ref_name, pr.value, cur_tig_name, pt.value, row['INDEX']))
                if rc == 2:
                    raise CigarError('Missing length in CIGAR string for contig {} alignment starting at {}:{}: CIGAR index {}'.format(
                        row['QRY_ID'], row['#CHROM'], row['POS'], tp.value))
                if rc == 3:
                    raise CigarError('Unknown CIGAR operation for contig {} alignment starting at {}:{}: CIGAR operation {}'.format(
                        row['QRY_ID'], row['#CHROM'], row['POS'], chr(ch.value)))
                if rc == 4:
                    raise IndexError('string index out of range')
                raise MemoryError('oracle: out of memory')
        n_snv, n_indel = L.orc_walk_n_snv(w), L.orc_walk_n_indel(w)
        snv = np.zeros(n_snv, dtype=SNV_DTYPE)
        indel = np.zeros(n_indel, dtype=INDEL_DTYPE)
        if n_snv:
            ctypes.memmove(snv.ctypes.data, L.orc_walk_snv(w), n_snv * SNV_DTYPE.itemsize)
        if n_indel:
            ctypes.memmove(indel.ctypes.data, L.orc_walk_indel(w), n_indel * INDEL_DTYPE.itemsize)
    finally:
        L.orc_walk_free(w)
    return snv, indel, ctx


def _frame_like_reference(rows, columns):
    """Row container built the way the reference builds it (pavlib/cigarcall.py:114-135,188-210,315,338):
    one ``pd.Series`` per variant, then ``pd.concat(axis=1).T``. This is where the reference spends ~93 % of
    its time, so the CPU *baseline* legs of bench.py use this mode to stand in for the
    Python reference, which cannot travel to the GPU box; the parity tests use the fast mode."""
    series = [pd.Series(list(r), index=columns) for r in rows]
    return pd.concat(series, axis=1).T


def make_insdel_snv_calls(df_align, ref_fa_name, tig_fa_name, hap, version_id=True, reference_containers=False):
    """Oracle restatement of pavlib.cigarcall.make_insdel_snv_calls -> (df_snv, df_insdel).

    ``reference_containers=True`` assembles the frames with the reference's per-variant Series + concat
    (same result, reference-like cost); the default builds them from tuples (same result, fast checker)."""
    _vid = globals()['version_id']
    snv, indel, ctx = walk_rows(df_align, ref_fa_name, tig_fa_name)

    rows = []
    for r in snv.tolist():
        pos, qp, rec, rb, ab, _ = r
        chrom, tig, is_rev, ai, _, _ = ctx[rec]
        rb, ab = chr(rb), chr(ab)
        rows.append((chrom, pos, pos + 1, f'{chrom}-{pos + 1}-SNV-{rb.upper()}{ab.upper()}', 'SNV', 1, rb, ab, hap,
                     f'{tig}:{qp + 1}-{qp + 1}', '-' if is_rev else '+', 0, ai, 'CIGAR'))
    if rows:
        df_snv = _frame_like_reference(rows, SNV_COLS) if reference_containers else pd.DataFrame(rows, columns=SNV_COLS).astype(object)
        if version_id:
            df_snv['ID'] = pd.Series(_vid(list(df_snv['ID'])), index=df_snv.index, dtype=object)
        df_snv.sort_values(['#CHROM', 'POS', 'END', 'ID'], inplace=True)
    else:
        df_snv = pd.DataFrame([], columns=SNV_COLS)

    rows = []
    for r in indel.tolist():
        pos, end, svlen, qp, qe, ls, hrl, hrr, htl, htr, ss, rec, svtype = r
        chrom, tig, is_rev, ai, qseq, rseq = ctx[rec]
        if svtype == 0:
            seq = qseq[ss:ss + svlen].decode()
            rows.append((chrom, pos, end, f'{chrom}-{pos + 1}-INS-{svlen}', 'INS', svlen, hap, f'{tig}:{qp + 1}-{qe}',
                         '-' if is_rev else '+', 0, ai, ls, f'{hrl},{hrr}', f'{htl},{htr}', 'CIGAR', seq))
        else:
            seq = rseq[ss:ss + svlen].decode()
            rows.append((chrom, pos, end, f'{chrom}-{pos + 1}-DEL-{svlen}', 'DEL', svlen, hap, f'{tig}:{qp + 1}-{qp + 1}',
                         '-' if is_rev else '+', 0, ai, ls, f'{hrl},{hrr}', f'{htl},{htr}', 'CIGAR', seq))
    if rows:
        df_insdel = _frame_like_reference(rows, INSDEL_COLS) if reference_containers else pd.DataFrame(rows, columns=INSDEL_COLS).astype(object)
        if version_id:
            df_insdel['ID'] = pd.Series(_vid(list(df_insdel['ID'])), index=df_insdel.index, dtype=object)
        df_insdel.sort_values(['#CHROM', 'POS', 'END', 'ID'], inplace=True)
    else:
        df_insdel = pd.DataFrame([], columns=INSDEL_COLS)
    return df_snv, df_insdel


# ---------------------------------------------------------------------------------------------
# Path B
# ---------------------------------------------------------------------------------------------

def kmer_stream(seq, k):
    """kanapy.util.kmer.stream(seq, KmerUtil(k), index=True) -> (uint64 k-mers, int32 start index). k <= 32."""
    L = lib()
    b = seq if isinstance(seq, bytes) else seq.encode()
    n = L.orc_kmer_stream(b, len(b), k, None, None)
    km = np.zeros(n, dtype=np.uint64)
    ix = np.zeros(n, dtype=np.int32)
    if n:
        L.orc_kmer_stream(b, len(b), k, km.ctypes.data, ix.ctypes.data)
    return km, ix


def kmer_rc(kmer, k):
    return int(lib().orc_kmer_rc(int(kmer), k))


ERR_INV_FAIL = 125


def density_arrays(ref_seq, tig_seq, k=31, rev=False, min_inf=2000, smooth=1.0, min_state=20, srs=20, delta=0.005):
    """scripts/density.py on in-memory windows. Returns (returncode, dict of arrays or None)."""
    L = lib()
    rb = ref_seq if isinstance(ref_seq, bytes) else bytes(ref_seq)
    tb = tig_seq if isinstance(tig_seq, bytes) else bytes(tig_seq)
    h = ctypes.c_void_p()
    rc = L.orc_density(rb, len(rb), tb, len(tb), k, int(rev), min_inf, float(smooth), min_state, srs, float(delta),
                       ctypes.byref(h))
    if rc != 0:
        return rc, None
    try:
        n = L.orc_density_rows(h)
        sm = bool(L.orc_density_smoothed(h))

        def arr(ptr, dt):
            a = np.zeros(n, dtype=dt)
            if n:
                ctypes.memmove(a.ctypes.data, ptr, n * a.itemsize)
            return a
        out = {'KMER': arr(L.orc_density_kmer(h), np.uint64), 'INDEX': arr(L.orc_density_index(h), np.int32),
               'STATE_MER': arr(L.orc_density_state_mer(h), np.int8), 'STATE': arr(L.orc_density_state(h), np.int8),
               'smoothed': sm, 'n_eval': int(L.orc_density_n_eval(h))}
        if sm:
            for s, nm in enumerate(('KERN_FWD', 'KERN_FWDREV', 'KERN_REV')):
                out[nm] = arr(L.orc_density_kern(h, s), np.float64)
    finally:
        L.orc_density_free(h)
    return 0, out


def density_frame(out):
    """Arrays -> DataFrame shaped like scripts/density.py's result (:341-342, or the raw frame of :193-194)."""
    if out['smoothed']:
        df = pd.DataFrame({'INDEX': out['INDEX'].astype(np.int64), 'STATE_MER': out['STATE_MER'].astype(np.int64),
                           'STATE': out['STATE'].astype(np.int64), 'KERN_FWD': out['KERN_FWD'],
                           'KERN_FWDREV': out['KERN_FWDREV'], 'KERN_REV': out['KERN_REV'],
                           'KMER': out['KMER'].astype(object)})
        df['KMER'] = [int(x) for x in out['KMER']]
    else:
        df = pd.DataFrame({'KMER': [int(x) for x in out['KMER']], 'INDEX': out['INDEX'].astype(np.int64),
                           'STATE': out['STATE'].astype(np.int64), 'STATE_MER': out['STATE_MER'].astype(np.int64)})
    return df


def rl_encoder(df, state_col='STATE'):
    """pavlib/density.py:330-361."""
    state = None
    count = 0
    pos = end = None
    for st, ix in zip(df[state_col].tolist(), df['INDEX'].tolist()):
        if st == state:
            count += 1
            end = ix
        else:
            if state is not None:
                yield (state, count, pos, end)
            state, count, pos, end = st, 1, ix, ix
    if state is not None:
        yield (state, count, pos, end)
