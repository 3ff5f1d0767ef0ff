"""Test-only stand-in for pysam (absent from this image), used ONLY to run the unmodified
reference from /root/reference when generating golden fixtures (tests/golden/make_golden.py).

Implements the subset the reference's hot path touches:
pysam.FastaFile(fn) as a context manager with fetch(name, start=None, end=None)
(pavlib/cigarcall.py:59-66, pavlib/seq.py:339-351). A plain FASTA with a .fai next to it is read by offset like htslib's
faidx does (one seek + one read per fetch: a worker of the CPU baseline fetching one chromosome of a 3.1 Gbp reference must not
parse the whole file); anything else (gzip, no index) is loaded whole.
"""
import gzip
import os

_CACHE = {}


class _Indexed:
    """name -> (length, offset, line_bases, line_width) from the .fai; fetch = seek + read + newline removal."""

    def __init__(self, fn):
        self.fn = fn
        self.idx = {}
        with open(fn + '.fai') as fh:
            for line in fh:
                t = line.rstrip('\n').split('\t')
                if len(t) >= 5:
                    self.idx[t[0]] = (int(t[1]), int(t[2]), int(t[3]), int(t[4]))

    def keys(self):
        return self.idx.keys()

    def length(self, name):
        return self.idx[name][0]

    def fetch(self, name, start, end):
        length, off, lb, lw = self.idx[name]
        start = 0 if start is None else max(0, start)
        end = length if end is None else min(length, end)
        if end <= start:
            return ''
        b0 = off + (start // lb) * lw + start % lb
        b1 = off + ((end - 1) // lb) * lw + (end - 1) % lb + 1
        with open(self.fn, 'rb') as fh:
            fh.seek(b0)
            raw = fh.read(b1 - b0)
        return raw.replace(b'\n', b'').replace(b'\r', b'').decode('ascii')


def _load(fn):
    if fn in _CACHE:
        return _CACHE[fn]
    if not str(fn).endswith('.gz') and os.path.exists(str(fn) + '.fai'):
        _CACHE[fn] = _Indexed(str(fn))
        return _CACHE[fn]
    opener = gzip.open if str(fn).endswith('.gz') else open
    seqs, name, chunks = {}, None, []
    with opener(fn, 'rt') as fh:
        for line in fh:
            if line.startswith('>'):
                if name is not None:
                    seqs[name] = ''.join(chunks)
                name = line[1:].split()[0]
                chunks = []
            else:
                chunks.append(line.strip())
    if name is not None:
        seqs[name] = ''.join(chunks)
    _CACHE[fn] = seqs
    return seqs


class FastaFile:
    def __init__(self, filename):
        self.filename = filename
        self._seqs = _load(filename)
        self.references = list(self._seqs.keys())
        if isinstance(self._seqs, _Indexed):
            self.lengths = [self._seqs.length(n) for n in self.references]
        else:
            self.lengths = [len(v) for v in self._seqs.values()]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def fetch(self, reference=None, start=None, end=None):
        if isinstance(self._seqs, _Indexed):
            return self._seqs.fetch(reference, start, end)
        seq = self._seqs[reference]
        if start is None and end is None:
            return seq
        if start is None:
            start = 0
        if end is None:
            end = len(seq)
        return seq[start:end]

    def get_reference_length(self, reference):
        if isinstance(self._seqs, _Indexed):
            return self._seqs.length(reference)
        return len(self._seqs[reference])


class AlignmentFile:  # import-time placeholder only
    def __init__(self, *a, **k):
        raise NotImplementedError('pysam stub: AlignmentFile is not available')


# ---------------------------------------------------------------------------------------------------------
# Minimal SAM-text reader with the record attributes pavlib.align.get_align_bed uses
# (pavlib/align/align.py:687-770). Attribute semantics follow the pysam documentation:
#   reference_start  0-based leftmost coordinate            reference_end   start + bases consumed on the reference (M D N = X)
#   query_alignment_start  index of the first aligned base in the stored SEQ (= leading soft clip; hard clips are not stored)
#   query_alignment_end    one past the last aligned base in the stored SEQ (= start + M I = X bases)
#   cigartuples  [(op code, length)] with BAM codes M0 I1 D2 N3 S4 H5 P6 =7 X8
# ---------------------------------------------------------------------------------------------------------
_CIGAR_CODE = {c: i for i, c in enumerate('MIDNSHP=X')}


class AlignedSegment:
    def __init__(self, line):
        import re
        tok = line.rstrip('\n').split('\t')
        self.query_name = tok[0]
        self.flag = int(tok[1])
        self.reference_name = tok[2]
        self.reference_start = int(tok[3]) - 1
        self.mapping_quality = int(tok[4])
        cig = tok[5]
        self.cigartuples = [] if cig == '*' else [(_CIGAR_CODE[o], int(n)) for n, o in re.findall(r'(\d+)([MIDNSHP=X])', cig)]
        self.cigar = self.cigartuples
        self.is_unmapped = bool(self.flag & 0x4)
        self.is_reverse = bool(self.flag & 0x10)
        ref_bp = sum(n for o, n in self.cigartuples if o in (0, 2, 3, 7, 8))
        self.reference_end = self.reference_start + ref_bp
        lead_s = 0
        for o, n in self.cigartuples:
            if o == 5:
                continue
            if o == 4:
                lead_s += n
            break
        self.query_alignment_start = lead_s
        self.query_alignment_end = lead_s + sum(n for o, n in self.cigartuples if o in (0, 1, 7, 8))
        self._tags = []
        for t in tok[11:]:
            k, ty, v = t.split(':', 2)
            self._tags.append((k, int(v) if ty == 'i' else (float(v) if ty == 'f' else v)))

    def get_tags(self):
        return list(self._tags)


class AlignmentFile:  # noqa: F811  (replaces the import-time placeholder above)
    def __init__(self, filename, mode='r', **kw):
        opener = gzip.open if str(filename).endswith('.gz') else open
        self._fh = opener(filename, 'rt')

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._fh.close()
        return False

    def __iter__(self):
        for line in self._fh:
            if line.startswith('@') or not line.strip():
                continue
            yield AlignedSegment(line)
