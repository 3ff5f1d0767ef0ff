"""Test-only stand-in for pysam (absent from this image), used ONLY to run the unmodified
reference from /root/reference when generating golden fixtures (tests/golden/make_golden.py).

Implements the subset the reference's hot path touches:
pysam.FastaFile(fn) as a context manager with fetch(name, start=None, end=None)
(pavlib/cigarcall.py:59-66, pavlib/seq.py:339-351). Plain or gzip FASTA; whole file in memory.
"""
import gzip

_CACHE = {}


def _load(fn):
    if fn in _CACHE:
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
        self.lengths = [len(v) for v in self._seqs.values()]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def fetch(self, reference=None, start=None, end=None):
        seq = self._seqs[reference]
        if start is None and end is None:
            return seq
        if start is None:
            start = 0
        if end is None:
            end = len(seq)
        return seq[start:end]

    def get_reference_length(self, reference):
        return len(self._seqs[reference])


class AlignmentFile:  # import-time placeholder only
    def __init__(self, *a, **k):
        raise NotImplementedError('pysam stub: AlignmentFile is not available')
