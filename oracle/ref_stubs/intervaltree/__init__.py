"""Test-only stand-in for the `intervaltree` package (absent from this image); see pysam stub
header. Linear-scan implementation of the subset used by pavlib/align/lift.py and pavlib/inv.py:
t[a:b] = data, t[a:b], t[p], len(t), iteration, overlap(), bool."""
import collections

Interval = collections.namedtuple('Interval', ['begin', 'end', 'data'])


class IntervalTree:
    def __init__(self, intervals=None):
        self._iv = list(intervals) if intervals is not None else []

    def __setitem__(self, key, data):
        if not isinstance(key, slice):
            raise TypeError('IntervalTree stub: slice expected')
        if key.start >= key.stop:
            raise ValueError('IntervalTree: null interval')
        self._iv.append(Interval(key.start, key.stop, data))

    def addi(self, begin, end, data=None):
        self[begin:end] = data

    def overlap(self, begin, end=None):
        if end is None and hasattr(begin, 'begin'):
            begin, end = begin.begin, begin.end
        if begin >= end:
            return set()
        return {iv for iv in self._iv if iv.begin < end and iv.end > begin}

    def at(self, p):
        return {iv for iv in self._iv if iv.begin <= p < iv.end}

    def __getitem__(self, key):
        if isinstance(key, slice):
            return self.overlap(key.start, key.stop)
        return self.at(key)

    def __len__(self):
        return len(self._iv)

    def __iter__(self):
        return iter(self._iv)

    def __bool__(self):
        return len(self._iv) > 0
