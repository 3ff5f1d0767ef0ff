"""Breakpoint homology with the reference's signatures (pavlib/call.py:542-647), evaluated on the GPU.

The CIGAR walk computes these inside ``homology_kernel`` for every indel of a batch; the functions
below are the single-call form other PAV code uses (pavlib/lgsv.py:194-309). Each call packs the two
strings into HBM and runs the same device routine -- there is no CPU implementation.
"""
from .. import device


def _empty_sv(base_getter):
    """The reference with an empty SV sequence: the first ACGT base it looks at ends in ``x % 0`` (pavlib/call.py:582,637)."""
    base = base_getter()
    if base not in {'A', 'C', 'G', 'T'}:
        return 0
    raise ZeroDivisionError('integer modulo by zero')


def left_homology(pos_tig, seq_tig, seq_sv):
    """Perfect-homology bases upstream of ``pos_tig`` (0-based, inclusive start of the leftward scan)."""
    if seq_sv is None or seq_tig is None:
        return 0
    if pos_tig < 0:
        return 0
    if pos_tig >= len(seq_tig):
        raise IndexError('string index out of range')
    if len(seq_sv) == 0:
        return _empty_sv(lambda: seq_tig[pos_tig])
    left, _ = device.homology(seq_tig, seq_sv, [pos_tig])
    return int(left[0])


def right_homology(pos_tig, seq_tig, seq_sv):
    """Perfect-homology bases downstream starting at ``pos_tig``."""
    if seq_sv is None or seq_tig is None:
        return 0
    if pos_tig >= len(seq_tig):
        return 0
    if pos_tig < 0:
        # The reference's loop indexes seq_tig[pos_tig + hom_len] with hom_len < len(seq_tig) - pos_tig (pavlib/call.py:627-634):
        # a negative position reads the last |pos_tig| bases through Python's negative indexing and then carries on from the start
        # of the sequence, i.e. it scans seq_tig[pos_tig:] + seq_tig from its first base. Same scan on the device over that string.
        if len(seq_tig) + pos_tig < 0:
            raise IndexError('string index out of range')
        if len(seq_sv) == 0:
            return _empty_sv(lambda: seq_tig[pos_tig])
        _, right = device.homology(seq_tig[pos_tig:] + seq_tig, seq_sv, [0])
        return int(right[0])
    if len(seq_sv) == 0:
        return _empty_sv(lambda: seq_tig[pos_tig])
    _, right = device.homology(seq_tig, seq_sv, [pos_tig])
    return int(right[0])
