"""Breakpoint homology with the reference's signatures (pavlib/call.py:542-647), evaluated on the GPU.

The CIGAR walk computes these inside ``homology_kernel`` for every indel of a batch; the functions
below are the single-call form other PAV code uses (pavlib/lgsv.py:194-309). Each call packs the two
strings into HBM and runs the same device routine -- there is no CPU implementation.
"""
from .. import device


def left_homology(pos_tig, seq_tig, seq_sv):
    """Perfect-homology bases upstream of ``pos_tig`` (0-based, inclusive start of the leftward scan)."""
    if seq_sv is None or seq_tig is None:
        return 0
    if pos_tig < 0:
        return 0
    if pos_tig >= len(seq_tig):
        raise IndexError('string index out of range')
    left, _ = device.homology(seq_tig, seq_sv, [pos_tig])
    return int(left[0])


def right_homology(pos_tig, seq_tig, seq_sv):
    """Perfect-homology bases downstream starting at ``pos_tig``."""
    if seq_sv is None or seq_tig is None:
        return 0
    if pos_tig >= len(seq_tig):
        return 0
    if pos_tig < 0:
        pos_tig_py = len(seq_tig) + pos_tig
        if pos_tig_py < 0:
            raise IndexError('string index out of range')
        # Python negative indexing semantics of the reference loop: position wraps, limit does not.
        raise NotImplementedError('right_homology with a negative position is not supported on the GPU path')
    _, right = device.homology(seq_tig, seq_sv, [pos_tig])
    return int(right[0])
