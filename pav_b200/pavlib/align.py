"""CIGAR tokenizer with the reference's interface and error text (pavlib/align/align.py:286-322).

The hot path does not call this generator (records are tokenised in bulk by
``pavgpu_cigar_parse``); it exists because other PAV code imports it from ``pavlib.align``.
"""
import pandas as pd

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
