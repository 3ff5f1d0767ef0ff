"""Test-only stand-in for Biopython (absent from this image); see pysam stub header."""
from . import Seq, SeqIO, bgzf, SeqRecord  # noqa: F401
