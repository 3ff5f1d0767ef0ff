"""Region value type and FASTA region access with the reference's interface (pavlib/seq.py:20-360)."""
import re

import numpy as np
import pandas as pd

from .. import fasta

_RGN_RE = re.compile(r'^([^:]+):(\d+)-(\d+)$')


class Region:
    """0-based half-open region with an orientation flag (reference: pavlib/seq.py:20-258).

    ``is_rev`` defaults to ``pos > end`` (coordinates are swapped in that case). ``pos_min/max`` and
    ``end_min/max`` carry breakpoint uncertainty; ``*_aln_index`` the alignment record of each end.
    """

    def __init__(self, chrom, pos, end, is_rev=None, pos_min=None, pos_max=None, end_min=None, end_max=None,
                 pos_aln_index=None, end_aln_index=None):
        self.chrom = str(chrom)
        self.pos = int(pos)
        self.end = int(end)
        self.pos_min = self.pos if pos_min is None else int(pos_min)
        self.pos_max = self.pos if pos_max is None else int(pos_max)
        self.end_min = self.end if end_min is None else int(end_min)
        self.end_max = self.end if end_max is None else int(end_max)
        self.pos_aln_index = pos_aln_index
        self.end_aln_index = end_aln_index
        if self.pos > self.end:
            self.pos, self.end = self.end, self.pos
            # same (quirky) min/max handling as the reference when coordinates arrive reversed
            self.end_min = self.pos if pos_min is None else int(pos_min)
            self.end_max = self.pos if pos_max is None else int(pos_max)
            self.pos_min = self.end if end_min is None else int(end_min)
            self.pos_max = self.end if end_max is None else int(end_max)
            self.pos_aln_index, self.end_aln_index = self.end_aln_index, self.pos_aln_index
            if is_rev is None:
                is_rev = True
        self.is_rev = False if is_rev is None else is_rev

    def __repr__(self):
        return self.to_base1_string()

    def to_base1_string(self):
        return '{}:{}-{}'.format(self.chrom, self.pos + 1, self.end)

    def to_bed_string(self):
        return '{}\t{}\t{}'.format(self.chrom, self.pos + 1, self.end)

    def __len__(self):
        return self.end - self.pos

    def region_id(self):
        return '{}-{}-RGN-{}'.format(self.chrom, self.pos, self.end - self.pos)

    def expand(self, expand_bp, min_pos=0, max_end=None, shift=True, balance=0.5):
        """Grow by ``expand_bp`` (``int(expand_bp * balance)`` upstream, the rest downstream), clamped to
        ``[min_pos, max_end]``; with ``shift`` the clipped amount moves to the free side.
        ``max_end`` may be a Series of chromosome lengths (reference: pavlib/seq.py:112-188)."""
        if balance is None:
            balance = 0.5
        try:
            if not (0 <= balance <= 1):
                raise RuntimeError('balance must be in range [0, 1]: {}'.format(balance))
        except ValueError:
            raise RuntimeError('balance is not numeric: {}'.format(balance))
        up = int(expand_bp * balance)
        down = np.max([0, expand_bp - up])
        new_pos = int(self.pos - up)
        new_end = int(self.end + down)
        if min_pos is not None and new_pos < min_pos:
            if shift:
                new_end += min_pos - new_pos
            new_pos = min_pos
        if max_end is not None:
            if max_end.__class__ == pd.core.series.Series and self.chrom in max_end.index:
                max_end = max_end[self.chrom]
            else:
                max_end = None
        if max_end is not None and new_end > max_end:
            if shift:
                new_pos -= new_end - max_end
                if new_pos < min_pos:
                    new_pos = min_pos
            new_end = max_end
        if new_end < new_pos:
            new_end = new_pos = (new_end + new_pos) // 2
        self.pos, self.end = new_pos, new_end
        self.pos_min = self.pos_max = self.pos
        self.end_min = self.end_max = self.end

    def __getitem__(self, key):
        if key not in {'chrom', 'pos', 'pos1', 'end'}:
            raise IndexError('No key in Region: {}'.format(key))
        return self.pos + 1 if key == 'pos1' else self.__dict__[key]

    def __eq__(self, other):
        return self.chrom == other.chrom and self.pos == other.pos and self.end == other.end

    def __lt__(self, other):
        return (self.chrom, self.pos, self.end) < (other.chrom, other.pos, other.end)

    def copy(self):
        return Region(self.chrom, self.pos, self.end, self.is_rev, self.pos_min, self.pos_max, self.end_min, self.end_max)


def region_from_string(rgn_str, is_rev=None, base0half=False):
    """``chrom:pos-end`` (1-based closed unless ``base0half``) -> Region (reference: pavlib/seq.py:260-285)."""
    m = _RGN_RE.match(rgn_str.replace(',', ''))
    if m is None:
        raise RuntimeError('Region is not in expected format (chrom:pos-end): {}'.format(rgn_str))
    pos, end = int(m[2]), int(m[3])
    if not base0half:
        pos -= 1
    return Region(m[1], pos, end, is_rev=is_rev)


def region_from_id(region_id):
    """``CHROM-POS-SVTYPE-LEN`` -> Region (reference: pavlib/seq.py:288-302)."""
    tok = region_id.split('-')
    if len(tok) != 4:
        raise RuntimeError('Unrecognized region ID: {}'.format(region_id))
    return Region(tok[0], int(tok[1]) - 1, int(tok[1]) - 1 + int(tok[3]))


def region_seq_fasta(region, fa_file_name, rev_compl=None):
    """Sequence of a Region (or whole record if ``region`` is a str) from an indexed FASTA
    (reference: pavlib/seq.py:328-360)."""
    fa = fasta.open_fasta(fa_file_name)
    if region.__class__ == str:
        arr, is_region = fa.fetch_array(region), False
    elif region.__class__ == Region:
        arr, is_region = fa.fetch_array(region.chrom, region.pos, region.end), True
    else:
        raise RuntimeError('Unrecognized region type: {}: Expected Region (pavlib.seq) or str'.format(str(region.__class__.__name__)))
    do_rc = (is_region and region.is_rev) if rev_compl is None else bool(rev_compl)
    if do_rc:
        arr = fasta.reverse_complement(arr)
    return arr.tobytes().decode('ascii')


def get_df_fai(fai_file_name):
    """Series of sequence lengths keyed by name (reference: dep/svpop/svpoplib/ref.py:109-139 defaults)."""
    df = pd.read_csv(fai_file_name, sep='\t', names=['CHROM', 'LEN', 'POS', 'LINE_BP', 'LINE_BYTES'],
                     usecols=('CHROM', 'LEN'), dtype={'CHROM': str, 'LEN': int})
    s = df.set_index('CHROM')['LEN']
    s.index = s.index.astype(str)
    return s
