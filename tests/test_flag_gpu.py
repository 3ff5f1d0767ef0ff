"""rule call_cigar end to end (GPU walk + FILTER) against the tables the reference's own rule body wrote
(tests/golden/flag/filter, see tests/golden/make_golden_flag.py)."""
import gzip
import os

import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'flag', 'filter')


@pytest.mark.parametrize('batch', [0, 1])
def test_call_cigar_rule(batch):
    from pav_b200.pavlib import flag
    df_align = pd.read_csv(os.path.join(GOLDEN, 'wl_align.bed'), sep='\t', dtype={'#CHROM': str}, keep_default_na=False, low_memory=False)
    df_trim = pd.read_csv(os.path.join(GOLDEN, 'wl_align_trim.bed'), sep='\t', usecols=['POS', 'END', 'INDEX'], index_col='INDEX').astype(int)
    df_snv, df_insdel = flag.call_cigar(df_align, batch, os.path.join(GOLDEN, 'wl_ref.fa'), os.path.join(GOLDEN, 'wl_tig.fa'), 'h1', df_trim)
    assert df_snv.to_csv(sep='\t', index=False) == gzip.open(os.path.join(GOLDEN, f'snv_{batch}.bed.gz'), 'rt').read()
    assert df_insdel.to_csv(sep='\t', index=False) == gzip.open(os.path.join(GOLDEN, f'insdel_{batch}.bed.gz'), 'rt').read()


@pytest.mark.parametrize('batch', [0, 1])
def test_call_cigar_rule_to_files(batch, tmp_path):
    """Same rule, tables written straight from device rows (C TSV writer + parallel gzip members): the files decompress to the bytes
    the reference's rule wrote."""
    from pav_b200.pavlib import flag
    df_align = pd.read_csv(os.path.join(GOLDEN, 'wl_align.bed'), sep='\t', dtype={'#CHROM': str}, keep_default_na=False, low_memory=False)
    df_trim = pd.read_csv(os.path.join(GOLDEN, 'wl_align_trim.bed'), sep='\t', usecols=['POS', 'END', 'INDEX'], index_col='INDEX').astype(int)
    out_i, out_s = str(tmp_path / 'insdel.bed.gz'), str(tmp_path / 'snv.bed.gz')
    n = flag.call_cigar_to_files(df_align, batch, os.path.join(GOLDEN, 'wl_ref.fa'), os.path.join(GOLDEN, 'wl_tig.fa'), 'h1', df_trim, out_i, out_s, threads=2)
    assert gzip.open(out_s, 'rt').read() == gzip.open(os.path.join(GOLDEN, f'snv_{batch}.bed.gz'), 'rt').read()
    assert gzip.open(out_i, 'rt').read() == gzip.open(os.path.join(GOLDEN, f'insdel_{batch}.bed.gz'), 'rt').read()
    assert n[0] > 0 and n[1] > 0
    # a batch without records writes header-only tables, like the reference's rule
    n = flag.call_cigar_to_files(df_align, 7, os.path.join(GOLDEN, 'wl_ref.fa'), os.path.join(GOLDEN, 'wl_tig.fa'), 'h1', df_trim, out_i, out_s)
    assert n == (0, 0) and pd.read_csv(out_s, sep='\t').shape[0] == 0 and 'FILTER' in pd.read_csv(out_i, sep='\t').columns
