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


@pytest.mark.parametrize('batch', [0, 1, 5])
def test_call_inv_batch_rule(batch, tmp_path):
    """rule call_inv_batch (flagged regions -> inversion calls): the table and the log the reference's own rule body wrote
    (tests/golden/flag/inv_batch: two inversions, one found twice, one region without an inversion, one row of another batch)."""
    import io

    from pav_b200.pavlib import flag, seq
    d = os.path.join(os.path.dirname(GOLDEN), 'inv_batch')
    df_flag = pd.read_csv(os.path.join(d, 'flagged.bed.gz'), sep='\t', header=0)
    df_aln = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t')
    log = io.StringIO()
    df_bed = flag.call_inv_batch(df_flag, batch, os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), df_aln, seq.get_df_fai(os.path.join(d, 'tig.fa.fai')),
                                 'h1', log=log, density_out_dir=str(tmp_path / 'density'))
    assert df_bed.to_csv(sep='\t', index=False) == gzip.open(os.path.join(d, f'inv_call_{batch}.bed.gz'), 'rt').read()
    gold_log = os.path.join(d, f'inv_call_{batch}.log')
    if os.path.exists(gold_log):
        assert log.getvalue() == open(gold_log).read()
    if batch == 0:
        assert sorted(os.listdir(tmp_path / 'density')) == ['density_chr1-26001-INV-8000_h1.tsv.gz', 'density_chr1-85958-INV-5086_h1.tsv.gz']
