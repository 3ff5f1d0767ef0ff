"""Flagging rules between Path A and Path B (pav_b200/pavlib/flag.py) against golden tables produced by executing the
reference's own rule bodies (tests/golden/make_golden_flag.py): call_inv_cluster, call_inv_flag_insdel_cluster,
call_inv_merge_flagged_loci and the FILTER step of call_cigar. Tables are compared as the TSV text the rules write."""
import os

import pandas as pd
import pytest

from pav_b200.pavlib import flag

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'flag')
CASES = ['a', 'b', 'empty', 'no_snv', 'no_sv']
CONFIG = {'b': dict(flank=5000, batch_count=7)}


def tsv(df):
    return df.to_csv(sep='\t', index=False)


def golden_text(path):
    import gzip
    return gzip.open(path, 'rt').read()


def read(path):
    return pd.read_csv(path, sep='\t', header=0, low_memory=False, dtype={'#CHROM': str})


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('vartype', ['snv', 'indel'])
def test_cluster(case, vartype):
    d = os.path.join(GOLDEN, case)
    src = read(os.path.join(d, 'snv.bed.gz' if vartype == 'snv' else 'insdel.bed.gz'))
    got = flag.cluster_variants([src], vartype)
    assert tsv(got) == golden_text(os.path.join(d, f'cluster_{vartype}.bed.gz'))


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('vartype', ['sv', 'indel'])
def test_flag_insdel_cluster(case, vartype):
    d = os.path.join(GOLDEN, case)
    got = flag.flag_insdel_cluster(read(os.path.join(d, 'insdel.bed.gz')), vartype)
    assert tsv(got) == golden_text(os.path.join(d, f'insdel_{vartype}.bed.gz'))


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('filt', ['svindel', 'sv', 'single_cluster'])
def test_merge_flagged_loci(case, filt):
    d = os.path.join(GOLDEN, case)
    tabs = [pd.read_csv(os.path.join(d, f'{n}.bed.gz'), sep='\t') for n in ('insdel_sv', 'insdel_indel', 'cluster_indel', 'cluster_snv')]
    got = flag.merge_flagged_loci(*tabs, inv_sig_filter=filt, **CONFIG.get(case, {}))
    assert tsv(got) == golden_text(os.path.join(d, f'flagged_regions_{filt}.bed.gz'))


def test_merge_rejects_unknown_filter():
    e = pd.DataFrame([], columns=['#CHROM', 'POS', 'END'])
    c = pd.DataFrame([], columns=['#CHROM', 'POS', 'END', 'COUNT'])
    with pytest.raises(RuntimeError, match='Unrecognized region filter'):
        flag.merge_flagged_loci(e, e, c, c, inv_sig_filter='x')


@pytest.mark.parametrize('batch', [0, 1])
@pytest.mark.parametrize('kind', ['snv', 'insdel'])
def test_cigar_filter(batch, kind):
    """FILTER column recomputed from the golden call table + the trimmed alignment table."""
    d = os.path.join(GOLDEN, 'filter')
    gold = pd.read_csv(os.path.join(d, f'{kind}_{batch}.bed.gz'), sep='\t', keep_default_na=False)
    df_trim = pd.read_csv(os.path.join(d, 'wl_align_trim.bed'), sep='\t', usecols=['POS', 'END', 'INDEX'], index_col='INDEX').astype(int)
    got = flag.cigar_filter(gold.drop(columns=['FILTER']), df_trim)
    assert got.tolist() == gold['FILTER'].tolist()
    assert set(got) == {'PASS', 'TRIM'}


def test_call_cigar_merge(tmp_path):
    """rule call_cigar_merge over three batches == the tables the reference's rule body wrote."""
    d = os.path.join(GOLDEN, 'filter')
    out_i, out_s = str(tmp_path / 'i.bed.gz'), str(tmp_path / 's.bed.gz')
    flag.call_cigar_merge([os.path.join(d, f'insdel_{b}.bed.gz') for b in (0, 1, 2)], [os.path.join(d, f'snv_{b}.bed.gz') for b in (0, 1, 2)],
                          out_i, out_s, threads=2)
    assert golden_text(out_i) == golden_text(os.path.join(d, 'merged_insdel.bed.gz'))
    assert golden_text(out_s) == golden_text(os.path.join(d, 'merged_snv.bed.gz'))
