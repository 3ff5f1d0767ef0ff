"""Pin the CPU oracle (oracle/pav_oracle.c + oracle/pyoracle.py) against golden vectors produced by
the unmodified reference (tests/golden/make_golden.py). CPU only."""
import gzip
import io
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import pyoracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CIGAR_CASES = sorted(os.listdir(os.path.join(GOLDEN, 'cigar')))
DENSITY_CASES = sorted(os.listdir(os.path.join(GOLDEN, 'density')))


def read_align(path):
    return pd.read_csv(path, sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)


def tsv_bytes(df):
    return df.to_csv(sep='\t', index=False).encode()


@pytest.mark.parametrize('case', CIGAR_CASES)
def test_cigar_golden(case):
    d = os.path.join(GOLDEN, 'cigar', case)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    df_align = read_align(os.path.join(d, 'align.bed'))
    args = (df_align, os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), meta['hap'])
    if 'exception' in meta:
        exc = {'RuntimeError': RuntimeError, 'IndexError': IndexError}[meta['exception']]
        with pytest.raises(exc) as ei:
            pyoracle.make_insdel_snv_calls(*args, version_id=meta['version_id'])
        assert str(ei.value) == meta['message']
        return
    df_snv, df_insdel = pyoracle.make_insdel_snv_calls(*args, version_id=meta['version_id'])
    assert tsv_bytes(df_snv) == open(os.path.join(d, 'snv.tsv'), 'rb').read()
    assert tsv_bytes(df_insdel) == open(os.path.join(d, 'insdel.tsv'), 'rb').read()
    assert [int(i) for i in df_snv.index] == meta['snv_index']
    assert [int(i) for i in df_insdel.index] == meta['insdel_index']
    assert [str(t) for t in df_snv.dtypes] == meta['snv_dtypes']
    assert [str(t) for t in df_insdel.dtypes] == meta['insdel_dtypes']
    if meta['n_snv'] + meta['n_insdel'] < 600:  # the reference-container mode gives the same frames
        r_snv, r_insdel = pyoracle.make_insdel_snv_calls(*args, version_id=meta['version_id'], reference_containers=True)
        assert tsv_bytes(r_snv) == tsv_bytes(df_snv) and tsv_bytes(r_insdel) == tsv_bytes(df_insdel)
        assert (r_snv.index == df_snv.index).all() and (r_insdel.index == df_insdel.index).all()
        assert [str(t) for t in r_snv.dtypes] == meta['snv_dtypes']


def test_homology_golden():
    cases = json.load(open(os.path.join(GOLDEN, 'homology.json')))
    for c in cases:
        if c['left'] is not None:
            assert pyoracle.left_homology(c['pos'], c['seq'], c['sv']) == c['left'], c
        if c['right'] is not None:
            assert pyoracle.right_homology(c['pos'], c['seq'], c['sv']) == c['right'], c


def test_kmer_golden():
    cases = json.load(open(os.path.join(GOLDEN, 'kmer.json')))
    for c in cases:
        if c['k'] > 32:
            continue
        km, ix = pyoracle.kmer_stream(c['seq'], c['k'])
        assert [str(int(x)) for x in km] == c['kmers']
        assert [int(x) for x in ix] == c['index']
        assert [str(pyoracle.kmer_rc(int(x), c['k'])) for x in km] == c['rc']


def load_density_case(case):
    d = os.path.join(GOLDEN, 'density', case)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    fa_r = pyoracle.read_fasta(os.path.join(d, 'ref.fa'))['chrW']
    fa_t = pyoracle.read_fasta(os.path.join(d, 'tig.fa'))['tigW']

    def sub(seq, rgn):
        a, b = rgn.split(':')[1].split('-')
        return seq[int(a) - 1:int(b)]
    df = None
    p = os.path.join(d, 'density.tsv.gz')
    if os.path.exists(p):
        df = pd.read_csv(p, sep='\t')
    return meta, sub(fa_r, meta['refregion']), sub(fa_t, meta['tigregion']), df


KERN_RTOL = 1e-9   # float64 KDE: different summation order / libm exp than scipy's Cython loop
KERN_ATOL = 1e-300


@pytest.mark.parametrize('case', DENSITY_CASES)
def test_density_golden(case):
    meta, ref_seq, tig_seq, gold = load_density_case(case)
    rc, out = pyoracle.density_arrays(ref_seq, tig_seq, k=meta['k'], rev=meta['rev'], srs=meta['srs'])
    assert rc == meta['returncode']
    if rc != 0:
        return
    df = pyoracle.density_frame(out)
    assert list(df.columns) == meta['columns']
    assert df.shape[0] == meta['n_rows']
    for col in ('INDEX', 'STATE_MER', 'STATE', 'KMER'):
        assert (df[col].to_numpy().astype(np.int64) == gold[col].to_numpy().astype(np.int64)).all(), col
    if out['smoothed']:
        for col in ('KERN_FWD', 'KERN_FWDREV', 'KERN_REV'):
            np.testing.assert_allclose(df[col].to_numpy(), gold[col].to_numpy(), rtol=KERN_RTOL, atol=KERN_ATOL, err_msg=col)
    assert [list(map(int, r)) for r in pyoracle.rl_encoder(df)] == meta['rl_state']
