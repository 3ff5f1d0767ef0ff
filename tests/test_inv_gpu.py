"""GPU parity for the scan_for_inv driver (pav_b200.pavlib.inv) against the golden call produced by the
unmodified reference (tests/golden/inv/kat3: 60 kbp locus, 8 kbp inversion, two expansions)."""
import gzip
import io
import json
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class _K:
    def __init__(self, k):
        self.k_size = k


@pytest.mark.parametrize('case', sorted(os.listdir(os.path.join(GOLDEN, 'inv'))))
def test_scan_for_inv_golden(case, capsys):
    from pav_b200.pavlib import inv, lift, seq
    d = os.path.join(GOLDEN, 'inv', case)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    df_align = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str})
    fai = seq.get_df_fai(os.path.join(d, 'tig.fa.fai'))
    al = lift.AlignLift(df_align, fai)
    log = io.StringIO()
    call = inv.scan_for_inv(seq.region_from_string(meta['flag']), os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), al, _K(31), log=log)
    assert call is not None
    assert call.id == meta['id'] and call.svlen == meta['svlen']
    for key in ('region_ref_outer', 'region_ref_inner', 'region_tig_outer', 'region_tig_inner', 'region_ref_discovery',
                'region_tig_discovery'):
        assert str(getattr(call, key)) == meta[key], key
    gold = pd.read_csv(os.path.join(d, 'density.tsv.gz'), sep='\t', keep_default_na=False, na_values=[''])
    assert list(call.df.columns) == meta['df_columns'] and call.df.shape[0] == meta['df_rows']
    for col in ('INDEX', 'STATE_MER', 'STATE', 'KMER'):
        assert (call.df[col].to_numpy() == gold[col].to_numpy()).all(), col
    for col in ('KERN_FWD', 'KERN_FWDREV', 'KERN_REV'):
        np.testing.assert_allclose(call.df[col].to_numpy(), gold[col].to_numpy(), rtol=1e-9, atol=1e-300)
    assert call.df['FLANK'].fillna('').tolist() == gold['FLANK'].fillna('').tolist()
    assert call.df['MATCH'].fillna('').tolist() == gold['MATCH'].fillna('').tolist()
    text = log.getvalue()
    assert 'Scanning region: chr1:26001-34000' in text
    assert 'Found inversion: ' + meta['id'] in text
    assert 'INV Found: outer=' + meta['region_tig_outer'] in capsys.readouterr().out


def test_scan_for_inv_negative_and_limits(tmp_path):
    """No inversion => None after min_exp_count expansions; region larger than max_region_size => None."""
    from pav_b200 import synth
    from pav_b200.pavlib import inv, lift, seq
    rng = np.random.default_rng(8)
    s = synth.random_seq(rng, 40_000)
    ref_fa = synth.write_fasta(str(tmp_path / 'ref.fa'), {'chr1': s})
    tig_fa = synth.write_fasta(str(tmp_path / 'tig.fa'), {'tig1': s})
    df_align = pd.DataFrame([('chr1', 0, 40_000, 0, 'tig1', 0, 40_000, 40_000, False, '40000=')],
                            columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END', 'QRY_LEN', 'REV', 'CIGAR'])
    al = lift.AlignLift(df_align, seq.get_df_fai(tig_fa + '.fai'))
    log = io.StringIO()
    assert inv.scan_for_inv(seq.Region('chr1', 18_000, 22_000), ref_fa, tig_fa, al, _K(31), log=log) is None
    assert 'Found no inverted k-mer states after 1 expansion(s)' in log.getvalue()
    log = io.StringIO()
    assert inv.scan_for_inv(seq.Region('chr1', 18_000, 22_000), ref_fa, tig_fa, al, _K(31), log=log, max_region_size=5000) is None
    assert 'Region size exceeds max' in log.getvalue()


def test_scan_for_inv_batch_equals_single(tmp_path):
    """Many flagged loci scored per GPU batch give the same calls as one scan_for_inv per locus."""
    from pav_b200 import synth
    from pav_b200.pavlib import inv, lift, seq
    rng = np.random.default_rng(77)
    n = 400_000
    s = synth.random_seq(rng, n)
    t = s.copy()
    flags, truth = [], []
    for i, start in enumerate(range(20_000, n - 60_000, 60_000)):
        ln = int(rng.integers(3000, 9000))
        if i % 3 != 2:                      # two of three loci carry an inversion
            t[start:start + ln] = synth.revcomp(s[start:start + ln])
        flags.append(seq.Region('chr1', start + ln // 3, start + 2 * ln // 3))
        truth.append(i % 3 != 2)
    ref_fa = synth.write_fasta(str(tmp_path / 'ref.fa'), {'chr1': s})
    tig_fa = synth.write_fasta(str(tmp_path / 'tig.fa'), {'tig1': t})
    df_align = pd.DataFrame([('chr1', 0, n, 0, 'tig1', 0, n, n, False, f'{n}=')],
                            columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END', 'QRY_LEN', 'REV', 'CIGAR'])
    al = lift.AlignLift(df_align, seq.get_df_fai(tig_fa + '.fai'))
    batch = inv.scan_for_inv_batch(flags, ref_fa, tig_fa, al, _K(31))
    single = [inv.scan_for_inv(f, ref_fa, tig_fa, al, _K(31)) for f in flags]
    assert len(batch) == len(flags)
    for b, s1, has_inv in zip(batch, single, truth):
        assert (b is None) == (s1 is None) == (not has_inv)
        if b is not None:
            assert b.id == s1.id and str(b.region_ref_inner) == str(s1.region_ref_inner) and str(b.region_tig_outer) == str(s1.region_tig_outer)
            assert b.df.shape == s1.df.shape and (b.df['STATE'].to_numpy() == s1.df['STATE'].to_numpy()).all()
