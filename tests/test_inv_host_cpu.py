"""Host logic of the inversion scan (pav_b200.pavlib.inv: expansion driver, lift, run-length decisions, dup-mer annotation, log
text) and of `rule call_inv_batch` (pav_b200.pavlib.flag.call_inv_batch) against the reference's goldens WITHOUT a GPU: the
device density batch is replaced, for these tests only, by the CPU oracle (oracle/pav_oracle.c) behind the same
`density_windows` interface. The `-m gpu` twins (tests/test_inv_gpu.py, tests/test_flag_gpu.py) run the same goldens through CUDA."""
import gzip
import io
import json
import os

import numpy as np
import pandas as pd
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class _K:
    def __init__(self, k):
        self.k_size = k


@pytest.fixture
def oracle_density(monkeypatch):
    """pavdensity.density_windows served by the oracle: same arguments, same list of per-window dicts."""
    from oracle import pyoracle
    from pav_b200.pavlib import density

    def density_windows(windows, k=31, ctx=None, lazy=False, min_informative=2000, min_state_count=20, smooth=1.0, delta=0.005, max_ref_kmer_count=100):
        assert max_ref_kmer_count == 100
        out = []
        for ref, tig, rev, srs in windows:
            rc, d = pyoracle.density_arrays(np.asarray(ref, np.uint8).tobytes(), np.asarray(tig, np.uint8).tobytes(), k=k, rev=bool(rev),
                                            min_inf=min_informative, smooth=smooth, min_state=min_state_count, srs=int(srs), delta=delta)
            if rc != 0:
                out.append({'status': rc, 'smoothed': False, 'n_eval': 0})
                continue
            d['status'] = 0
            if not d['smoothed']:
                d['STATE'] = np.full(len(d['INDEX']), -1, np.int8)
            out.append(d)
        return out
    monkeypatch.setattr(density, 'density_windows', density_windows)
    return density_windows


@pytest.mark.parametrize('case', sorted(os.listdir(os.path.join(GOLDEN, 'inv'))))
def test_scan_for_inv_golden_host(case, oracle_density, capsys):
    from pav_b200.pavlib import inv, lift, seq
    d = os.path.join(GOLDEN, 'inv', case)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    df_align = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str})
    al = lift.AlignLift(df_align, seq.get_df_fai(os.path.join(d, 'tig.fa.fai')))
    log = io.StringIO()
    call = inv.scan_for_inv(seq.region_from_string(meta['flag']), os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), al, _K(31), log=log)
    assert call is not None and call.id == meta['id'] and call.svlen == meta['svlen']
    for key in ('region_ref_outer', 'region_ref_inner', 'region_tig_outer', 'region_tig_inner', 'region_ref_discovery', 'region_tig_discovery'):
        assert str(getattr(call, key)) == meta[key], key
    gold = pd.read_csv(os.path.join(d, 'density.tsv.gz'), sep='\t', keep_default_na=False, na_values=[''])
    assert list(call.df.columns) == meta['df_columns'] and call.df.shape[0] == meta['df_rows']
    for col in ('INDEX', 'STATE_MER', 'STATE', 'KMER'):
        assert (call.df[col].to_numpy() == gold[col].to_numpy()).all(), col
    assert call.df['FLANK'].fillna('').tolist() == gold['FLANK'].fillna('').tolist()
    assert call.df['MATCH'].fillna('').tolist() == gold['MATCH'].fillna('').tolist()
    assert 'Found inversion: ' + meta['id'] in log.getvalue()
    assert 'INV Found: outer=' + meta['region_tig_outer'] in capsys.readouterr().out
    # the batch driver on the same locus: same call, same log
    log_b = io.StringIO()
    call_b = inv.scan_for_inv_batch([seq.region_from_string(meta['flag'])], os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), al, _K(31), log=log_b)[0]
    assert call_b.id == call.id and str(call_b.region_tig_inner) == str(call.region_tig_inner) and log_b.getvalue() == log.getvalue()
    assert call_b.df.equals(call.df)


@pytest.mark.parametrize('batch', [0, 1, 5])
def test_call_inv_batch_rule_host(batch, oracle_density, tmp_path):
    """Table and log of `rule call_inv_batch` equal what the reference's rule body wrote (tests/golden/flag/inv_batch)."""
    from pav_b200.pavlib import flag, seq
    d = os.path.join(GOLDEN, 'flag', 'inv_batch')
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


def test_scan_for_inv_limits_host(oracle_density, tmp_path):
    """No inversion => None after the minimum number of expansions; a region beyond max_region_size => None; both say so in the log."""
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
