"""Randomised GPU parity: arbitrary op mixes (not derived from a real alignment) and windows with many
state runs, CUDA vs the CPU oracle. Seeds are fixed so failures reproduce."""
import os

import numpy as np
import pandas as pd
import pytest

from pav_b200 import synth

pytestmark = pytest.mark.gpu

ALPHABET = np.frombuffer(b'ACGTACGTACGTACGTacgtNnRY', np.uint8)   # mostly ACGT, some lower case / N / IUPAC


def _random_seq(rng, n, homopolymer_frac=0.15):
    s = ALPHABET[rng.integers(0, len(ALPHABET), n)]
    # plant repeats so that homology scans run long and wrap around the SV sequence
    k = int(n * homopolymer_frac / 40)
    for st in rng.integers(0, max(n - 60, 1), size=k):
        unit = np.frombuffer(b'ACGT', np.uint8)[rng.integers(0, 4, int(rng.integers(1, 5)))]
        ln = int(rng.integers(10, 60))
        s[st:st + ln] = np.resize(unit, ln)
    return s


def _random_record(rng, n_ops):
    ops = []
    codes = rng.choice(list('=XIDSH'), size=n_ops, p=[.38, .2, .17, .17, .04, .04])
    for c in codes:
        ln = int(rng.integers(1, 4)) if c == 'X' else int(rng.integers(1, 40))
        if c in 'ID' and rng.random() < 0.1:
            ln = int(rng.integers(40, 300))
        ops.append((ln, c))
    ref_span = sum(n for n, c in ops if c in '=XD')
    qry_span = sum(n for n, c in ops if c in '=XISH')
    return ''.join(f'{n}{c}' for n, c in ops), ref_span, qry_span


@pytest.mark.parametrize('seed', [101, 102, 103])
def test_random_op_mixes_vs_oracle(tmp_path, seed):
    from oracle import pyoracle
    from pav_b200.pavlib import cigarcall
    rng = np.random.default_rng(seed)
    n_rec = 60
    recs = [_random_record(rng, int(rng.integers(1, 400))) for _ in range(n_rec)]
    ref_len = max(r[1] for r in recs) + 500
    ref = {'chrA': _random_seq(rng, ref_len), 'chrB': _random_seq(rng, ref_len)}
    tigs, rows = {}, []
    for i, (cigar, rspan, qspan) in enumerate(recs):
        name = f't{i:03d}'
        tigs[name] = _random_seq(rng, qspan + int(rng.integers(0, 5)))   # a few spare bases past the walk
        pos = int(rng.integers(0, ref_len - rspan))
        if i % 7 == 0:
            pos = 0
        rows.append(('chrA' if i % 2 else 'chrB', pos, pos + rspan, 1000 + i, name, 0, qspan, len(tigs[name]), bool(rng.random() < 0.5), cigar))
    df = pd.DataFrame(rows, columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END', 'QRY_LEN', 'REV', 'CIGAR'])
    ref_fa = synth.write_fasta(str(tmp_path / 'ref.fa'), ref, line_width=61)
    tig_fa = synth.write_fasta(str(tmp_path / 'tig.fa'), tigs, line_width=53)
    for vid in (True, False):
        g = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h2', version_id=vid)
        o = pyoracle.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h2', version_id=vid)
        for a, b in zip(g, o):
            assert a.to_csv(sep='\t', index=False) == b.to_csv(sep='\t', index=False)
            assert (a.index == b.index).all()
    assert g[0].shape[0] > 100 and g[1].shape[0] > 100


@pytest.mark.parametrize('seed', [201, 202])
def test_many_state_runs_vs_oracle(seed):
    """Contig stitched from forward / reverse-complement / novel segments of the reference: hundreds of STATE_MER runs."""
    from oracle import pyoracle
    from pav_b200.pavlib import density
    rng = np.random.default_rng(seed)
    wins = []
    for w in range(4):
        n = int(rng.integers(12_000, 30_000))
        ref = synth.random_seq(rng, n)
        parts, p = [], 0
        while p < n:
            ln = int(rng.integers(60, 500))
            seg = ref[p:p + ln]
            kind = rng.random()
            if kind < 0.45:
                parts.append(seg)
            elif kind < 0.8:
                parts.append(synth.revcomp(seg))
            else:
                parts.append(synth.random_seq(rng, len(seg)))
            p += ln
        tig = np.concatenate(parts)
        # an inverted duplicate inside the reference gives FWDREV k-mers
        ref2 = ref.copy()
        a = int(rng.integers(0, n - 3000))
        ref2[a + 1500:a + 2500] = synth.revcomp(ref2[a:a + 1000])
        wins.append((ref2, tig, bool(w % 2), [20, 7, 33, 20][w]))
    res = density.density_windows(wins)
    n_runs_seen = 0
    for (r, t, rev, srs), g in zip(wins, res):
        rc, o = pyoracle.density_arrays(r.tobytes(), t.tobytes(), rev=rev, srs=srs)
        assert g['status'] == rc == 0 and g['smoothed'] == o['smoothed']
        for c in ('KMER', 'INDEX', 'STATE_MER', 'STATE'):
            assert (g[c].astype(np.int64) == o[c].astype(np.int64)).all(), c
        assert g['n_eval'] == o['n_eval']
        for c in ('KERN_FWD', 'KERN_FWDREV', 'KERN_REV'):
            np.testing.assert_allclose(g[c], o[c], rtol=1e-9, atol=1e-300, err_msg=c)
        n_runs_seen += int((np.diff(g['STATE_MER']) != 0).sum()) + 1
    assert n_runs_seen > 40
