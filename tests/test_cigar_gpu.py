"""GPU parity for Path A: pav_b200.pavlib.cigarcall (CUDA, through the C ABI) against
(1) golden fixtures produced by the unmodified reference and (2) the CPU oracle on seeded inputs."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from pav_b200 import synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CIGAR_CASES = sorted(os.listdir(os.path.join(GOLDEN, 'cigar')))


def read_align(path):
    return pd.read_csv(path, sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)


def tsv_bytes(df):
    return df.to_csv(sep='\t', index=False).encode()


@pytest.mark.parametrize('case', CIGAR_CASES)
def test_cigar_golden_gpu(case):
    from pav_b200.pavlib import cigarcall
    d = os.path.join(GOLDEN, 'cigar', case)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    df_align = read_align(os.path.join(d, 'align.bed'))
    before = df_align.copy()
    args = (df_align, os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), meta['hap'])
    if 'exception' in meta:
        exc = {'RuntimeError': RuntimeError, 'IndexError': IndexError}[meta['exception']]
        with pytest.raises(exc) as ei:
            cigarcall.make_insdel_snv_calls(*args, version_id=meta['version_id'])
        assert str(ei.value) == meta['message']
        return
    df_snv, df_insdel = cigarcall.make_insdel_snv_calls(*args, version_id=meta['version_id'])
    assert tsv_bytes(df_snv) == open(os.path.join(d, 'snv.tsv'), 'rb').read()
    assert tsv_bytes(df_insdel) == open(os.path.join(d, 'insdel.tsv'), 'rb').read()
    assert [int(i) for i in df_snv.index] == meta['snv_index']
    assert [int(i) for i in df_insdel.index] == meta['insdel_index']
    assert [str(t) for t in df_snv.dtypes] == meta['snv_dtypes']
    assert [str(t) for t in df_insdel.dtypes] == meta['insdel_dtypes']
    assert list(df_snv.columns) == cigarcall.SNV_COLUMNS and list(df_insdel.columns) == cigarcall.INSDEL_COLUMNS
    pd.testing.assert_frame_equal(df_align, before)  # caller's table must not be mutated


def _workload(tmp_path, seed, **kw):
    ref, tigs, df = synth.make_cigar_workload(seed, **kw)
    ref_fa, tig_fa, _ = synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    return ref_fa, tig_fa, df


@pytest.mark.parametrize('seed,kw', [
    (11, dict(n_chrom=2, chrom_len=400_000, n_contig=40, contig_len=20_000, edit_rate=0.01, rev_frac=0.5)),
    (12, dict(n_chrom=1, chrom_len=300_000, n_contig=3, contig_len=100_000, edit_rate=0.02, rev_frac=0.5, clip=(11, 3),
              soft_mask_frac=0.5, n_block_frac=0.05)),
    (13, dict(n_chrom=3, chrom_len=60_000, n_contig=180, contig_len=1_000, edit_rate=0.004, rev_frac=0.3)),  # tiny records
    (14, dict(n_chrom=1, chrom_len=2_000_000, n_contig=1, contig_len=2_000_000, edit_rate=0.01, rev_frac=1.0)),  # one long record
])
def test_cigar_vs_oracle(tmp_path, seed, kw):
    """Bit-exact DataFrames against the CPU oracle on seeded synthetic alignments."""
    from oracle import pyoracle
    from pav_b200.pavlib import cigarcall
    ref_fa, tig_fa, df = _workload(tmp_path, seed, **kw)
    for vid in (False, True):
        g_snv, g_indel = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=vid)
        o_snv, o_indel = pyoracle.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=vid)
        assert tsv_bytes(g_snv) == tsv_bytes(o_snv)
        assert tsv_bytes(g_indel) == tsv_bytes(o_indel)
        assert (g_snv.index == o_snv.index).all() and (g_indel.index == o_indel.index).all()
        assert g_snv.shape[0] > 0 and g_indel.shape[0] > 0


def test_cigar_rows_full_size_properties(tmp_path):
    """C2-shaped slice (100 x 200 kbp): numeric rows equal the oracle's, row-count identities hold."""
    from oracle import pyoracle
    from pav_b200 import device, fasta
    ref, tigs, df = synth.config_c2(n_contig=100)
    ref_fa, tig_fa, _ = synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    o_snv, o_indel, _ = pyoracle.walk_rows(df, ref_fa, tig_fa)
    ctx = device.get_context()
    names_r, names_t = list(ref), list(tigs)
    rs = device.SeqStore(ctx, names_r, [ref[n] for n in names_r])
    ts = device.SeqStore(ctx, names_t, [tigs[n] for n in names_t])
    ops, op_off, perr = device.parse_cigars(df['CIGAR'].tolist())
    assert perr.code == 0
    rid = np.array([names_r.index(c) for c in df['#CHROM']], np.int32)
    qid = np.array([names_t.index(c) for c in df['QRY_ID']], np.int32)
    snv, indel, err, st = device.cigar_call(ctx, rs, ts, rid, qid, df['POS'].to_numpy(np.int32), df['REV'].to_numpy(np.uint8), ops, op_off)
    assert err.code == 0
    # row-count identities straight from the packed ops
    code, ln = ops & 15, ops >> 4
    assert len(snv) == int(ln[code == 8].sum()) and len(indel) == int(((code == 1) | (code == 2)).sum())
    assert (snv['pos_ref'] == o_snv['pos_ref']).all() and (snv['qry_pos'] == o_snv['qry_pos']).all() and (snv['rec'] == o_snv['rec']).all()
    for a, b in [('pos', 'pos'), ('end', 'end'), ('svlen', 'svlen'), ('qry_pos', 'qry_pos'), ('qry_end', 'qry_end'),
                 ('left_shift', 'left_shift'), ('hom_ref_l', 'hom_ref_l'), ('hom_ref_r', 'hom_ref_r'), ('hom_tig_l', 'hom_tig_l'),
                 ('hom_tig_r', 'hom_tig_r'), ('rec', 'rec'), ('svtype', 'svtype')]:
        assert (indel[a] == o_indel[b]).all(), a
    # emission order is (record, op, base)
    key = snv['rec'].astype(np.int64) * (1 << 32) + snv['op_idx']
    assert (np.diff(key) >= 0).all()
    rs.close(); ts.close()


def test_homology_golden_gpu():
    from pav_b200 import device
    cases = json.load(open(os.path.join(GOLDEN, 'homology.json')))
    groups = {}
    for c in cases:
        groups.setdefault((c['seq'], c['sv']), []).append(c)
    for (seq, sv), lst in groups.items():
        if not seq:
            continue
        pos = [c['pos'] for c in lst]
        left, right = device.homology(seq, sv, pos)
        for c, l, r in zip(lst, left.tolist(), right.tolist()):
            if c['left'] is not None:
                assert l == c['left'], c
            if c['right'] is not None:
                assert r == c['right'], c


def test_homology_wraparound_and_empty_sv():
    """pavlib.call.left_homology / right_homology at the edges the reference answers in its own way: a negative position for the
    rightward scan (Python's negative indexing: the scan reads the sequence's tail, then carries on from its start), positions
    outside the sequence, an empty SV sequence -- values and exception types stored from the reference (homology_wrap.json)."""
    from pav_b200.pavlib import call
    cases = json.load(open(os.path.join(GOLDEN, 'homology_wrap.json')))
    assert len(cases) > 500
    for c in cases:
        for name, fn in (('left', call.left_homology), ('right', call.right_homology)):
            want = c[name]
            if isinstance(want, dict):
                with pytest.raises({'IndexError': IndexError, 'ZeroDivisionError': ZeroDivisionError}[want['error']]):
                    fn(c['pos'], c['seq'], c['sv'])
            else:
                assert fn(c['pos'], c['seq'], c['sv']) == want, (name, c)


def test_pavlib_call_signatures():
    from pav_b200.pavlib import call
    assert call.left_homology(5, None, 'A') == 0 and call.right_homology(5, 'ACGT', None) == 0
    assert call.left_homology(15, 'ACGATTACAGCAGCAG', 'CAG') == 9
    assert call.right_homology(7, 'ACGATTACAGCAGCAGT', 'CAG') == 9


def test_seqstore_pack_roundtrip():
    """2-bit plane and N-mask plane decode back to the upper-cased ACGT / non-ACGT classes."""
    from pav_b200 import device
    rng = np.random.default_rng(3)
    seqs = [np.frombuffer(b'ACGTacgtNnRYKM-*', np.uint8)[rng.integers(0, 16, n)] for n in (0, 1, 31, 32, 33, 127, 128, 129, 5000)]
    ctx = device.get_context()
    st = device.SeqStore(ctx, [f's{i}' for i in range(len(seqs))], seqs)
    pack2, nmask = st.export()
    for i, s in enumerate(seqs):
        off = st.offset(i)
        g = off + np.arange(len(s))
        code = (pack2[g >> 5] >> (62 - 2 * (g & 31)).astype(np.uint64)) & np.uint64(3)
        isn = (nmask[g >> 5] >> (g & 31).astype(np.uint32)) & 1
        up = s & 0xDF
        exp_n = ~np.isin(up, np.frombuffer(b'ACGT', np.uint8))
        assert (isn.astype(bool) == exp_n).all()
        exp_code = np.searchsorted(np.frombuffer(b'ACGT', np.uint8), up[~exp_n])
        assert (code[~exp_n] == exp_code).all()
    st.close()


def test_multipass_walk_matches_single_pass(tmp_path, monkeypatch):
    """The three-kernel walk (kept for oversized batches, PAVGPU_CIGAR_MULTIPASS=1) gives the same DataFrames as the
    single-pass kernel and as the oracle."""
    from oracle import pyoracle
    from pav_b200.pavlib import cigarcall
    ref_fa, tig_fa, df = _workload(tmp_path, 15, n_chrom=2, chrom_len=300_000, n_contig=25, contig_len=24_000, edit_rate=0.012,
                                   rev_frac=0.5, clip=(4, 2))
    single = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=True)
    assert cigarcall.last_stats['walk_passes'] == 1 and cigarcall.last_stats['kernel_launches'] in (3, 5)   # count (+ record scan by its last CTA), walk, homology (1 or 3 launches)
    monkeypatch.setenv('PAVGPU_CIGAR_MULTIPASS', '1')
    multi = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=True)
    assert cigarcall.last_stats['walk_passes'] == 3 and cigarcall.last_stats['kernel_launches'] == 4   # reduce, chunk scan, emit, homology
    monkeypatch.delenv('PAVGPU_CIGAR_MULTIPASS')
    orc = pyoracle.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=True)
    for a, b, c in zip(single, multi, orc):
        assert tsv_bytes(a) == tsv_bytes(b) == tsv_bytes(c)
        assert (a.index == b.index).all() and (a.index == c.index).all()


@pytest.mark.parametrize('case', CIGAR_CASES)
@pytest.mark.parametrize('kernel', ['gather', 'tiled', 'nbr', 'bulk', 'queue', 'split'])
def test_cigar_golden_gpu_tiled_homology(case, kernel, monkeypatch):
    """The golden cases again with every homology kernel named explicitly (gathers, per-warp shared-memory tiles, per-indel
    neighbourhoods through cp.async and through bulk copies + mbarrier, gathers with CTA-pooled scan rests, the three-launch split with global queues): partial warps, REV records, N runs, tandem repeats that
    leave the staged words, planes smaller than a neighbourhood."""
    monkeypatch.setenv('PAVGPU_HOMOLOGY', kernel)
    test_cigar_golden_gpu(case)


@pytest.mark.parametrize('seed,kw', [
    (11, dict(n_chrom=2, chrom_len=400_000, n_contig=40, contig_len=20_000, edit_rate=0.01, rev_frac=0.5)),
    (12, dict(n_chrom=1, chrom_len=300_000, n_contig=3, contig_len=100_000, edit_rate=0.02, rev_frac=0.5, clip=(11, 3),
              soft_mask_frac=0.5, n_block_frac=0.05)),
    (13, dict(n_chrom=3, chrom_len=60_000, n_contig=180, contig_len=1_000, edit_rate=0.004, rev_frac=0.3)),  # warps span many records
    (14, dict(n_chrom=1, chrom_len=2_000_000, n_contig=1, contig_len=2_000_000, edit_rate=0.01, rev_frac=1.0)),
    (16, dict(n_chrom=1, chrom_len=3_000_000, n_contig=3, contig_len=1_000_000, edit_rate=0.0004, rev_frac=0.5)),  # sparse: spans overflow the tile
])
def test_homology_kernels_agree(tmp_path, monkeypatch, seed, kw):
    """All six homology kernels and the oracle give the same indel rows; the stats say which kernel ran."""
    from oracle import pyoracle
    from pav_b200.pavlib import cigarcall
    ref_fa, tig_fa, df = _workload(tmp_path, seed, **kw)
    orc = pyoracle.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=False)
    for kernel, name in enumerate(['gather', 'tiled', 'nbr', 'bulk', 'queue', 'split']):   # gathers, warp tiles, per-indel neighbourhoods (cp.async / bulk copies), pooled rests
        monkeypatch.setenv('PAVGPU_HOMOLOGY', name)
        got = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=False)
        assert cigarcall.last_stats['homology_tiled'] == kernel
        assert tsv_bytes(got[1]) == tsv_bytes(orc[1]) and tsv_bytes(got[0]) == tsv_bytes(orc[0])
        assert got[1].shape[0] > 0


def test_homology_kernel_choice_c2_slice(tmp_path, monkeypatch):
    """A C2-shaped slice (1 indel / ~540 bp) through every kernel choice, old switches included (PAVGPU_HOMOLOGY_TILED=auto picks the
    tiled kernel for a dense batch): all give the same bytes and equal the oracle."""
    from oracle import pyoracle
    from pav_b200 import device
    monkeypatch.delenv('PAVGPU_HOMOLOGY_TILED', raising=False)
    monkeypatch.delenv('PAVGPU_HOMOLOGY', raising=False)
    ref, tigs, df = synth.config_c2(n_contig=100)
    ref_fa, tig_fa, _ = synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    _, o_indel, _ = pyoracle.walk_rows(df, ref_fa, tig_fa)
    ctx = device.get_context()
    names_r, names_t = list(ref), list(tigs)
    rs = device.SeqStore(ctx, names_r, [ref[n] for n in names_r])
    ts = device.SeqStore(ctx, names_t, [tigs[n] for n in names_t])
    ops, op_off, _ = device.parse_cigars(df['CIGAR'].tolist())
    rid = np.array([names_r.index(c) for c in df['#CHROM']], np.int32)
    qid = np.array([names_t.index(c) for c in df['QRY_ID']], np.int32)
    out = {}
    monkeypatch.delenv('PAVGPU_HOMOLOGY_NBR', raising=False)
    for mode in ('gather', 'auto', '1', 'nbr', 'bulk', 'tiled-auto', 'queue', 'split', 'default'):
        monkeypatch.delenv('PAVGPU_HOMOLOGY', raising=False)
        if mode == 'nbr':
            monkeypatch.setenv('PAVGPU_HOMOLOGY_TILED', '0')
            monkeypatch.setenv('PAVGPU_HOMOLOGY_NBR', '1')
        elif mode in ('auto', '1'):
            monkeypatch.setenv('PAVGPU_HOMOLOGY_TILED', mode)
            monkeypatch.setenv('PAVGPU_HOMOLOGY_NBR', '0')
        elif mode == 'default':
            monkeypatch.delenv('PAVGPU_HOMOLOGY_TILED', raising=False)
            monkeypatch.delenv('PAVGPU_HOMOLOGY_NBR', raising=False)
        else:
            monkeypatch.setenv('PAVGPU_HOMOLOGY', mode)
        _, indel, err, st = device.cigar_call(ctx, rs, ts, rid, qid, df['POS'].to_numpy(np.int32), df['REV'].to_numpy(np.uint8), ops, op_off)
        assert err.code == 0
        out[mode] = (indel.copy(), st.homology_tiled)
    assert [out[m][1] for m in ('gather', 'auto', '1', 'nbr', 'bulk', 'tiled-auto', 'queue', 'split', 'default')] == [0, 1, 1, 2, 3, 1, 4, 5, 4]
    assert all(out[m][0].tobytes() == out['gather'][0].tobytes() for m in out)
    for f in ('pos', 'end', 'svlen', 'qry_pos', 'qry_end', 'left_shift', 'hom_ref_l', 'hom_ref_r', 'hom_tig_l', 'hom_tig_r', 'rec', 'svtype'):
        assert (out['gather'][0][f] == o_indel[f]).all(), f
    rs.close(); ts.close()
