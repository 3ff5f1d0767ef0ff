"""Host-side row formatting (pav_b200/pavlib/cigarcall.build_frames + pav_b200/csrc/pyrows.c) without a GPU: the numeric
rows come from the CPU oracle, converted to the device row layout (pavgpu_snv_row / pavgpu_indel_row); the frames built
from them must equal the golden tables of the unmodified reference and the oracle's own frames, byte for byte."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import pyoracle
from pav_b200 import _capi, fasta, synth
from pav_b200.pavlib import cigarcall

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
OK_CASES = [c for c in sorted(os.listdir(os.path.join(GOLDEN, 'cigar')))
            if 'exception' not in json.load(open(os.path.join(GOLDEN, 'cigar', c, 'meta.json')))]


def device_rows_from_oracle(df_align, ref_fa, tig_fa):
    snv, indel, _ = pyoracle.walk_rows(df_align, ref_fa, tig_fa)
    s = np.zeros(len(snv), _capi.SNV_ROW)
    for c in ('pos_ref', 'qry_pos', 'rec'):
        s[c] = snv[c]
    i = np.zeros(len(indel), _capi.INDEL_ROW)
    for c in ('rec', 'svtype', 'svlen', 'pos', 'end', 'qry_pos', 'qry_end', 'left_shift', 'hom_ref_l', 'hom_ref_r', 'hom_tig_l',
              'hom_tig_r', 'seq_start'):
        i[c] = indel[c]
    return s, i


def frames_from_oracle_rows(df_align, ref_fa, tig_fa, hap, version_id):
    if df_align.shape[0] == 0:
        return cigarcall._empty(cigarcall.SNV_COLUMNS), cigarcall._empty(cigarcall.INSDEL_COLUMNS)
    s, i = device_rows_from_oracle(df_align, ref_fa, tig_fa)
    table = cigarcall.AlignTable(df_align)
    rf, tf = fasta.open_fasta(ref_fa), fasta.open_fasta(tig_fa)
    ref_arr = [rf.fetch_array(nm) for nm in table.ref_names]
    tig_arr = [tf.fetch_array(nm) for nm in table.tig_names]
    return cigarcall.build_frames(s, i, table.chrom, table.qry, table.rev, table.align_index, ref_arr, tig_arr, table.ref_id,
                                  table.qry_id, hap, version_id)


def tsv(df):
    return df.to_csv(sep='\t', index=False).encode()


@pytest.mark.parametrize('case', OK_CASES)
def test_frames_golden(case):
    d = os.path.join(GOLDEN, 'cigar', case)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    df_align = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)
    df_snv, df_insdel = frames_from_oracle_rows(df_align, os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), meta['hap'], meta['version_id'])
    assert tsv(df_snv) == open(os.path.join(d, 'snv.tsv'), 'rb').read()
    assert tsv(df_insdel) == open(os.path.join(d, 'insdel.tsv'), 'rb').read()
    assert [int(x) for x in df_snv.index] == meta['snv_index']
    assert [int(x) for x in df_insdel.index] == meta['insdel_index']
    assert [str(t) for t in df_snv.dtypes] == meta['snv_dtypes']
    assert [str(t) for t in df_insdel.dtypes] == meta['insdel_dtypes']


@pytest.mark.parametrize('vid', [False, True])
def test_frames_vs_oracle_overlapping_records(tmp_path, vid):
    """Overlapping contigs (ties on #CHROM, POS, END -> ID tie-break and ID versioning), REV records, clips, soft-mask, N."""
    ref, tigs, df = synth.make_cigar_workload(21, n_chrom=2, chrom_len=60_000, n_contig=30, contig_len=12_000, edit_rate=0.02,
                                              rev_frac=0.5, clip=(7, 2), soft_mask_frac=0.3, n_block_frac=0.02)
    # duplicate some records under new contig names so that identical variants collide
    dup = df.iloc[:8].copy()
    dup['INDEX'] = np.arange(len(df), len(df) + len(dup))
    for q in dup['QRY_ID']:
        tigs[q + '_dup'] = tigs[q]
    dup['QRY_ID'] = dup['QRY_ID'] + '_dup'
    df = pd.concat([df, dup], ignore_index=True).sort_values(['#CHROM', 'POS', 'END', 'QRY_ID'], ascending=[True, True, False, True]).reset_index(drop=True)
    ref_fa, tig_fa, _ = synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    g_snv, g_indel = frames_from_oracle_rows(df, ref_fa, tig_fa, 'h2', vid)
    o_snv, o_indel = pyoracle.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h2', version_id=vid)
    assert tsv(g_snv) == tsv(o_snv) and tsv(g_indel) == tsv(o_indel)
    assert (g_snv.index == o_snv.index).all() and (g_indel.index == o_indel.index).all()
    assert all(str(t) == 'object' for t in g_snv.dtypes) and all(str(t) == 'object' for t in g_indel.dtypes)
    assert len(g_snv) > 100 and len(g_indel) > 20


def test_frames_non_ascii_names(tmp_path):
    """Contig / chromosome names outside ASCII take the generic formatter; the result is the same table."""
    ref, tigs, df = synth.make_cigar_workload(22, n_chrom=1, chrom_len=30_000, n_contig=4, contig_len=6_000, edit_rate=0.02, rev_frac=0.5)
    ren = {q: q + 'é' for q in tigs}
    tigs = {ren[q]: v for q, v in tigs.items()}
    df = df.copy()
    df['QRY_ID'] = [ren[q] for q in df['QRY_ID']]
    ref_fa, tig_fa = str(tmp_path / 'r.fa'), str(tmp_path / 't.fa')
    synth.write_fasta(ref_fa, ref)
    with open(tig_fa, 'w', encoding='utf-8') as fh:
        for q, v in tigs.items():
            fh.write(f'>{q}\n{bytes(v).decode()}\n')
    s, i = device_rows_from_oracle(df, ref_fa, tig_fa)
    table = cigarcall.AlignTable(df)
    rf = fasta.open_fasta(ref_fa)
    ref_arr = [rf.fetch_array(nm) for nm in table.ref_names]
    tig_arr = [tigs[nm] for nm in table.tig_names]
    g_snv, g_indel = cigarcall.build_frames(s, i, table.chrom, table.qry, table.rev, table.align_index, ref_arr, tig_arr, table.ref_id,
                                            table.qry_id, 'h1', True)
    o_snv, o_indel = pyoracle.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=True)
    assert tsv(g_snv) == tsv(o_snv) and tsv(g_indel) == tsv(o_indel)


@pytest.mark.parametrize('seed', range(6))
def test_sort_order_equals_pandas(seed):
    """_sort_order (with its already-sorted shortcuts) == pandas' stable sort_values(['#CHROM', 'POS', 'END', 'ID'])."""
    rng = np.random.default_rng(seed)
    n = 400
    chrom_codes = np.sort(rng.integers(0, 3, n)).astype(np.int64)
    pos = rng.integers(0, 60, n).astype(np.int64)
    if seed % 2 == 0:      # rows that arrive sorted by (#CHROM, POS), with ties
        o = np.lexsort((pos, chrom_codes))
        chrom_codes, pos = chrom_codes[o], pos[o]
    if seed == 4:          # strictly increasing: the tie-free shortcut
        pos = np.arange(n, dtype=np.int64)
    ids = np.array([f'id{int(x)}' for x in rng.integers(0, 50, n)], dtype=object)
    for snv_like in (True, False):
        end = pos + 1 if snv_like else pos + rng.integers(1, 4, n)
        got = cigarcall._sort_order(chrom_codes, pos, end, ids.__getitem__, end_is_pos_plus_1=snv_like)
        df = pd.DataFrame({'#CHROM': chrom_codes, 'POS': pos, 'END': end, 'ID': ids})
        exp = df.sort_values(['#CHROM', 'POS', 'END', 'ID']).index.to_numpy()
        assert (got == exp).all()


@pytest.mark.parametrize('case', OK_CASES)
def test_tables_tsv_equals_to_csv(case):
    """The C TSV writer (no frames) produces the bytes DataFrame.to_csv writes for the frames, FILTER column included."""
    d = os.path.join(GOLDEN, 'cigar', case)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    df_align = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)
    if df_align.shape[0] == 0:
        return
    ref_fa, tig_fa = os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa')
    s, i = device_rows_from_oracle(df_align, ref_fa, tig_fa)
    table = cigarcall.AlignTable(df_align)
    rf, tf = fasta.open_fasta(ref_fa), fasta.open_fasta(tig_fa)
    ref_arr = [rf.fetch_array(nm) for nm in table.ref_names]
    tig_arr = [tf.fetch_array(nm) for nm in table.tig_names]
    args = (s, i, table.chrom, table.qry, table.rev, table.align_index, ref_arr, tig_arr, table.ref_id, table.qry_id, meta['hap'], meta['version_id'])
    text = cigarcall.tables_tsv(*args)
    assert text[0] == open(os.path.join(d, 'snv.tsv'), 'rb').read() and text[1] == open(os.path.join(d, 'insdel.tsv'), 'rb').read()
    rng = np.random.default_rng(0)
    ps, pi = rng.integers(0, 2, len(s)).astype(np.uint8), rng.integers(0, 2, len(i)).astype(np.uint8)
    text = cigarcall.tables_tsv(*args, pass_snv=ps, pass_indel=pi)
    df_snv, df_insdel = cigarcall.build_frames(*args)
    df_snv['FILTER'] = pd.Series(np.where(ps, 'PASS', 'TRIM'), dtype=object).to_numpy()[df_snv.index.to_numpy()] if len(s) else []
    df_insdel['FILTER'] = pd.Series(np.where(pi, 'PASS', 'TRIM'), dtype=object).to_numpy()[df_insdel.index.to_numpy()] if len(i) else []
    assert text[0] == tsv(df_snv) and text[1] == tsv(df_insdel)


def test_gzip_members_roundtrip(tmp_path):
    import gzip
    from pav_b200.pavlib import flag
    data = b''.join(b'line %d\tsome text\n' % k for k in range(200_000))
    p = str(tmp_path / 'x.tsv.gz')
    flag.write_gzip_members(p, data, threads=3, block=1 << 18)
    assert gzip.open(p, 'rb').read() == data
    assert pd.read_csv(p, sep='\t', header=None).shape == (200_000, 2)
    flag.write_gzip_members(p, b'', threads=2)
    assert gzip.open(p, 'rb').read() == b''


def test_any_repeat_and_versioned_ids_many_records(tmp_path):
    """_any_repeat (the numeric test that decides whether the string pass of version_id is needed) against a set-based count, and
    version_id=True on a 60-record table (REF/ALT gathered per run of rows of a record) equal to the oracle's frames."""
    rng = np.random.default_rng(1)
    for _ in range(2000):
        n = int(rng.integers(0, 12))
        a, b = rng.integers(0, 4, n), rng.integers(0, 3, n)
        assert cigarcall._any_repeat(a, b) == (len(set(zip(a.tolist(), b.tolist()))) < n)
        assert cigarcall._any_repeat(a) == (len(set(a.tolist())) < n)
        s = np.sort(a)
        assert cigarcall._any_repeat(s, b) == (len(set(zip(s.tolist(), b.tolist()))) < n)
    ref, tigs, df = synth.make_cigar_workload(41, 3, 80_000, 60, 4_000, edit_rate=0.02, rev_frac=0.5)
    twin = df.iloc[[5, 17]].copy()
    twin['INDEX'] = [900, 901]
    df = pd.concat([df, twin], axis=0)
    ref_fa, tig_fa, _ = synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    for vid in (True, False):
        got = frames_from_oracle_rows(df, ref_fa, tig_fa, 'h1', vid)
        want = pyoracle.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=vid)
        assert tsv(got[0]) == tsv(want[0]) and tsv(got[1]) == tsv(want[1])
        assert (got[0].index == want[0].index).all() and (got[1].index == want[1].index).all()
    assert got[0].shape[0] > 1000
