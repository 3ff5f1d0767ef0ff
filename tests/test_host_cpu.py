"""CPU-only checks: the C-ABI library loads and exports every symbol of include/pavgpu.h, the host
tokenizer, and the host-side mirrors (FASTA reader, Region, rl_encoder, version_id, AlignLift, srs tree)."""
import json
import os
import re

import numpy as np
import pandas as pd
import pytest

from oracle import pyoracle, refenv
from pav_b200 import _capi, fasta, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def test_capi_exports_match_header():
    hdr = open(os.path.join(REPO, 'include', 'pavgpu.h')).read()
    declared = set(re.findall(r'\b(pavgpu_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    L = _capi.lib()
    for s in declared:
        assert hasattr(L, s), s


def test_struct_layouts_match_header():
    import ctypes
    assert _capi.SNV_ROW.itemsize == 16 and _capi.INDEL_ROW.itemsize == 64
    assert _capi.DENSITY_WINDOW.itemsize == 32 and _capi.DENSITY_RESULT.itemsize == 32
    assert ctypes.sizeof(_capi.DensityParams) == 32
    assert ctypes.sizeof(_capi.ParseErr) == 32 and ctypes.sizeof(_capi.CigarErr) == 32


def test_no_gpu_fails_loudly():
    """Without a CUDA device the hot path must raise, not fall back."""
    from pav_b200 import device
    if _capi.lib().pavgpu_device_count() > 0:
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        device.Context(0)
    from pav_b200.pavlib import cigarcall
    d = os.path.join(GOLDEN, 'cigar', 'kat1')
    df = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t')
    with pytest.raises(RuntimeError):
        cigarcall.make_insdel_snv_calls(df, os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), 'h1')


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(REPO, 'pav_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f


def test_cigar_parse_host():
    from pav_b200 import device
    from pav_b200.pavlib import align
    cigs = ['2H9=1X20=3I10=2D16=1H', '', '100=', '5M3N2P7S']
    ops, off, err = device.parse_cigars(cigs)
    assert err.code == 0 and off.tolist() == [0, 9, 9, 10, 14]
    chars = 'MIDNSHP=X'
    for i, c in enumerate(cigs):
        exp = list(align.cigar_str_to_tuples(c))
        got = [(int(o >> 4), chars[int(o & 15)]) for o in ops[off[i]:off[i + 1]]]
        assert got == exp
    for bad, code, tp in [('20==100=', 2, 3), ('20=5Q95=', 3, 3), ('120=5', 4, 5)]:
        ops, off, err = device.parse_cigars(['10=', bad, '7='])
        assert (err.code, err.rec, err.text_pos) == (code, 1, tp)
        assert off[3] == off[2]  # nothing after the malformed record


def test_cigar_parse_oversize_length():
    from pav_b200 import device
    ops, off, err = device.parse_cigars(['%d=' % (2 ** 28 - 1)])
    assert err.code == 0 and int(ops[0] >> 4) == 2 ** 28 - 1
    # 2^64 + 5 would wrap to a legal length if the digits were accumulated blindly
    for big in (2 ** 28, 2 ** 32 + 1, 2 ** 64 + 5, 10 ** 40):
        with pytest.raises(RuntimeError, match='exceeds 2\\^28-1'):
            device.parse_cigars(['10=', '%d=' % big])


def test_cigar_str_to_tuples_errors():
    from pav_b200.pavlib import align
    row = pd.Series({'CIGAR': '20=5Q95=', 'QRY_ID': 'tigA', '#CHROM': 'chrA', 'POS': 0})
    with pytest.raises(RuntimeError, match='Unknown CIGAR operation for contig tigA alignment starting at chrA:0: CIGAR operation 5'):
        list(align.cigar_str_to_tuples(row))
    row['CIGAR'] = '20==100='
    with pytest.raises(RuntimeError, match='Missing length in CIGAR string for contig tigA alignment starting at chrA:0: CIGAR index 3'):
        list(align.cigar_str_to_tuples(row))
    with pytest.raises(IndexError):
        list(align.cigar_str_to_tuples('120=5'))


def test_fasta_reader(tmp_path):
    rng = np.random.default_rng(2)
    seqs = {'a': synth.random_seq(rng, 1000), 'b b2': synth.random_seq(rng, 61), 'c': synth.random_seq(rng, 0), 'd': synth.random_seq(rng, 60)}
    seqs = {k.split()[0]: v for k, v in seqs.items()}
    p = synth.write_fasta(str(tmp_path / 'x.fa'), seqs, line_width=60)
    for with_fai in (True, False):
        if not with_fai:
            os.remove(p + '.fai')
        fa = fasta.Fasta(p)
        for n, s in seqs.items():
            assert fa.length(n) == len(s)
            assert (fa.fetch_array(n) == s).all()
            for a, b in [(0, 1), (59, 61), (60, 120), (5, 999)]:
                if b <= len(s):
                    assert fa.fetch(n, a, b) == s[a:b].tobytes().decode()
    import gzip
    with open(p, 'rb') as fi, gzip.open(p + '.gz', 'wb') as fo:
        fo.write(fi.read())
    fz = fasta.Fasta(p + '.gz')
    assert (fz.fetch_array('a') == seqs['a']).all() and fz.length('d') == 60
    assert (fasta.reverse_complement(np.frombuffer(b'ACGTNacgtRy', np.uint8)) == np.frombuffer(b'rYacgtNACGT', np.uint8)).all()


def test_region_and_expand():
    from pav_b200.pavlib import seq
    r = seq.region_from_string('chr1:1,001-2,000')
    assert (r.chrom, r.pos, r.end, r.is_rev, len(r)) == ('chr1', 1000, 2000, False, 1000)
    assert str(r) == 'chr1:1001-2000' and r.region_id() == 'chr1-1000-RGN-1000'
    assert seq.region_from_id('chr1-1001-RGN-1000') == r
    rr = seq.Region('t', 50, 10)
    assert (rr.pos, rr.end, rr.is_rev) == (10, 50, True)
    fai = pd.Series({'chr1': 5000})
    a = r.copy(); a.expand(4000, min_pos=0, max_end=fai, shift=True)
    assert (a.pos, a.end) == (0, 5000)
    b = r.copy(); b.expand(np.int32(1500), min_pos=0, max_end=fai, shift=True, balance=0.25)
    assert (b.pos, b.end) == (1000 - 375, 2000 + 1125)


@pytest.mark.skipif(not refenv.available(), reason='reference tree only exists in the build container')
def test_region_expand_matches_reference():
    refenv.activate()
    import pavlib as ref_pavlib
    from pav_b200.pavlib import seq
    rng = np.random.default_rng(5)
    fai = pd.Series({'c': 100000})
    for _ in range(300):
        p = int(rng.integers(0, 90000)); e = p + int(rng.integers(1, 9000))
        bp = int(rng.integers(0, 60000)); bal = float(rng.choice([0.25, 0.5, 0.75]))
        a = seq.Region('c', p, e); b = ref_pavlib.seq.Region('c', p, e)
        a.expand(np.int32(bp), min_pos=0, max_end=fai, shift=True, balance=bal)
        b.expand(np.int32(bp), min_pos=0, max_end=fai, shift=True, balance=bal)
        assert (a.pos, a.end) == (b.pos, b.end)


def test_rl_encoder_and_version_id():
    from pav_b200.pavlib import density, variant
    df = pd.DataFrame({'STATE': [0, 0, 2, 2, 2, 0, 1], 'INDEX': [3, 4, 9, 10, 11, 50, 51]})
    assert list(density.rl_encoder(df)) == [(0, 2, 3, 4), (2, 3, 9, 11), (0, 1, 50, 50), (1, 1, 51, 51)]
    assert list(density.rl_encoder(df)) == list(pyoracle.rl_encoder(df))
    assert list(density.rl_encoder(df.iloc[:0])) == []
    ids = pd.Series(['a', 'b', 'a', 'a.1', 'c', 'a', 'b'])
    assert variant.version_id(ids).tolist() == pyoracle.version_id(ids.tolist()) == ['a', 'b', 'a.2', 'a.1', 'c', 'a.3', 'b.1']
    same = pd.Series(['x', 'y'])
    assert variant.version_id(same) is same


def test_srs_tree():
    from pav_b200.pavlib import inv
    t = inv.get_srs_tree(None)
    assert list(t[12345])[0].data == 20
    t = inv.get_srs_tree([(1000, 10), (50000, 20), (500000, 40)])
    assert [list(t[x])[0].data for x in (5, 999, 1000, 49999, 50000, 10 ** 7)] == [20, 20, 10, 10, 20, 40]


def _lift_table():
    rng = np.random.default_rng(21)
    ref, tigs, df = synth.make_cigar_workload(21, 1, 120_000, 4, 30_000, edit_rate=0.01, rev_frac=0.5, clip=(13, 7))
    fai = pd.Series({k: len(v) for k, v in tigs.items()})
    return df.reset_index(drop=True), fai, rng


@pytest.mark.skipif(not refenv.available(), reason='reference tree only exists in the build container')
def test_align_lift_matches_reference():
    refenv.activate()
    import pavlib as ref_pavlib
    from pav_b200.pavlib import lift, seq
    df, fai, rng = _lift_table()
    mine, theirs = lift.AlignLift(df, fai), ref_pavlib.align.AlignLift(df, fai)
    for _ in range(400):
        row = df.iloc[int(rng.integers(0, df.shape[0]))]
        p = int(rng.integers(row['POS'] - 50, row['END'] + 50))
        assert mine.lift_to_qry(row['#CHROM'], p) == theirs.lift_to_qry(row['#CHROM'], p)
        q = int(rng.integers(max(row['QRY_POS'] - 30, 0), row['QRY_END'] + 30))
        try:
            exp = theirs.lift_to_sub(row['QRY_ID'], q)
        except RuntimeError:
            with pytest.raises(RuntimeError):
                mine.lift_to_sub(row['QRY_ID'], q)
            continue
        assert mine.lift_to_sub(row['QRY_ID'], q) == exp
    for _ in range(100):
        row = df.iloc[int(rng.integers(0, df.shape[0]))]
        a = int(rng.integers(row['POS'], row['END'] - 2000)); b = a + int(rng.integers(10, 1900))
        r1, r2 = mine.lift_region_to_qry(seq.Region(row['#CHROM'], a, b)), theirs.lift_region_to_qry(ref_pavlib.seq.Region(row['#CHROM'], a, b))
        assert (r1 is None) == (r2 is None)
        if r1 is not None:
            assert (r1.chrom, r1.pos, r1.end, r1.is_rev) == (r2.chrom, r2.pos, r2.end, r2.is_rev)
            b1, b2 = mine.lift_region_to_sub(r1), theirs.lift_region_to_sub(r2)
            assert (b1 is None) == (b2 is None)
            if b1 is not None:
                assert (b1.chrom, b1.pos, b1.end) == (b2.chrom, b2.pos, b2.end)


def test_align_lift_roundtrip():
    """Runs everywhere: ref -> contig -> ref returns the start position inside aligned blocks."""
    from pav_b200.pavlib import lift
    df, fai, rng = _lift_table()
    L = lift.AlignLift(df, fai)
    ok = 0
    for _ in range(200):
        row = df.iloc[int(rng.integers(0, df.shape[0]))]
        p = int(rng.integers(row['POS'] + 100, row['END'] - 100))
        q = L.lift_to_qry(row['#CHROM'], p)
        assert q is not None and q[0] == row['QRY_ID'] and q[2] == row['REV']
        back = L.lift_to_sub(q[0], q[1])
        if back is not None and abs(back[1] - p) <= 1:
            ok += 1
    assert ok > 150


def test_get_align_bed_matches_reference_golden():
    """SAM text -> alignment table equals the table produced by the reference's get_align_bed."""
    from pav_b200.pavlib import align, seq
    d = os.path.join(GOLDEN, 'align', 'sam1')
    meta = json.load(open(os.path.join(d, 'meta.json')))
    fai = seq.get_df_fai(os.path.join(d, 'tig.fa.fai'))
    df = align.get_align_bed(os.path.join(d, 'align.sam'), fai, 'h1', min_mapq=meta['min_mapq'])
    assert df.to_csv(sep='\t', index=False).encode() == open(os.path.join(d, 'align.bed'), 'rb').read()
    assert [int(i) for i in df.index] == meta['index'] and [str(t) for t in df.dtypes] == meta['dtypes']
    with pytest.raises(RuntimeError) as ei:
        align.get_align_bed(os.path.join(d, 'bad_m.sam'), fai, 'h1')
    assert str(ei.value) == meta['bad_m'][1]


def test_count_cigar_and_check_record():
    from pav_b200.pavlib import align
    assert align.count_cigar('3H2S10=2X4I5D7=1S2H') == (24, 23, 3, 2, 2, 1)
    with pytest.raises(RuntimeError, match='CIGAR op "M" is not allowed'):
        align.count_cigar('10M')
    with pytest.raises(RuntimeError, match='Found clipped bases before last non-clipped'):
        align.count_cigar('5=2S5=')
    row = pd.Series({'#CHROM': 'c', 'POS': 10, 'END': 34, 'INDEX': 0, 'QRY_ID': 'q', 'QRY_POS': 5, 'QRY_END': 28, 'QRY_LEN': 31,
                     'CIGAR': '3H2S10=2X4I5D7=1S2H'})
    align.check_record(row, pd.Series({'q': 31}))
    row['END'] = 35
    with pytest.raises(RuntimeError, match='END mismatch'):
        align.check_record(row, pd.Series({'q': 31}))


def test_fasta_bgzf_random_access(tmp_path):
    """bgzip-compressed FASTA + .fai (+/- .gzi): only the blocks of the requested record are inflated."""
    import gzip
    import shutil
    rng = np.random.default_rng(4)
    seqs = {'chrA': synth.random_seq(rng, 300_000), 'chrB': synth.random_seq(rng, 7), 'chrC': synth.random_seq(rng, 150_001)}
    plain = synth.write_fasta(str(tmp_path / 'g.fa'), seqs, line_width=60)
    data = open(plain, 'rb').read()
    for with_gzi in (True, False):
        gz = str(tmp_path / f'g{int(with_gzi)}.fa.gz')
        synth.write_bgzf(gz, data, write_gzi=with_gzi)
        shutil.copy(plain + '.fai', gz + '.fai')
        assert gzip.open(gz, 'rb').read() == data      # a valid multi-member gzip stream
        fa = fasta.Fasta(gz)
        assert fa._bgzf is not None and fa._buf is None
        for n, s in seqs.items():
            assert (fa.fetch_array(n) == s).all()
        assert fa.fetch('chrA', 65_000, 66_500) == seqs['chrA'][65_000:66_500].tobytes().decode()
        assert fa.fetch('chrC', 149_990, 150_001) == seqs['chrC'][149_990:].tobytes().decode()


def test_fasta_fetch_into(tmp_path):
    """Whole-record reads straight into a caller buffer (pinned staging): every line width, partial last lines, empty records,
    a file without a trailing newline, and the BGZF fallback."""
    import numpy as np

    from pav_b200 import fasta, synth
    rng = np.random.default_rng(1)
    seqs = {'a': synth.random_seq(rng, 1000), 'b': synth.random_seq(rng, 80), 'c': synth.random_seq(rng, 81), 'd': synth.random_seq(rng, 7),
            'e': np.zeros(0, np.uint8), 'f': synth.random_seq(rng, 160)}
    for lw in (80, 60, 7):
        p = str(tmp_path / f'x{lw}.fa')
        synth.write_fasta(p, seqs, line_width=lw)
        fa = fasta.open_fasta(p)
        for k, v in seqs.items():
            out = np.full(len(v) + 5, 255, np.uint8)
            got = fa.fetch_into(k, out)
            assert (got == v).all() and (out[len(v):] == 255).all(), (lw, k)
            assert (fa.fetch_array(k) == v).all()
    p = str(tmp_path / 'nonl.fa')
    open(p, 'wb').write(b'>s\nACGTACGT\nACGTACGT')
    out = np.zeros(16, np.uint8)
    assert bytes(fasta.open_fasta(p).fetch_into('s', out)) == b'ACGTACGTACGTACGT'
    # bgzip-compressed FASTA: block-wise reader behind the same call
    plain = str(tmp_path / 'x80.fa')
    gz = str(tmp_path / 'z.fa.gz')
    synth.write_bgzf(gz, open(plain, 'rb').read(), block=300)
    import shutil
    shutil.copy(plain + '.fai', gz + '.fai')
    fz = fasta.open_fasta(gz)
    for k, v in seqs.items():
        out = np.zeros(len(v), np.uint8)
        assert (fz.fetch_into(k, out) == v).all()
    import pytest
    with pytest.raises(KeyError):
        fa.fetch_into('nope', np.zeros(4, np.uint8))


def test_density_cli_host_side():
    """The parts of the scripts/density.py twin that need no GPU: argument errors raised before any work, and the text of the
    two soft failures, against what the reference's own process printed (tests/golden/density_cli.json)."""
    import json
    from pav_b200 import fasta
    from pav_b200.scripts import density as cli
    gold = json.load(open(os.path.join(REPO, 'tests', 'golden', 'density_cli.json')))
    d = os.path.join(REPO, 'tests', 'golden', 'density')
    base = ['--tigregion', 'tigW:1-1500', '--refregion', 'chrW:1-1500', '--ref', os.path.join(d, 'few_informative', 'ref.fa'),
            '--tig', os.path.join(d, 'few_informative', 'tig.fa')]
    with pytest.raises(RuntimeError) as e:
        cli.main(base + ['-r', 'maybe'])
    assert 'RuntimeError: ' + str(e.value) == gold['bad_bool']['stderr_last']
    with pytest.raises(RuntimeError) as e:
        cli.main(base + ['-r', 'F', 'TMP/out.csv'])
    assert 'RuntimeError: ' + str(e.value) == gold['bad_extension']['stderr_last']
    assert [cli.get_bool(s) for s in ('TRUE', 't', '1', 'False', 'f', '0')] == [True, True, True, False, False, False]
    ref = fasta.Fasta(os.path.join(d, 'exit125_repeat', 'ref.fa')).fetch_array('chrW')
    n, mer = cli.ref_kmer_failure(ref, 31)
    assert 'K-mer count exceeds max: {} > {} ({}): {}\n'.format(n, cli.MAX_REF_KMER_COUNT, mer, 'chrW:1-5200') == gold['exit125_repeat']['stderr']
    ref = fasta.Fasta(os.path.join(d, 'exit125_empty', 'ref.fa')).fetch_array('chrW')
    assert cli.ref_kmer_failure(ref, 31) is None and cli.ref_kmer_failure(ref[:10], 31) is None


def test_count_cigar_matches_reference_golden():
    """count_cigar (vectorised fast path and op-by-op path) on 2,500 random CIGAR strings, legal and not: the tuple or the
    exception class + text of the reference's own pavlib.align.count_cigar (tests/golden/count_cigar.json)."""
    import json
    from pav_b200.pavlib import align
    cases = json.load(open(os.path.join(REPO, 'tests', 'golden', 'count_cigar.json')))
    n_fast = 0
    for c in cases:
        row = pd.Series({'CIGAR': c['cigar'], 'QRY_ID': 'tigA', '#CHROM': 'chrA', 'POS': 5})
        n_fast += align._count_cigar_fast(c['cigar'], c['allow_m']) is not None
        try:
            got = {'result': [int(x) for x in align.count_cigar(row, allow_m=c['allow_m'])]}
        except Exception as ex:   # noqa: BLE001
            got = {'error': [type(ex).__name__, str(ex)]}
        want = {k: c[k] for k in ('result', 'error') if k in c}
        assert got == want, c['cigar']
    assert n_fast > 500      # the fast path really took part


def test_align_lift_matches_reference_golden():
    """AlignLift against the answers of the reference's own pavlib.align.AlignLift stored with the table (tests/golden/lift: 3,000
    point lifts in both directions incl. gap=True, 400 region lifts there and back)."""
    import json
    from pav_b200.pavlib import lift, seq
    d = os.path.join(REPO, 'tests', 'golden', 'lift')
    df = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t')
    fai = pd.read_csv(os.path.join(d, 'tig.fai.tsv'), sep='\t', header=None, index_col=0)[1]
    al = lift.AlignLift(df, fai)

    def plain(x):
        return None if x is None else [x[0], int(x[1]), bool(x[2]), int(x[3]), int(x[4]), [int(i) for i in x[5]]]
    n = 0
    for q in json.load(open(os.path.join(d, 'queries.json'))):
        if q['f'] == 'region':
            rq = al.lift_region_to_qry(seq.Region(q['chrom'], q['pos'], q['end']))
            assert (None if rq is None else [rq.chrom, rq.pos, rq.end, bool(rq.is_rev)]) == q['qry'], q
            if rq is not None:
                for gap, key in ((False, 'sub'), (True, 'sub_gap')):
                    rs = al.lift_region_to_sub(rq, gap=gap)
                    assert (None if rs is None else [rs.chrom, rs.pos, rs.end, bool(rs.is_rev)]) == q[key], (q, gap)
        else:
            f = al.lift_to_qry if q['f'] == 'to_qry' else al.lift_to_sub
            kw = {'gap': q['gap']} if q['f'] == 'to_sub' else {}
            if 'error' in q:
                with pytest.raises(RuntimeError):
                    f(q['id'], q['pos'], **kw)
            else:
                assert plain(f(q['id'], q['pos'], **kw)) == q['result'], q
        n += 1
    assert n == 3400


def test_region_expand_matches_reference_golden():
    """Region.expand against stored results of the reference's Region.expand (tests/golden/region_expand.json)."""
    import json
    from pav_b200.pavlib import seq
    fai = pd.Series({'c': 100000})
    for c in json.load(open(os.path.join(REPO, 'tests', 'golden', 'region_expand.json'))):
        r = seq.Region('c', c['pos'], c['end'])
        r.expand(c['bp'], min_pos=c['min_pos'], max_end={'fai': fai, 'int': 100000, 'none': None}[c['max_end']], shift=c['shift'], balance=c['balance'])
        assert [r.pos, r.end] == c['result'], c
        r = seq.Region('c', c['pos'], c['end'])
        r.expand(np.int32(c['bp']), min_pos=c['min_pos'], max_end={'fai': fai, 'int': 100000, 'none': None}[c['max_end']], shift=c['shift'], balance=c['balance'])
        assert [r.pos, r.end] == c['result'], c


def test_binding_layouts_match_header(tmp_path):
    """Every struct of include/pavgpu.h as the C compiler lays it out (sizeof + offsetof of each field, from a program compiled
    against the header) equals the ctypes Structure / numpy dtype the binding uses for it."""
    import ctypes
    import subprocess
    from pav_b200 import _capi
    bound = {
        'pavgpu_parse_err': _capi.ParseErr, 'pavgpu_cigar_err': _capi.CigarErr, 'pavgpu_cigar_stats': _capi.CigarStats,
        'pavgpu_density_params': _capi.DensityParams, 'pavgpu_density_stats': _capi.DensityStats,
        'pavgpu_snv_row': _capi.SNV_ROW, 'pavgpu_indel_row': _capi.INDEL_ROW, 'pavgpu_density_window': _capi.DENSITY_WINDOW,
        'pavgpu_density_result': _capi.DENSITY_RESULT, 'pavgpu_state_run': _capi.STATE_RUN, 'pavgpu_cigar_rec_stats': _capi.CIGAR_REC_STATS,
    }

    def fields(b):
        if isinstance(b, np.dtype):
            return [(n, b.fields[n][1]) for n in b.names], b.itemsize
        return [(n, getattr(b, n).offset) for n, _ in b._fields_], ctypes.sizeof(b)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pavgpu.h"', 'int main(void) {']
    for name, b in bound.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for f, _ in fields(b)[0]:
            lines.append(f'  printf("{name} {f} %zu\\n", offsetof({name}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = str(tmp_path / 'layout')
    subprocess.check_call(['gcc', '-std=c11', '-I', os.path.join(REPO, 'include'), '-o', exe, str(src)])
    got = {}
    for ln in subprocess.check_output([exe]).decode().split('\n'):
        if ln:
            s, f, v = ln.split()
            got[(s, f)] = int(v)
    for name, b in bound.items():
        fl, size = fields(b)
        assert got[(name, 'size')] == size, name
        for f, off in fl:
            assert got[(name, f)] == off, (name, f)
    # every struct typedef of the header is bound (a new struct must come with its binding)
    import re
    hdr = open(os.path.join(REPO, 'include', 'pavgpu.h')).read()
    assert set(re.findall(r'^\} (pavgpu_\w+);', hdr, flags=re.M)) == set(bound)


def test_binding_arity_matches_header():
    """Every prototype of include/pavgpu.h has as many parameters as the argtypes the binding declares for it."""
    import re
    from pav_b200 import _capi
    L = _capi.lib()
    hdr = re.sub(r'/\*.*?\*/', '', open(os.path.join(REPO, 'include', 'pavgpu.h')).read(), flags=re.S)
    protos = re.findall(r'\b(pavgpu_\w+)\s*\(([^;{]*?)\)\s*;', hdr)
    assert {n for n, _ in protos} == set(_capi.EXPORTS)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ('', 'void') else params.count(',') + 1
        fn = getattr(L, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, (name, n, len(fn.argtypes))
        else:
            assert n == 0 or name in ('pavgpu_device_count', 'pavgpu_last_error'), name


def test_read_sequences_splits_large_records_across_jobs(tmp_path):
    """cigarcall.read_sequences: large records are copied as runs of lines by several jobs, small ones grouped; every layout of the
    last line (full, partial, no trailing newline) gives the bases of the plain reader."""
    from concurrent.futures import ThreadPoolExecutor

    from pav_b200 import fasta, synth
    from pav_b200.pavlib import cigarcall
    rng = np.random.default_rng(17)
    seqs = {f'c{i}': synth.random_seq(rng, n) for i, n in enumerate([9_000_123, 2_400_000, 61, 60, 59, 1, 4_345_678, 7, 3_000_060])}
    p = synth.write_fasta(str(tmp_path / 'x.fa'), seqs)
    for strip_newline in (False, True):
        if strip_newline:
            raw = open(p, 'rb').read().rstrip(b'\n')
            p = str(tmp_path / 'y.fa')
            open(p, 'wb').write(raw)
            import subprocess
            idx = fasta.Fasta(p)     # builds y.fa.fai
            assert idx.length('c8') == 3_000_060
        fa = fasta.open_fasta(p)
        with ThreadPoolExecutor(6) as pool:
            out, futs = cigarcall.read_sequences(fa, list(seqs), pool)
            for f in futs:
                f.result()
        assert len(futs) > len(seqs) // 2
        for nm, a in zip(seqs, out):
            assert np.array_equal(a, seqs[nm]), (nm, strip_newline)


def test_record_stats_host_agree_with_count_cigar():
    """The per-record sums get_align_bed uses (align._record_stats_host, the host twin of pavgpu_cigar_record_stats) against
    count_cigar on well-formed records: aligned spans and clips must tell the same story."""
    from pav_b200 import device
    from pav_b200.pavlib import align
    rng = np.random.default_rng(8)
    cigars = []
    for _ in range(400):
        body = ''.join('%d%s' % (int(rng.integers(1, 3000)), '=XID'[int(rng.integers(0, 4))]) for _ in range(int(rng.integers(1, 60))))
        body = '7=' + body + '3='          # (aligned ends, so that clips sit next to aligned bases)
        lead = ['', '5H', '7S', '3H9S'][int(rng.integers(0, 4))]
        tail = ['', '4S', '6H', '8S1H'][int(rng.integers(0, 4))]
        cigars.append(lead + body + tail)
    ops, op_off, perr = device.parse_cigars(cigars)
    assert perr.code == 0
    st = align._record_stats_host((ops & 15).astype(np.int64), (ops >> 4).astype(np.int64), op_off)
    for c, s_ in zip(cigars, st):
        ref_bp, tig_bp, h_l, s_l, h_r, s_r = align.count_cigar(c)
        assert (int(s_['ref_bp']), int(s_['qry_bp'])) == (ref_bp, tig_bp), c
        assert int(s_['lead']) == h_l + s_l and int(s_['trail']) == h_r + s_r and int(s_['clip_h_first']) == h_l and int(s_['lead_s']) == s_l, c
        assert int(s_['flags']) == 0 and s_['first_body'] >= 0
