#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(/root/reference, PAV 2.4.6.0) in the build container.

The reference ships no tests and no golden vectors, so every pin is produced here by
the reference's own code: pavlib.cigarcall.make_insdel_snv_calls (pavlib/cigarcall.py:24),
pavlib.call.left_homology/right_homology (pavlib/call.py:542,595), scripts/density.py (spawned
exactly like pavlib/inv.py:249-266 does) and pavlib.inv.scan_for_inv (pavlib/inv.py:149), with the
third-party modules that are absent from this image replaced by oracle/ref_stubs/.

Run:  python tests/golden/make_golden.py            (container only; needs /root/reference)
Outputs are small and committed; inputs are committed with them so the fixtures are
self-contained on the GPU box (where /root/reference does not exist).
"""
import base64
import codecs
import gzip
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refenv  # noqa: E402
from pav_b200 import synth  # noqa: E402

refenv.activate()
import pavlib  # noqa: E402  (the reference)
import kanapy.util.kmer  # noqa: E402


def _write_fa(path, seqs):
    """name -> str/bytes/uint8 array; plain FASTA + .fai, gzip-free (small)."""
    arrs = {}
    for k, v in seqs.items():
        if isinstance(v, str):
            v = np.frombuffer(v.encode(), dtype=np.uint8)
        arrs[k] = v
    synth.write_fasta(path, arrs, line_width=60)


def _save_df(df, path):
    df.to_csv(path, sep='\t', index=False)


def _align_df(rows):
    return pd.DataFrame(rows, columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END',
                                       'QRY_LEN', 'REV', 'CIGAR'])


def cigar_case(name, ref, tigs, df_align, hap='h1', version_id=True):
    """Run reference Path A on one case; store inputs + outputs (or the exception)."""
    d = os.path.join(HERE, 'cigar', name)
    os.makedirs(d, exist_ok=True)
    ref_fa = os.path.join(d, 'ref.fa')
    tig_fa = os.path.join(d, 'tig.fa')
    _write_fa(ref_fa, ref)
    _write_fa(tig_fa, tigs)
    df_align.to_csv(os.path.join(d, 'align.bed'), sep='\t', index=False)
    meta = {'hap': hap, 'version_id': version_id}
    try:
        df_snv, df_insdel = pavlib.cigarcall.make_insdel_snv_calls(df_align, ref_fa, tig_fa, hap, version_id=version_id)
        _save_df(df_snv, os.path.join(d, 'snv.tsv'))
        _save_df(df_insdel, os.path.join(d, 'insdel.tsv'))
        meta['n_snv'] = int(df_snv.shape[0])
        meta['n_insdel'] = int(df_insdel.shape[0])
        meta['snv_index'] = [int(i) for i in df_snv.index]
        meta['insdel_index'] = [int(i) for i in df_insdel.index]
        meta['snv_dtypes'] = [str(t) for t in df_snv.dtypes]
        meta['insdel_dtypes'] = [str(t) for t in df_insdel.dtypes]
    except Exception as ex:  # noqa: BLE001
        meta['exception'] = type(ex).__name__
        meta['message'] = str(ex)
    with open(os.path.join(d, 'meta.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)
    print('cigar', name, {k: v for k, v in meta.items() if k in ('n_snv', 'n_insdel', 'exception')})


def make_cigar_cases():
    # --- KAT 1: lowercase prefix, N, fwd + rev record, version_id=True
    ref1 = 'ttgaccgtACGATTACAGCAGCAGCAGCAGTTGACCTGANCCGTAGGCTTAAGGCCTA'
    q = list(ref1)
    q[9] = 't'
    qs = ''.join(q)
    qs = qs[:30] + 'CAG' + qs[30:40] + qs[42:]
    contig = 'GG' + qs + 'A'
    cig = '2H9=1X20=3I10=2D16=1H'
    rc = str(__import__('Bio').Seq.Seq(contig).reverse_complement())
    df = _align_df([
        ('chr1', 0, 58, 0, 'tigF', 2, 61, 62, False, cig),
        ('chr1', 0, 58, 1, 'tigR', 1, 60, 62, True, cig),
    ])
    cigar_case('kat1', {'chr1': ref1}, {'tigF': contig, 'tigR': rc}, df, version_id=True)

    # --- KAT 2: DEL left-shift quirk
    ref2 = 'ACGTACGTTTGACCAGTAGGGGGGCATCATCATCAGTTTACGATCGGATCAGCTAGCAAGT'
    # 5= 1X 21= 3D 13= 2I rest=
    q2 = ref2[:5] + 'A' + ref2[6:27] + ref2[30:43] + 'GG' + ref2[43:]
    cig2 = f'5=1X21=3D13=2I{len(ref2) - 43}='
    df = _align_df([('chr1', 0, len(ref2), 0, 'tig1', 0, len(q2), len(q2), False, cig2)])
    cigar_case('kat2', {'chr1': ref2}, {'tig1': q2}, df, version_id=True)

    # --- C1 (BASELINE configs[0])
    ref, tigs, dfa = synth.config_c1()
    cigar_case('c1', ref, tigs, dfa, version_id=False)
    cigar_case('c1_vid', ref, tigs, dfa, version_id=True)

    # --- multi-record: REV, clips, soft-mask, N blocks, two chromosomes, duplicate-ID producing overlap
    ref, tigs, dfa = synth.make_cigar_workload(77, 2, 60_000, 6, 20_000, edit_rate=0.012, rev_frac=0.5,
                                               clip=(5, 7), soft_mask_frac=0.4, n_block_frac=0.03)
    # a second haplotype-like copy of record 0 so that IDs collide (exercises version_id)
    dup = dfa.iloc[[0]].copy()
    dup['INDEX'] = 100
    dfa2 = pd.concat([dfa, dup], axis=0)
    cigar_case('multi', ref, tigs, dfa2, hap='h2', version_id=True)
    cigar_case('multi_novid', ref, tigs, dfa, hap='h2', version_id=False)

    # --- edge cases, hand built
    rng = np.random.default_rng(5)
    base = ''.join('ACGT'[i] for i in rng.integers(0, 4, 300))
    # (a) record starting at POS=0 with I then D as the first ops; INS at the contig end; X run;
    #     IUPAC / lowercase / N bases in the query under X and inside an insertion
    refe = 'ACACACACACAC' + base[:200] + 'GTGTGTGTGT'
    body = refe
    qe = 'AC' + body[:12] + body[12:50] + 'nRy' + body[53:100] + 'tttNNN' + body[100:150] + body[156:] + 'GT'
    cige = f'2I50=3X47=6I50=6D{len(refe) - 156}=2I'
    df = _align_df([('chrE', 0, len(refe), 3, 'tigE', 0, len(qe), len(qe), False, cige)])
    cigar_case('edge_ends', {'chrE': refe}, {'tigE': qe}, df, version_id=False)
    # (b) D as the first op at POS=0, S clips, reverse strand, homopolymers
    refh = 'AAAAAAAAAAAAAAAAAAAACCCCCCCCCCCCCCCCCCCCGGGGGGGGGGTTTTTTTTTTTTTTTTTTTT' + base[:60]
    qh_ref = 'NN' + refh[3:30] + 'CCCC' + refh[30:55] + refh[57:]
    cigh = f'2S3D27=4I25=2D{len(refh) - 57}='
    qh = str(__import__('Bio').Seq.Seq(qh_ref).reverse_complement())
    df = _align_df([('chrH', 0, len(refh), 9, 'tigH', 0, len(qh) - 2, len(qh), True, cigh)])
    cigar_case('edge_homopolymer_rev', {'chrH': refh}, {'tigH': qh}, df, version_id=False)
    # (c) adjacent indels / X (last_op not '=') and zero-length corner
    refa = base[:120]
    qa = refa[:20] + 'TT' + refa[23:40] + 'G' + refa[45:60] + ('A' if refa[60] != 'A' else 'C') + 'CC' + refa[61:]
    ciga = f'20=2I3D17=1I5D15=1X2I{len(refa) - 61}='
    df = _align_df([('chrA', 10, 10 + len(refa), 4, 'tigA', 0, len(qa), len(qa), False, ciga)])
    cigar_case('edge_adjacent', {'chrA': 'GATTACAGAT' + refa}, {'tigA': qa}, df, version_id=False)
    # (d) empty table and a record with no variants
    df0 = _align_df([])
    cigar_case('empty', {'chrA': refa}, {'tigA': refa}, df0, version_id=True)
    df = _align_df([('chrA', 0, len(refa), 0, 'tigA', 0, len(refa), len(refa), False, f'{len(refa)}=')])
    cigar_case('novariants', {'chrA': refa}, {'tigA': refa}, df, version_id=True)

    # --- error behaviour (messages must match)
    for nm, cg in [('err_M', '20=5M95='), ('err_N', '20=5N95='), ('err_P', '10=1X2P109='),
                   ('err_nolen', '20==100='), ('err_badop', '20=5Q95='), ('err_trailing_digits', '120=5')]:
        df = _align_df([('chrA', 0, 120, 7, 'tigA', 0, 120, 120, False, f'{len(refa)}='),
                        ('chrA', 0, 120, 8, 'tigA', 0, 120, 120, False, cg)])
        cigar_case(nm, {'chrA': refa}, {'tigA': refa}, df, version_id=False)


def homology_cases():
    """Direct known answers for left_homology / right_homology (pavlib/call.py:542-647)."""
    rng = np.random.default_rng(11)
    out = []
    seqs = ['ACGTACGTACGTACGT', 'AAAAAAAAAA', 'ACGNNACGTTTTACAC', 'CAGCAGCAGCAGCAGTTGA', 'A', 'GATTACA' * 6]
    svs = ['ACGT', 'A', 'CAG', 'AC', 'TTGA', 'GATTACA', 'N', 'ACGTACGTACGTACGTAC']
    for s in seqs:
        for v in svs:
            for p in list(range(-1, len(s) + 1)):
                lh = pavlib.call.left_homology(p, s, v) if p < len(s) else None
                rh = pavlib.call.right_homology(p, s, v) if p >= 0 else None
                out.append({'seq': s, 'sv': v, 'pos': p, 'left': lh, 'right': rh})
    for _ in range(200):
        n = int(rng.integers(1, 40))
        s = ''.join('ACGTN'[i] for i in rng.choice(5, n, p=[.3, .3, .18, .18, .04]))
        m = int(rng.integers(1, 6))
        v = ''.join('ACGT'[i] for i in rng.integers(0, 2, m))
        p = int(rng.integers(0, n))
        out.append({'seq': s, 'sv': v, 'pos': p, 'left': pavlib.call.left_homology(p, s, v),
                    'right': pavlib.call.right_homology(p, s, v)})
    with open(os.path.join(HERE, 'homology.json'), 'w') as fh:
        json.dump(out, fh)
    print('homology', len(out))


def kmer_cases():
    """kanapy k-mer stream / rev_complement known answers (dep/svpop/dep/kanapy/util/kmer.py)."""
    rng = np.random.default_rng(13)
    out = []
    for k in (5, 16, 31, 32):
        ku = kanapy.util.kmer.KmerUtil(k)
        for n in (0, k - 1, k, 3 * k + 7, 200):
            s = ''.join('ACGTacgtNn'[i] for i in rng.choice(10, n, p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02]))
            st = list(kanapy.util.kmer.stream(s, ku, index=True))
            out.append({'k': k, 'seq': s, 'kmers': [str(a) for a, _ in st], 'index': [int(b) for _, b in st],
                        'rc': [str(ku.rev_complement(a)) for a, _ in st],
                        'canon': [str(ku.canonical_complement(a)) for a, _ in st]})
    with open(os.path.join(HERE, 'kmer.json'), 'w') as fh:
        json.dump(out, fh)
    print('kmer', len(out))


def _run_density(ref_fa, tig_fa, refregion, tigregion, k=31, rev=False, srs=20, extra=()):
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join(refenv.pythonpath_entries())
    args = [sys.executable, os.path.join(refenv.REF_ROOT, 'scripts', 'density.py'),
            '--tigregion', tigregion, '--refregion', refregion, '--ref', ref_fa, '--tig', tig_fa,
            '-k', str(k), '-t', '1', '-r', 'true' if rev else 'false', '--staterunsmooth', str(srs)] + list(extra)
    proc = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    if proc.returncode != 0:
        return proc.returncode, None, proc.stderr.decode()
    return 0, pickle.loads(codecs.decode(proc.stdout, 'base64')), proc.stderr.decode()


def density_case(name, ref_arr, tig_arr, k=31, rev=False, srs=20, sub=None, extra=()):
    d = os.path.join(HERE, 'density', name)
    os.makedirs(d, exist_ok=True)
    ref_fa, tig_fa = os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa')
    _write_fa(ref_fa, {'chrW': ref_arr})
    _write_fa(tig_fa, {'tigW': tig_arr})
    rs, re_ = (0, len(ref_arr)) if sub is None else sub[0]
    ts, te = (0, len(tig_arr)) if sub is None else sub[1]
    refregion, tigregion = f'chrW:{rs + 1}-{re_}', f'tigW:{ts + 1}-{te}'
    rc, df, err = _run_density(ref_fa, tig_fa, refregion, tigregion, k, rev, srs, extra)
    meta = {'k': k, 'rev': rev, 'srs': srs, 'refregion': refregion, 'tigregion': tigregion,
            'returncode': rc, 'extra': list(extra)}
    if df is not None:
        meta['columns'] = list(df.columns)
        meta['n_rows'] = int(df.shape[0])
        meta['rl_state'] = [[int(x) for x in r] for r in pavlib.density.rl_encoder(df)]
        meta['dtypes'] = [str(t) for t in df.dtypes]
        with gzip.open(os.path.join(d, 'density.tsv.gz'), 'wt') as fh:
            df.to_csv(fh, sep='\t', index=False)  # floats written with repr => round-trip exact
    with open(os.path.join(d, 'meta.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)
    print('density', name, rc, None if df is None else (df.shape, meta['rl_state'][:6]))


def make_density_cases():
    import random
    # KAT 4: random.seed(1), 20 kbp, inversion of [8000,12000)
    random.seed(1)
    s = ''.join(random.choice('ACGT') for _ in range(20000))
    comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
    t = s[:8000] + ''.join(comp[c] for c in reversed(s[8000:12000])) + s[12000:]
    density_case('kat4', np.frombuffer(s.encode(), np.uint8), np.frombuffer(t.encode(), np.uint8))

    rng = np.random.default_rng(2024)
    # inverted-repeat flanks + divergence => FWDREV states, many state changes
    r, t, _ = synth.make_inv_window(rng, 12000, 3000, flank_rep=800, divergence=0.004)
    density_case('flankrep_div', r, t)
    # same shape scored with -r true (reference k-mer set reverse-complemented)
    density_case('flankrep_div_rtrue', r, t, rev=True)
    # N run and lowercase in the contig, srs=10, sub-region of the records
    r, t, _ = synth.make_inv_window(rng, 10000, 2500, divergence=0.002, n_run=150)
    t = t.copy()
    t[4000:4600] |= 0x20
    density_case('nrun_lower_sub', r, t, srs=10, sub=((500, 9500), (700, 9400)))
    # negative control (no inversion): single FWD run
    r, t, _ = synth.make_inv_window(rng, 6000, 1000, negative=True, divergence=0.003)
    density_case('negative', r, t)
    # fewer than 2000 informative k-mers => un-smoothed frame (STATE=-1)
    r, t, _ = synth.make_inv_window(rng, 9000, 800, divergence=0.0)
    density_case('few_informative', r[:1500], t[:1500])
    # a low-count state (<20 k-mers) that must be dropped: tiny inverted segment of 40 bp
    r = synth.random_seq(rng, 7000)
    t = r.copy()
    t[3000:3040] = synth.revcomp(r[3000:3040])
    density_case('lowcount_state', r, t)
    # small isolated REV cluster just above the min state count (narrow bandwidth, density near/above 1)
    t = r.copy()
    t[3000:3060] = synth.revcomp(r[3000:3060])
    density_case('small_rev_cluster', r, t)
    # repeated reference k-mer > 100 copies => exit 125
    r = np.concatenate([synth.random_seq(rng, 2000), np.tile(np.frombuffer(b'ACGTTGCA', np.uint8), 150),
                        synth.random_seq(rng, 2000)])
    density_case('exit125_repeat', r, r.copy())
    # no reference k-mers at all (all N) => exit 125
    density_case('exit125_empty', np.full(500, ord('N'), np.uint8), synth.random_seq(rng, 500))
    # k = 21, srs = 7
    r, t, _ = synth.make_inv_window(rng, 8000, 2000, flank_rep=300, divergence=0.003)
    density_case('k21_srs7', r, t, k=21, srs=7)


def make_inv_case(name='kat3', rev=False, seed=3, flank_rep=0):
    """KAT 3: full pavlib.inv.scan_for_inv incl. AlignLift, 60 kbp, 8 kbp inversion.
    ``rev``: the contig is stored reverse-complemented and aligned on the minus strand (exercises -r true and the
    reverse lift); ``flank_rep``: inverted-repeat flanks (inner != outer breakpoints, FLANK / MATCH annotation)."""
    import random
    random.seed(seed)
    n = 60000
    s = ''.join(random.choice('ACGT') for _ in range(n))
    comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
    rc = lambda x: ''.join(comp[c] for c in reversed(x))  # noqa: E731
    if flank_rep:
        s = s[:34000] + rc(s[26000 - flank_rep:26000]) + s[34000 + flank_rep:]
    t = s[:26000] + rc(s[26000:34000]) + s[34000:]
    if rev:
        t = rc(t)
    d = os.path.join(HERE, 'inv', name)
    os.makedirs(d, exist_ok=True)
    ref_fa, tig_fa = os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa')
    _write_fa(ref_fa, {'chr1': s})
    _write_fa(tig_fa, {'tig1': t})
    df_align = pd.DataFrame([('chr1', 0, n, 0, 'tig1', 0, n, n, rev, f'{n}=')],
                            columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END', 'QRY_LEN',
                                     'REV', 'CIGAR'])
    df_align.to_csv(os.path.join(d, 'align.bed'), sep='\t', index=False)
    import svpoplib
    df_fai = svpoplib.ref.get_df_fai(tig_fa + '.fai')
    lift = pavlib.align.AlignLift(df_align, df_fai)
    k_util = kanapy.util.kmer.KmerUtil(31)
    flag = pavlib.seq.Region('chr1', 28000, 32000)
    # scan_for_inv spawns `python3 scripts/density.py`; give the child the stub path too
    os.environ['PYTHONPATH'] = os.pathsep.join(refenv.pythonpath_entries())
    log_path = os.path.join(d, 'scan.log')
    with open(log_path, 'w') as log:
        call = pavlib.inv.scan_for_inv(flag, ref_fa, tig_fa, lift, k_util, log=log)
    meta = {'flag': 'chr1:28001-32000', 'id': call.id, 'svlen': int(call.svlen),
            'region_ref_outer': str(call.region_ref_outer), 'region_ref_inner': str(call.region_ref_inner),
            'region_tig_outer': str(call.region_tig_outer), 'region_tig_inner': str(call.region_tig_inner),
            'region_ref_discovery': str(call.region_ref_discovery),
            'region_tig_discovery': str(call.region_tig_discovery),
            'df_columns': list(call.df.columns), 'df_rows': int(call.df.shape[0])}
    with gzip.open(os.path.join(d, 'density.tsv.gz'), 'wt') as fh:
        call.df.to_csv(fh, sep='\t', index=False)
    with open(os.path.join(d, 'meta.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)
    print('inv', name, meta['id'], meta['region_ref_outer'], meta['region_ref_inner'], meta['region_tig_outer'])


def make_align_case():
    """SAM text -> alignment table with the reference's get_align_bed (pavlib/align/align.py:666-794) on the stub SAM reader."""
    import svpoplib
    d = os.path.join(HERE, 'align', 'sam1')
    os.makedirs(d, exist_ok=True)
    ref, tigs, dfa = synth.make_cigar_workload(55, 2, 40_000, 7, 9_000, edit_rate=0.01, rev_frac=0.5, clip=(6, 9))
    tig_fa = os.path.join(d, 'tig.fa')
    _write_fa(tig_fa, tigs)
    extra = ['tigU\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*',                                   # unmapped
             'tig00001\t0\tchr1\t101\t3\t50=\t*\t0\t0\t*\t*',                          # below min_mapq
             'tig00002\t0\tchr1\t201\t60\t*\t*\t0\t0\t*\t*']                           # no CIGAR
    sam = synth.write_sam(os.path.join(d, 'align.sam'), dfa, ref, extra_lines=extra)
    fai = svpoplib.ref.get_df_fai(tig_fa + '.fai')
    df = pavlib.align.get_align_bed(sam, fai, 'h1', min_mapq=10)
    df.to_csv(os.path.join(d, 'align.bed'), sep='\t', index=False)
    meta = {'n': int(df.shape[0]), 'index': [int(i) for i in df.index], 'dtypes': [str(t) for t in df.dtypes], 'min_mapq': 10}
    # error: M operation
    with open(os.path.join(d, 'bad_m.sam'), 'w') as fh:
        fh.write('@HD\tVN:1.6\n')
        fh.write('tig00000\t0\tchr1\t1\t60\t100M\t*\t0\t0\t*\t*\n')
    try:
        pavlib.align.get_align_bed(os.path.join(d, 'bad_m.sam'), fai, 'h1')
    except Exception as ex:  # noqa: BLE001
        meta['bad_m'] = [type(ex).__name__, str(ex)]
    with open(os.path.join(d, 'meta.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)
    print('align sam1', meta['n'], meta.get('bad_m'))


def make_density_cli():
    """What the reference's scripts/density.py process itself says on the existing density cases: exit code, the text it
    prints for the two soft failures (one on stdout, one on stderr, scripts/density.py:510-527), the pickle's frame layout
    on stdout, and a .tsv written through the positional outfile argument."""
    import tempfile
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join(refenv.pythonpath_entries())
    out = {}
    for case in ('exit125_repeat', 'exit125_empty', 'few_informative', 'kat4'):
        d = os.path.join(HERE, 'density', case)
        meta = json.load(open(os.path.join(d, 'meta.json')))
        args = [sys.executable, os.path.join(refenv.REF_ROOT, 'scripts', 'density.py'),
                '--tigregion', meta['tigregion'], '--refregion', meta['refregion'], '--ref', 'ref.fa', '--tig', 'tig.fa',
                '-k', str(meta['k']), '-t', '1', '-r', 'true' if meta['rev'] else 'false', '--staterunsmooth', str(meta['srs'])]
        proc = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, cwd=d)
        rec = {'returncode': proc.returncode, 'stderr': proc.stderr.decode()}
        if proc.returncode == 0:
            df = pickle.loads(codecs.decode(proc.stdout, 'base64'))
            rec['index_name'] = df.index.name
            rec['index_equals_INDEX'] = bool((df.index.to_numpy() == df['INDEX'].to_numpy()).all())
            rec['columns'] = list(df.columns)
        else:
            rec['stdout'] = proc.stdout.decode()
        out[case] = rec
    d = os.path.join(HERE, 'density', 'few_informative')
    meta = json.load(open(os.path.join(d, 'meta.json')))
    with tempfile.TemporaryDirectory() as tmp:
        tsv = os.path.join(tmp, 'out.tsv')
        args = [sys.executable, os.path.join(refenv.REF_ROOT, 'scripts', 'density.py'),
                '--tigregion', meta['tigregion'], '--refregion', meta['refregion'], '--ref', 'ref.fa', '--tig', 'tig.fa',
                '-k', str(meta['k']), '-r', 'F', tsv]
        proc = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, cwd=d)
        out['few_informative_tsv'] = {'returncode': proc.returncode, 'stdout': proc.stdout.decode(), 'tsv': open(tsv).read()}
        proc = subprocess.run(args[:-1] + [os.path.join(tmp, 'out.csv')], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, cwd=d)
        out['bad_extension'] = {'returncode': proc.returncode, 'stderr_last': proc.stderr.decode().strip().splitlines()[-1].replace(tmp, 'TMP')}
        proc = subprocess.run(args[:-3] + ['-r', 'maybe'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, cwd=d)
        out['bad_bool'] = {'returncode': proc.returncode, 'stderr_last': proc.stderr.decode().strip().splitlines()[-1]}
    with open(os.path.join(HERE, 'density_cli.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
    for k, v in out.items():
        print('density cli', k, {a: (b if len(str(b)) < 100 else str(b)[:100] + '...') for a, b in v.items()})


def lift_cases(seed=21):
    """pavlib.align.AlignLift of the reference on a 4-record alignment table with clips, both strands and 1 % edits: point
    lifts in both directions (inside, at and beyond record ends, inside insertions / deletions, with and without gap=True) and
    region lifts; the table is stored with the answers."""
    d = os.path.join(HERE, 'lift')
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(seed)
    ref, tigs, df = synth.make_cigar_workload(seed, 1, 120_000, 4, 30_000, edit_rate=0.01, rev_frac=0.5, clip=(13, 7))
    df = df.reset_index(drop=True)
    fai = pd.Series({k: len(v) for k, v in tigs.items()})
    df.to_csv(os.path.join(d, 'align.bed'), sep='\t', index=False)
    fai.to_csv(os.path.join(d, 'tig.fai.tsv'), sep='\t', header=False)
    al = pavlib.align.AlignLift(df, fai)

    def plain(x):
        if x is None:
            return None
        return [x[0], int(x[1]), bool(x[2]), int(x[3]), int(x[4]), [int(i) for i in x[5]]]

    def call(f, *a, **k):
        try:
            return {'result': f(*a, **k)}
        except RuntimeError as ex:
            return {'error': str(ex)}
    out = []
    for _ in range(1500):
        row = df.iloc[int(rng.integers(0, df.shape[0]))]
        p = int(rng.integers(row['POS'] - 50, row['END'] + 50))
        r = call(al.lift_to_qry, row['#CHROM'], p)
        out.append({'f': 'to_qry', 'id': row['#CHROM'], 'pos': p, **({'result': plain(r['result'])} if 'result' in r else r)})
        q = int(rng.integers(max(row['QRY_POS'] - 30, 0), row['QRY_END'] + 30))
        gap = bool(rng.random() < 0.5)
        r = call(al.lift_to_sub, row['QRY_ID'], q, gap=gap)
        out.append({'f': 'to_sub', 'id': row['QRY_ID'], 'pos': q, 'gap': gap, **({'result': plain(r['result'])} if 'result' in r else r)})
    for _ in range(400):
        row = df.iloc[int(rng.integers(0, df.shape[0]))]
        a = int(rng.integers(row['POS'], row['END'] - 2000))
        b = a + int(rng.integers(10, 1900))
        rq = al.lift_region_to_qry(pavlib.seq.Region(row['#CHROM'], a, b))
        rec = {'f': 'region', 'chrom': row['#CHROM'], 'pos': a, 'end': b,
               'qry': None if rq is None else [rq.chrom, int(rq.pos), int(rq.end), bool(rq.is_rev)]}
        if rq is not None:
            for gap in (False, True):
                rs = al.lift_region_to_sub(rq, gap=gap)
                rec['sub_gap' if gap else 'sub'] = None if rs is None else [rs.chrom, int(rs.pos), int(rs.end), bool(rs.is_rev)]
        out.append(rec)
    with open(os.path.join(d, 'queries.json'), 'w') as fh:
        json.dump(out, fh)
    print('lift', len(out), 'queries,', sum(1 for r in out if r.get('result') is None and 'error' not in r and r['f'] != 'region'), 'None,',
          sum('error' in r for r in out), 'errors')


def make_cigar_cases_extra():
    """Cases added after the first fixture set (kept separate so the earlier fixtures are not rewritten)."""
    # records out of order (chromosomes interleaved, later positions first), a non-default frame index, two contigs over the
    # same reference span (ties on #CHROM, POS, END broken by ID; identical IDs versioned in sorted order)
    ref, tigs, dfa = synth.make_cigar_workload(78, 2, 40_000, 6, 12_000, edit_rate=0.015, rev_frac=0.5, clip=(3, 4))
    twin = dfa.iloc[[1, 4]].copy()
    twin['INDEX'] = [501, 502]
    dfx = pd.concat([dfa, twin], axis=0)
    rng = np.random.default_rng(78)
    dfx = dfx.iloc[rng.permutation(dfx.shape[0])]
    dfx.index = ['r%02d' % i for i in rng.permutation(dfx.shape[0])]
    cigar_case('unsorted_overlap_vid', ref, tigs, dfx, hap='h1', version_id=True)
    cigar_case('unsorted_overlap', ref, tigs, dfx, hap='h1', version_id=False)


def region_expand_cases(seed=5):
    """pavlib.seq.Region.expand of the reference: random regions, expansions, balances, with / without shifting, limits given as
    a Series of chromosome lengths, an int, or absent."""
    rng = np.random.default_rng(seed)
    fai = pd.Series({'c': 100000})
    out = []
    for _ in range(1200):
        p = int(rng.integers(0, 99000))
        e = p + int(rng.integers(1, 9000))
        bp = int(rng.integers(0, 60000))
        bal = float(rng.choice([0.25, 0.5, 0.75, 0.0, 1.0, 0.1]))
        shift = bool(rng.random() < 0.7)
        lim = ['fai', 'int', 'none'][int(rng.integers(0, 3))]
        min_pos = int(rng.choice([0, 0, 0, 500]))
        r = pavlib.seq.Region('c', p, min(e, 100000))
        r.expand(np.int32(bp) if rng.random() < 0.5 else bp, min_pos=min_pos, max_end={'fai': fai, 'int': 100000, 'none': None}[lim], shift=shift,
                 balance=bal)
        out.append({'pos': p, 'end': min(e, 100000), 'bp': bp, 'balance': bal, 'shift': shift, 'max_end': lim, 'min_pos': min_pos,
                    'result': [int(r.pos), int(r.end)]})
    with open(os.path.join(HERE, 'region_expand.json'), 'w') as fh:
        json.dump(out, fh)
    print('region_expand', len(out))


def count_cigar_cases(n_cases=2500, seed=31):
    """pavlib.align.count_cigar of the reference on random CIGAR strings: well-formed ones (clips in every legal and illegal
    arrangement, zero lengths, M with and without allow_m, N / P ops) and malformed text; result tuple or exception class + text."""
    import random
    rnd = random.Random(seed)
    out = []
    for _ in range(n_cases):
        kind = rnd.random()
        body = ''.join('{}{}'.format(rnd.choice([0, 1, 1, 2, 7, 30, 1234, 99999]) if rnd.random() < 0.2 else rnd.randint(1, 60),
                                     rnd.choice('==XXIDM' if rnd.random() < 0.15 else '==XXID')) for _ in range(rnd.randint(0, 12)))
        def clips():
            k = rnd.random()
            if k < 0.35:
                return []
            pool = ['S', 'H'] if k < 0.9 else ['S', 'H', 'S', 'H']
            return rnd.sample(pool, rnd.randint(1, len(pool)))
        lead = ''.join('{}{}'.format(rnd.choice([0, 3, 17, 250]), c) for c in clips())
        tail = ''.join('{}{}'.format(rnd.choice([0, 3, 17, 250]), c) for c in clips())
        cigar = lead + body + tail
        if kind < 0.08 and body:      # a clip or an exotic op in the middle
            ops = [m for m in __import__('re').findall(r'\d+.', cigar)]
            ops.insert(rnd.randint(0, len(ops)), '{}{}'.format(rnd.randint(1, 9), rnd.choice('SHNP')))
            cigar = ''.join(ops)
        elif kind < 0.12:             # malformed text
            cigar = rnd.choice([cigar + '12', '=' + cigar, cigar.replace('=', 'Q', 1) if '=' in cigar else cigar + '5Q', cigar + '3==', ''])
        allow_m = rnd.random() < 0.3
        rec = {'cigar': cigar, 'allow_m': allow_m}
        row = pd.Series({'CIGAR': cigar, 'QRY_ID': 'tigA', '#CHROM': 'chrA', 'POS': 5})
        try:
            rec['result'] = [int(x) for x in pavlib.align.count_cigar(row, allow_m=allow_m)]
        except Exception as ex:   # noqa: BLE001
            rec['error'] = [type(ex).__name__, str(ex)]
        out.append(rec)
    with open(os.path.join(HERE, 'count_cigar.json'), 'w') as fh:
        json.dump(out, fh)
    n_err = sum('error' in r for r in out)
    print('count_cigar', len(out), 'cases,', n_err, 'errors,', len({r['error'][1].split(' at ')[0][:40] for r in out if 'error' in r}), 'kinds of message')


if __name__ == '__main__':
    what = set(sys.argv[1:]) or {'cigar', 'homology', 'kmer', 'density', 'inv', 'align', 'density_cli', 'count_cigar', 'lift', 'region_expand', 'cigar_extra'}
    if 'density_cli' in what:
        make_density_cli()
    if 'count_cigar' in what:
        count_cigar_cases()
    if 'cigar_extra' in what:
        make_cigar_cases_extra()
    if 'lift' in what:
        lift_cases()
    if 'region_expand' in what:
        region_expand_cases()
    if 'align' in what:
        make_align_case()
    if 'cigar' in what:
        make_cigar_cases()
    if 'homology' in what:
        homology_cases()
    if 'kmer' in what:
        kmer_cases()
    if 'density' in what:
        make_density_cases()
    if 'inv' in what:
        make_inv_case()
        make_inv_case('kat3_rev', rev=True, seed=4)
        make_inv_case('kat3_flank', rev=False, seed=5, flank_rep=1500)
