#!/usr/bin/env python3
"""Golden fixtures for the flagging rules between Path A and Path B, produced by executing the
UNMODIFIED ``run:`` blocks of the reference's Snakemake rules in the build container:

    rules/call.snakefile      rule call_cigar                     (FILTER = PASS / TRIM)
    rules/call_inv.snakefile  rule call_inv_cluster               (SNV / indel clusters)
                              rule call_inv_flag_insdel_cluster   (matched INS / DEL)
                              rule call_inv_merge_flagged_loci    (merge, TRY_INV, BATCH)

The rule bodies are cut out of the snakefiles *as text at generation time* (nothing is copied into this repository),
wrapped into a function (Snakemake does the same, which is why they may ``return``) and run with stand-ins for
``input / output / params / wildcards / get_config``. Inputs and outputs are committed under tests/golden/flag/.

Run:  python tests/golden/make_golden_flag.py      (container only; needs /root/reference)
"""
import collections
import os
import re
import sys
import textwrap
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refenv  # noqa: E402
from pav_b200 import synth  # noqa: E402

refenv.activate()
import intervaltree  # noqa: E402  (stub)
import pavlib  # noqa: E402  (the reference)

OUT = os.path.join(HERE, 'flag')


def rule_body(snakefile, rule):
    """Source of the ``run:`` block of ``rule`` as a function ``run(input, output, params, wildcards)``."""
    lines = open(os.path.join(refenv.REF_ROOT, 'rules', snakefile)).read().split('\n')
    i = next(k for k, ln in enumerate(lines) if re.match(r'^rule\s+%s\s*:' % re.escape(rule), ln))
    j = next(k for k in range(i, len(lines)) if lines[k].rstrip() == '    run:')
    body = []
    for ln in lines[j + 1:]:
        if ln.strip() and len(ln) - len(ln.lstrip()) < 8:
            break
        body.append(ln)
    return 'def run(input, output, params, wildcards):\n' + textwrap.indent(textwrap.dedent('\n'.join(body)), '    ') + '\n'


def function_source(snakefile, name):
    lines = open(os.path.join(refenv.REF_ROOT, 'rules', snakefile)).read().split('\n')
    i = next(k for k, ln in enumerate(lines) if ln.startswith('def %s(' % name))
    body = [lines[i]]
    for ln in lines[i + 1:]:
        if ln.strip() and not ln.startswith(' '):
            break
        body.append(ln)
    return '\n'.join(body) + '\n'


def run_rule(snakefile, rule, input, output, params=None, wildcards=None, config=None, extra=None):
    g = {'pd': pd, 'np': np, 'collections': collections, 'intervaltree': intervaltree, 'pavlib': pavlib, 'os': os,
         'get_config': lambda wc, key=None, default=None, default_none=False: dict(config or {}) if key is None else (config or {}).get(key, default),
         'BATCH_COUNT_DEFAULT': 60}
    g.update(extra or {})
    exec(compile(rule_body(snakefile, rule), f'{snakefile}:{rule}', 'exec'), g)
    ns = types.SimpleNamespace

    class Wildcards(dict):   # attribute access and ``**wildcards`` both work in rule bodies
        __getattr__ = dict.__getitem__
    g['run'](ns(**input), ns(**output), ns(**(params or {})), Wildcards(wildcards or {}))


def variant_tables(seed):
    """Synthetic merged CIGAR call tables (the columns the flagging rules read): sparse background, planted SNV / indel
    clusters, matched INS/DEL pairs (SV- and indel-sized), some FILTER=TRIM rows, three chromosomes."""
    rng = np.random.default_rng(seed)
    snv, indel = [], []
    for chrom, clen in (('chr1', 3_000_000), ('chr2', 2_000_000), ('chrX', 800_000)):
        # background
        for p in np.sort(rng.choice(clen - 1000, size=clen // 1500, replace=False)).tolist():
            snv.append((chrom, p, p + 1, 'SNV', 1))
        for p in np.sort(rng.choice(clen - 1000, size=clen // 8000, replace=False)).tolist():
            ln = int(min(rng.geometric(0.2), 45))
            t = 'INS' if rng.random() < 0.5 else 'DEL'
            indel.append((chrom, p, p + (1 if t == 'INS' else ln), t, ln))
        # SNV clusters: 12..60 SNVs spaced 5..190 bp (some too short / too few to qualify)
        for c in range(clen // 150_000):
            p = int(rng.integers(10_000, clen - 50_000))
            for _ in range(int(rng.integers(12, 60))):
                p += int(rng.integers(5, 190))
                snv.append((chrom, p, p + 1, 'SNV', 1))
        # indel clusters: 6..30 indels spaced 10..190 bp
        for c in range(clen // 200_000):
            p = int(rng.integers(10_000, clen - 50_000))
            for _ in range(int(rng.integers(6, 30))):
                p += int(rng.integers(10, 190))
                ln = int(rng.integers(1, 49))
                t = 'INS' if rng.random() < 0.5 else 'DEL'
                indel.append((chrom, p, p + (1 if t == 'INS' else ln), t, ln))
        # matched INS/DEL pairs: SV-sized and indel-sized, some just outside the flank, some chained within the merge flank
        for c in range(clen // 100_000):
            p = int(rng.integers(10_000, clen - 50_000))
            ln = int(rng.integers(50, 6000)) if rng.random() < 0.6 else int(rng.integers(4, 49))
            off = int(rng.integers(0, int(ln * 2.4)))
            indel.append((chrom, p, p + ln, 'DEL', ln))
            indel.append((chrom, max(p + ln // 2 + (off if rng.random() < 0.5 else -off), 1), 0, 'INS', int(ln * rng.uniform(0.8, 1.2))))
            if rng.random() < 0.3:
                q = p + ln + int(rng.integers(100, 3000))
                indel.append((chrom, q, q + ln, 'DEL', ln))
                indel.append((chrom, q + 5, 0, 'INS', ln))
    df_snv = pd.DataFrame(snv, columns=['#CHROM', 'POS', 'END', 'SVTYPE', 'SVLEN'])
    df_indel = pd.DataFrame(indel, columns=['#CHROM', 'POS', 'END', 'SVTYPE', 'SVLEN'])
    df_indel.loc[df_indel['SVTYPE'] == 'INS', 'END'] = df_indel.loc[df_indel['SVTYPE'] == 'INS', 'POS'] + 1
    for df, tag in ((df_snv, 'SNV'), (df_indel, None)):
        df['ID'] = [f'{c}-{p + 1}-{t}-{ln}' for c, p, t, ln in zip(df['#CHROM'], df['POS'], df['SVTYPE'], df['SVLEN'])]
        df['FILTER'] = np.where(rng.random(len(df)) < 0.03, 'TRIM', 'PASS')
    cols = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'FILTER']
    return (df_snv.sort_values(['#CHROM', 'POS'])[cols].reset_index(drop=True),
            df_indel.sort_values(['#CHROM', 'POS', 'END', 'ID'])[cols].reset_index(drop=True))


def flag_case(name, seed, config=None):
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    df_snv, df_indel = variant_tables(seed)
    if name == 'empty':
        df_snv, df_indel = df_snv.iloc[:0], df_indel.iloc[:0]
    if name == 'no_snv':      # one of the four flag tables empty
        df_snv = df_snv.iloc[:0]
    if name == 'no_sv':       # no SV-sized pairs and no indel clusters
        df_indel = df_indel.loc[(df_indel['SVLEN'] < 50) & (df_indel['SVLEN'] >= 30)]
    p_snv, p_indel = os.path.join(d, 'snv.bed.gz'), os.path.join(d, 'insdel.bed.gz')
    df_snv.to_csv(p_snv, sep='\t', index=False, compression='gzip')
    df_indel.to_csv(p_indel, sep='\t', index=False, compression='gzip')
    out = {}
    for vartype, files in (('indel', [p_indel]), ('snv', [p_snv])):
        out[f'cluster_{vartype}'] = os.path.join(d, f'cluster_{vartype}.bed.gz')
        run_rule('call_inv.snakefile', 'call_inv_cluster', {'bed': files}, {'bed': out[f'cluster_{vartype}']},
                 params={'cluster_win': 200, 'cluster_win_min': 500, 'cluster_min_snv': 20, 'cluster_min_indel': 10}, wildcards={'vartype': vartype})
    for vartype in ('sv', 'indel'):
        out[f'insdel_{vartype}'] = os.path.join(d, f'insdel_{vartype}.bed.gz')
        run_rule('call_inv.snakefile', 'call_inv_flag_insdel_cluster', {'bed': p_indel}, {'bed': out[f'insdel_{vartype}']},
                 params={'flank_cluster': 2, 'flank_merge': 2000, 'cluster_min_svlen': 4}, wildcards={'vartype': vartype})
    accept = {}
    exec(function_source('call_inv.snakefile', '_call_inv_accept_flagged_region'), accept)
    for filt in ('svindel', 'sv', 'single_cluster'):
        run_rule('call_inv.snakefile', 'call_inv_merge_flagged_loci',
                 {'bed_insdel_sv': out['insdel_sv'], 'bed_insdel_indel': out['insdel_indel'], 'bed_cluster_indel': out['cluster_indel'],
                  'bed_cluster_snv': out['cluster_snv']},
                 {'bed': os.path.join(d, f'flagged_regions_{filt}.bed.gz')}, wildcards={'asm_name': 'x', 'hap': 'h1'},
                 config=dict(config or {}, inv_sig_filter=filt), extra={'_call_inv_accept_flagged_region': accept['_call_inv_accept_flagged_region']})
    print(name, {k: pd.read_csv(v, sep='\t').shape[0] for k, v in out.items()},
          pd.read_csv(os.path.join(d, 'flagged_regions_svindel.bed.gz'), sep='\t').shape[0])


def filter_case(name, seed):
    """rule call_cigar end to end on a small alignment table with a trimmed table that cuts some records."""
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    ref, tigs, df = synth.make_cigar_workload(seed, 2, 120_000, 12, 18_000, edit_rate=0.01, rev_frac=0.5)
    df = df.reset_index(drop=True)
    df['CALL_BATCH'] = df['INDEX'] % 3
    ref_fa, tig_fa, bed = synth.write_cigar_workload(d, ref, tigs, df)
    rng = np.random.default_rng(seed)
    trim = df[['#CHROM', 'POS', 'END', 'INDEX']].copy()
    trim['POS'] += rng.integers(0, 3000, size=len(trim))
    trim['END'] -= rng.integers(0, 3000, size=len(trim))
    trim = trim.iloc[[i for i in range(len(trim)) if i % 5 != 4]]     # some records vanish after trimming
    bed_trim = os.path.join(d, 'wl_align_trim.bed')
    trim.to_csv(bed_trim, sep='\t', index=False)
    for batch in (0, 1):
        run_rule('call.snakefile', 'call_cigar', {'bed': bed, 'bed_trim': bed_trim, 'tig_fa_name': tig_fa},
                 {'bed_insdel': os.path.join(d, f'insdel_{batch}.bed.gz'), 'bed_snv': os.path.join(d, f'snv_{batch}.bed.gz')},
                 wildcards={'batch': str(batch), 'hap': 'h1'}, extra={'REF_FA': ref_fa})
        print(name, batch, pd.read_csv(os.path.join(d, f'snv_{batch}.bed.gz'), sep='\t')['FILTER'].value_counts().to_dict())


def inv_batch_case(name, seed=17):
    """rule call_inv_batch end to end (rules/call_inv.snakefile:115-311): a 120 kbp locus with two inversions, a flagged table with
    a region per inversion, a second region that finds the first inversion again (dropped as a duplicate), a region without an
    inversion, and a row of another batch. The reference's scan_for_inv spawns scripts/density.py per expansion, as in production."""
    import gc
    import random

    import kanapy.util.kmer
    import svpoplib
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    random.seed(seed)
    n = 120_000
    s0 = ''.join(random.choice('ACGT') for _ in range(n))
    comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
    rc = lambda x: ''.join(comp[c] for c in reversed(x))  # noqa: E731
    t = s0[:26000] + rc(s0[26000:34000]) + s0[34000:86000] + rc(s0[86000:91000]) + s0[91000:]
    ref_fa, tig_fa = os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa')
    synth.write_fasta(ref_fa, {'chr1': np.frombuffer(s0.encode(), dtype=np.uint8)}, line_width=60)
    synth.write_fasta(tig_fa, {'tig1': np.frombuffer(t.encode(), dtype=np.uint8)}, line_width=60)
    pd.DataFrame([('chr1', 0, n, 0, 'tig1', 0, n, n, False, f'{n}=')],
                 columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END', 'QRY_LEN', 'REV', 'CIGAR']
                 ).to_csv(os.path.join(d, 'align.bed'), sep='\t', index=False)
    flag_rows = [('chr1', 28000, 32000, 'chr1-28000-RGN-4000', 'RGN', 4000, 'MATCH_SV', 0, 0, True, 0),
                 ('chr1', 29000, 31000, 'chr1-29000-RGN-2000', 'RGN', 2000, 'CLUSTER_SNV,MATCH_INDEL', 0, 25, True, 0),
                 ('chr1', 60000, 61000, 'chr1-60000-RGN-1000', 'RGN', 1000, 'MATCH_INDEL', 0, 0, True, 0),
                 ('chr1', 87000, 90000, 'chr1-87000-RGN-3000', 'RGN', 3000, 'MATCH_SV', 0, 0, True, 0),
                 ('chr1', 100000, 101000, 'chr1-100000-RGN-1000', 'RGN', 1000, 'MATCH_SV', 0, 0, True, 1)]
    pd.DataFrame(flag_rows, columns=['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'TYPE', 'COUNT_INDEL', 'COUNT_SNV', 'TRY_INV', 'BATCH']
                 ).to_csv(os.path.join(d, 'flagged.bed.gz'), sep='\t', index=False, compression='gzip')
    os.environ['PYTHONPATH'] = os.pathsep.join(refenv.pythonpath_entries())   # for the scripts/density.py children
    cwd = os.getcwd()
    work = os.path.join(d, '_work')
    os.makedirs(work, exist_ok=True)
    os.chdir(work)
    try:
        for batch in (0, 1, 5):
            run_rule('call_inv.snakefile', 'call_inv_batch',
                     {'bed_flag': os.path.join(d, 'flagged.bed.gz'), 'bed_aln': os.path.join(d, 'align.bed'), 'tig_fa': tig_fa, 'fai': tig_fa + '.fai'},
                     {'bed': os.path.join(d, f'inv_call_{batch}.bed.gz')}, wildcards={'asm_name': 'asm', 'hap': 'h1', 'batch': str(batch)},
                     extra={'REF_FA': ref_fa, 'kanapy': kanapy, 'svpoplib': svpoplib, 'gc': gc, 'threads': 1,
                            'log': types.SimpleNamespace(log=os.path.join(d, f'inv_call_{batch}.log'))})
            print(name, batch, pd.read_csv(os.path.join(d, f'inv_call_{batch}.bed.gz'), sep='\t').iloc[:, :6].to_string())
    finally:
        os.chdir(cwd)
    import shutil
    shutil.rmtree(work)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'inv':
        inv_batch_case('inv_batch')
        sys.exit(0)
    flag_case('a', 501)
    flag_case('b', 502, config={'inv_sig_merge_flank': 5000, 'inv_sig_batch_count': 7})
    flag_case('empty', 503)
    flag_case('no_snv', 504)
    flag_case('no_sv', 505)
    filter_case('filter', 511)
