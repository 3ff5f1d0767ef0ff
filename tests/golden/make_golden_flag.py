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
         'get_config': lambda wc, key, default=None, default_none=False: (config or {}).get(key, default), 'BATCH_COUNT_DEFAULT': 60}
    g.update(extra or {})
    exec(compile(rule_body(snakefile, rule), f'{snakefile}:{rule}', 'exec'), g)
    ns = types.SimpleNamespace
    g['run'](ns(**input), ns(**output), ns(**(params or {})), ns(**(wildcards or {})))


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


if __name__ == '__main__':
    flag_case('a', 501)
    flag_case('b', 502, config={'inv_sig_merge_flank': 5000, 'inv_sig_batch_count': 7})
    flag_case('empty', 503)
    flag_case('no_snv', 504)
    flag_case('no_sv', 505)
    filter_case('filter', 511)
