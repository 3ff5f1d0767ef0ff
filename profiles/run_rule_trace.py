#!/usr/bin/env python3
"""`rule call_cigar` end to end on the C2 workload (alignment table + FASTA files in, two bed.gz out), three ways:
  frames   flag.call_cigar (GPU walk -> DataFrames -> FILTER) + DataFrame.to_csv(compression='gzip'), as the rule does
  direct   flag.call_cigar_to_files (GPU walk -> TSV text in C -> parallel gzip members)
and checks that both write the same tables.   python profiles/run_rule_trace.py [n_calls]"""
import gzip
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pav_b200 import synth  # noqa: E402
from pav_b200.pavlib import flag  # noqa: E402

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ref, trs = synth.make_reference(1002, 4, 50_000_000)
tigs, df = synth.make_contigs(ref, trs, 1002, 1000, 200_000)
df = df.reset_index(drop=True)
df['CALL_BATCH'] = 0
tmp = tempfile.mkdtemp(prefix='pav_rule_')
ref_fa, tig_fa, _ = synth.write_cigar_workload(tmp, ref, tigs, df)
del ref, tigs
rng = np.random.default_rng(1)
trim = df[['POS', 'END', 'INDEX']].copy()
trim['POS'] += rng.integers(0, 3000, size=len(trim))
trim['END'] -= rng.integers(0, 3000, size=len(trim))
trim = trim.set_index('INDEX').astype(int)
a_s, a_i, b_s, b_i = (os.path.join(tmp, x) for x in ('a_snv.bed.gz', 'a_insdel.bed.gz', 'b_snv.bed.gz', 'b_insdel.bed.gz'))
for k in range(n_calls):
    t0 = time.perf_counter()
    df_snv, df_insdel = flag.call_cigar(df, 0, ref_fa, tig_fa, 'h1', trim)
    t1 = time.perf_counter()
    df_insdel.to_csv(a_i, sep='\t', index=False, compression='gzip')
    df_snv.to_csv(a_s, sep='\t', index=False, compression='gzip')
    t2 = time.perf_counter()
    rows = len(df_snv) + len(df_insdel)
    del df_snv, df_insdel
    t3 = time.perf_counter()
    n = flag.call_cigar_to_files(df, 0, ref_fa, tig_fa, 'h1', trim, b_i, b_s)
    t4 = time.perf_counter()
    print(f'call {k}: {rows} rows; frames: walk+frames+FILTER {t1 - t0:.2f}s + to_csv(gzip) {t2 - t1:.2f}s = {t2 - t0:.2f}s '
          f'({rows / (t2 - t0):.3e} rows/s); direct: {t4 - t3:.2f}s ({sum(n) / (t4 - t3):.3e} rows/s)', flush=True)
same = gzip.open(a_s, 'rb').read() == gzip.open(b_s, 'rb').read() and gzip.open(a_i, 'rb').read() == gzip.open(b_i, 'rb').read()
print('tables identical:', same, '; sizes', os.path.getsize(a_s), os.path.getsize(b_s), os.path.getsize(a_i), os.path.getsize(b_i))
assert same
