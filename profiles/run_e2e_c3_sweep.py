#!/usr/bin/env python3
"""End-to-end make_insdel_snv_calls on C3 (one haplotype) under different host settings: FASTA reader threads, pinned staging cap.
    python profiles/run_e2e_c3_sweep.py"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pav_b200 import fasta as fasta_mod  # noqa: E402
from pav_b200.pavlib import cigarcall  # noqa: E402

os.environ.setdefault('PAVGPU_TUNE_ALLOC', '1')
args = bench.parse_args()
tmp = tempfile.mkdtemp(prefix='c3sweep_')
bench.make_c3_files(args, tmp)
df = bench.read_align(os.path.join(tmp, 'h1_align.bed'))
ref_fa, tig_fa = os.path.join(tmp, 'ref.fa'), os.path.join(tmp, 'h1_tig.fa')
for readers, pin_mb in ((3, 1024), (6, 1024), (12, 1024), (6, 8192), (12, 8192)):
    cigarcall._READERS = readers
    cigarcall._PINNED_STAGING_MAX = pin_mb << 20
    for i in range(3):
        fasta_mod._CACHE.clear()
        out = None
        t0 = time.perf_counter()
        out = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=False)
        dt = time.perf_counter() - t0
        print(f'readers={readers} pinned_max={pin_mb} MB call {i}: {dt:.3f}s rows={len(out[0]) + len(out[1])} phases={ {k: (round(v, 3) if isinstance(v, float) else v) for k, v in cigarcall.last_phase_seconds.items()} } '
              f'walk={ {k: round(v, 3) for k, v in cigarcall.last_walk_seconds.items()} }', flush=True)
