#!/usr/bin/env python3
"""Phase-by-phase timing of the public-API call on the C2 workload (host wall clock + PAVGPU_TRACE=1 from the C layer).

    PAVGPU_TRACE=1 python profiles/run_e2e_trace.py [n_calls] 2> trace.log
"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pav_b200 import fasta, synth  # noqa: E402
from pav_b200.pavlib import cigarcall  # noqa: E402

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ref, trs = synth.make_reference(1002, 4, 50_000_000)
tigs, df = synth.make_contigs(ref, trs, 1002, 1000, 200_000)
tmp = tempfile.mkdtemp(prefix='pav_e2e_')
ref_fa, tig_fa, _ = synth.write_cigar_workload(tmp, ref, tigs, df)
del ref, tigs
for i in range(n_calls):
    fasta._CACHE.clear()
    print(f'--- call {i}', file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    a, b = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=False)
    dt = time.perf_counter() - t0
    print(f'call {i}: {dt:.3f}s rows={len(a) + len(b)} phases={ {k: round(v, 4) for k, v in cigarcall.last_phase_seconds.items()} } '
          f'walk={ {k: round(v, 4) for k, v in cigarcall.last_walk_seconds.items()} }', flush=True)
    del a, b
