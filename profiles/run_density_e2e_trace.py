#!/usr/bin/env python3
"""Phase timing of the public density call (pavlib.density.density_windows) on C5-shaped windows:
    PAVGPU_TRACE=1 python profiles/run_density_e2e_trace.py [n_windows] [n_calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pav_b200 import synth  # noqa: E402
from pav_b200.pavlib import density  # noqa: E402

n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n_calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ref, tig, meta = synth.make_inv_workload(seed=1005, n_win=n_win, win_len=50_000)
wins = [(ref[a], tig[b], False, 20) for a, b, _, _ in meta]
for i in range(2 * n_calls):
    lazy = i >= n_calls          # second half: run lengths only (columns stay in HBM until asked for)
    print(f'--- call {i} lazy={lazy}', file=sys.stderr, flush=True)
    out = None
    t0 = time.perf_counter()
    out = density.density_windows(wins, lazy=lazy)
    dt = time.perf_counter() - t0
    print(f'call {i} lazy={lazy}: {dt * 1e3:.1f} ms = {n_win * 50_000 / dt / 1e9:.3f} Gbases/s; seconds={ {k: round(v, 4) for k, v in density.last_stats["seconds"].items()} } '
          f'kernels={density.last_stats["ms_kernels"]:.2f} ms d2h={density.last_stats["ms_d2h"]:.2f} ms', flush=True)

if os.environ.get('PROFILE'):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    out = density.density_windows(wins, lazy=True)
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
