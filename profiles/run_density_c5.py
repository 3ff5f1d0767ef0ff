#!/usr/bin/env python3
"""Profiling driver: one C5-shaped density batch (N windows x 50 kbp, k=31, srs=20) run R times on cuda:0.
Used under ncu on the B200 box:  ncu --set full ... python profiles/run_density_c5.py 48 2"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from pav_b200 import _capi, device, synth  # noqa: E402
from pav_b200.pavlib import density  # noqa: E402

n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 48
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = device.get_context()
ref, tig, meta = synth.make_inv_workload(seed=1005, n_win=n_win, win_len=50_000)
rs = device.SeqStore(ctx, list(ref), [ref[n] for n in ref], keep_host=False)
ts = device.SeqStore(ctx, list(tig), [tig[n] for n in tig], keep_host=False)
win = np.zeros(n_win, dtype=_capi.DENSITY_WINDOW)
for i in range(n_win):
    win[i] = (i, i, 0, 50_000, 0, 50_000, 0, 20)
batch = density.DensityBatch(ctx, win, density.default_params())
for _ in range(reps):
    st = batch.run(rs, ts)
print(st.as_dict())
