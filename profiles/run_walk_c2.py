#!/usr/bin/env python3
"""Profiling driver: the C2 walk step (1,000 contigs x 200 kbp vs 200 Mbp, the input of bench.py's `c2` leg) run R times on cuda:0,
kernel by kernel (PAVGPU_NO_GRAPH=1) so that ncu sees every launch.   ncu --set full ... python profiles/run_walk_c2.py 4"""
import os
import sys

os.environ.setdefault('PAVGPU_NO_GRAPH', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from pav_b200 import device, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_contigs, contig_len = 1000, 200_000
ctx = device.get_context()
ref, trs = synth.make_reference(1002, 4, n_contigs * contig_len // 4)
tigs, df = synth.make_contigs(ref, trs, 1002, n_contigs, contig_len)
names_r, names_t = list(ref), list(tigs)
ref_store = device.SeqStore(ctx, names_r, [ref[n] for n in names_r])
tig_store = device.SeqStore(ctx, names_t, [tigs[n] for n in names_t])
rid = np.array([names_r.index(c) for c in df['#CHROM']], np.int32)
tidx = {n: i for i, n in enumerate(names_t)}
qid = np.array([tidx[c] for c in df['QRY_ID']], np.int32)
ops, op_off, perr = device.parse_cigars(df['CIGAR'].tolist())
batch = device.CigarBatch(ctx, rid, qid, df['POS'].to_numpy(np.int32), df['REV'].to_numpy(np.uint8), ops, op_off)
for _ in range(reps):
    ctx.l2_flush()
    st = batch.run(ref_store, tig_store)
print(st.as_dict())
