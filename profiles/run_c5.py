#!/usr/bin/env python3
"""BASELINE configs[4] at full size on one B200: 10,000 flagged 50 kbp windows (central inversion of 2-20 kbp, 30 % with
1-3 kbp inverted-repeat flanks, 0.5 % divergence, 10 % negative controls), k=31, srs=20, delta=0.005, density scan in
batches, with size-independent properties per window and an oracle comparison on a sample.

    python profiles/run_c5.py [--windows 10000] [--chunk 1024] [--out profiles/rNN_c5.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import properties as checks  # noqa: E402
from pav_b200 import _capi, device, synth  # noqa: E402
from pav_b200.pavlib import density  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def state_runs(state, index):
    brk = np.flatnonzero(state[1:] != state[:-1]) + 1
    starts = np.concatenate(([0], brk))
    ends = np.concatenate((brk, [len(state)]))
    return [(int(state[a]), int(b - a), int(index[a]), int(index[b - 1])) for a, b in zip(starts, ends)]


def run(n_windows=10_000, chunk=1024, win_len=50_000, oracle_windows=2, seed=1005):
    ctx = device.get_context()
    params = density.default_params()
    tot = {'ms': 0.0, 'ms_kmer': 0.0, 'ms_kde': 0.0, 'rows': 0, 'kde_pairs': 0}
    n_inv_found = n_inv = n_neg = n_neg_clean = n_flank = n_flank_seen = n_fail = 0
    oracle_done = 0
    t_gen = t_up = 0.0
    for c, a in enumerate(range(0, n_windows, chunk)):
        n = min(chunk, n_windows - a)
        t0 = time.perf_counter()
        ref, tig, meta = synth.make_inv_workload(seed=seed + 7919 * c, n_win=n, win_len=win_len)
        t_gen += time.perf_counter() - t0
        t0 = time.perf_counter()
        rs = device.SeqStore(ctx, list(ref), [ref[k] for k in ref], keep_host=False)
        ts = device.SeqStore(ctx, list(tig), [tig[k] for k in tig], keep_host=False)
        t_up += time.perf_counter() - t0
        win = np.zeros(n, dtype=_capi.DENSITY_WINDOW)
        for i in range(n):
            win[i] = (i, i, 0, win_len, 0, win_len, 0, 20)
        batch = density.DensityBatch(ctx, win, params)
        ctx.l2_flush()
        st = batch.run(rs, ts)
        res, cols = batch.fetch()
        batch.close()
        rs.close()
        ts.close()
        for k in ('ms_kmer', 'ms_kde'):
            tot[k] += getattr(st, k)
        tot['ms'] += st.ms_kernels
        tot['rows'] += int(st.rows)
        tot['kde_pairs'] += int(st.kde_pairs)
        for i, d in enumerate(density._split(res, cols)):
            r_name, t_name, (ia, ib), neg = meta[i]
            if not checks.check_density_window(d, win_len, 31):
                n_fail += 1
                continue
            runs = state_runs(d['STATE'], d['INDEX']) if d['smoothed'] else []
            rev_cov = sum(min(e, ib) - max(s, ia) for stt, _, s, e in runs if stt == 2 and min(e, ib) > max(s, ia))
            if neg:
                n_neg += 1
                n_neg_clean += all(stt != 2 for stt, _, _, _ in runs)   # inverted-repeat flanks give FWDREV runs, never REV
            else:
                n_inv += 1
                n_inv_found += rev_cov >= 0.8 * (ib - ia - 62)
            if oracle_done < oracle_windows and d['smoothed']:
                from oracle import pyoracle
                rc, o = pyoracle.density_arrays(ref[r_name].tobytes(), tig[t_name].tobytes())
                assert rc == 0 and all((d[k].astype(np.int64) == o[k].astype(np.int64)).all() for k in ('KMER', 'INDEX', 'STATE_MER', 'STATE')), \
                    'window differs from the oracle'
                for k in ('KERN_FWD', 'KERN_FWDREV', 'KERN_REV'):
                    np.testing.assert_allclose(d[k], o[k], rtol=1e-9, atol=1e-300)
                oracle_done += 1
        log(f'chunk {c}: {n} windows, {st.ms_kernels:.2f} ms')
    bases = n_windows * win_len
    out = {'config': f'C5: {n_windows} windows x {win_len} bp, k=31, srs=20, delta=0.005, batches of {chunk}', 'windows': n_windows, 'bases': bases,
           'ms': tot['ms'], 'ms_kmer': tot['ms_kmer'], 'ms_kde': tot['ms_kde'], 'gbases_per_s': bases / (tot['ms'] * 1e-3) / 1e9,
           'rows': tot['rows'], 'kde_pairs': tot['kde_pairs'], 'seconds_generate': t_gen, 'seconds_upload_pack': t_up,
           'properties': {'windows_with_inversion': n_inv, 'inversion_recovered_as_REV_run': int(n_inv_found), 'negative_controls': n_neg,
                          'negative_controls_without_REV_run': int(n_neg_clean), 'soft_failures': n_fail},
           'oracle': {'windows': oracle_done, 'identical_discrete_columns': True, 'kern_rtol': 1e-9}}
    assert n_inv_found >= 0.99 * n_inv, out
    assert n_neg_clean >= 0.99 * n_neg, out
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--windows', type=int, default=10_000)
    ap.add_argument('--chunk', type=int, default=1024)
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    from oracle import pyoracle
    from pav_b200 import build
    build.build()
    pyoracle.build()
    out = run(args.windows, args.chunk)
    print(json.dumps(out))
    if args.out:
        with open(args.out, 'w') as fh:
            json.dump(out, fh, indent=1)


if __name__ == '__main__':
    main()
