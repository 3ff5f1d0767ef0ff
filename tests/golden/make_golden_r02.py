#!/usr/bin/env python3
"""Round-2 additions to the golden fixtures, generated like make_golden.py by running the UNMODIFIED reference
(scripts/density.py spawned the way pavlib/inv.py:249-266 spawns it) in the build container:

  density/few_informative_gaps   fewer than 2000 informative k-mers AND rows missing from the k-mer stream (an N run, a segment
                                 that matches nothing): the raw frame keeps the row labels of the stream (scripts/density.py:161-194);
                                 the labels are stored in meta['index'].
  density/tie_argmax_*           near ties of the float contract: FWD, a short REV block, FWD, with lengths searched
                                 (numpy restatement of the KDE, below) so that at a lattice point on the REV block's slope the FWD and
                                 REV densities differ by a relative margin of ~1e-6 (scripts/density.py:250-254, :335-338)
  density/tie_delta_*            the same search for a sampled gap whose max |delta KERN| lies within ~1e-8 of the 0.005 threshold
                                 (scripts/density.py:275-278) with constant STATE_MER and equal argmax at both ends, i.e. the gap is
                                 interpolated or fully evaluated depending on the 9th digit
  density/spike_near_one         a 22-k-mer REV cluster in ~2,100 informative k-mers: bandwidth ~1.3 lattice steps, KERN_REV peaks at
                                 1 + O(1e-14), the `> 1.0 -> reciprocal` branch (scripts/density.py:330-332)

    python tests/golden/make_golden_r02.py          (container only; needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (activates the reference environment)
from pav_b200 import synth  # noqa: E402

K = 31
S = 20


def kde(points, n_total, at):
    """scripts/density.py:69-115 + scipy.stats.gaussian_kde restated: sum_i exp(-((x_i - j) / h)^2 / 2) / (h sqrt(2 pi)),
    h = std(points, ddof=1) * n_total^(-1/5)."""
    h = np.std(points, ddof=1) * n_total ** (-0.2)
    d = (points[None, :] - at[:, None]) / h
    return np.exp(-0.5 * d * d).sum(axis=1) / (h * np.sqrt(2 * np.pi))


def two_runs(m0, m2):
    n = m0 + m2
    return np.arange(m0, dtype=np.float64), np.arange(m0, n, dtype=np.float64), n


def search_argmax(rng, trials=20000):
    """FWD a, REV b, FWD c with a small REV block: the REV density rises from ~0 to its peak across the block's edges while the FWD
    density stays near 0.9, so the two cross on a slope whose position moves continuously with (a, b, c) -- unlike the boundary of
    two long runs, where the crossing is pinned half way between two lattice points (margin always ~1 / (h sqrt(2 pi)))."""
    best = []
    for _ in range(trials):
        a, b, c = int(rng.integers(900, 2200)), int(rng.integers(25, 260)), int(rng.integers(900, 2200))
        n = a + b + c
        x0 = np.concatenate((np.arange(a), np.arange(a + b, n))).astype(np.float64)
        x2 = np.arange(a, a + b, dtype=np.float64)
        at = np.arange(max(a - 60, 0), min(a + b + 60, n), dtype=np.float64)
        k0, k2 = kde(x0, n, at), kde(x2, n, at)
        sgn = np.sign(k0 - k2)
        cross = np.flatnonzero(sgn[1:] != sgn[:-1])
        for cpos in cross:
            for j in (cpos, cpos + 1):
                best.append((abs(k0[j] - k2[j]) / max(k0[j], k2[j]), a, b, c, int(at[j])))
    best.sort()
    return best[:3]


def search_delta(rng, trials=6000, delta=0.005):
    best = []
    for _ in range(trials):
        m0, m2 = int(rng.integers(1100, 2600)), int(rng.integers(1100, 2600))
        x0, x2, n = two_runs(m0, m2)
        samp = np.arange(0, n, S, dtype=np.float64)
        k0, k2 = kde(x0, n, samp), kde(x2, n, samp)
        dm = np.maximum(np.abs(np.diff(k0)), np.abs(np.diff(k2)))
        a = samp[:-1].astype(int)
        b = a + S
        same_mer = (b < m0) | (a >= m0)                                   # STATE_MER constant on [a, b]
        same_arg = (k0[:-1] > k2[:-1]) == (k0[1:] > k2[1:])
        ok = same_mer & same_arg
        if not ok.any():
            continue
        m = np.abs(dm - delta)
        m[~ok] = np.inf
        g = int(np.argmin(m))
        best.append((float(m[g]), m0, m2, int(a[g]), float(dm[g])))
    best.sort()
    return best[:3]


def build_window(rng, segs):
    """segs: list of ('F' | 'R' | 'X', length in bases). Returns (ref, tig): F copies the next reference bases, R their reverse
    complement, X random bases that match nothing; every segment consumes its own stretch of the reference."""
    total = sum(n for _, n in segs)
    ref = synth.random_seq(rng, total)
    parts, pos = [], 0
    for kind, n in segs:
        piece = ref[pos:pos + n]
        parts.append(piece if kind == 'F' else synth.revcomp(piece) if kind == 'R' else synth.random_seq(rng, n))
        pos += n
    return ref, np.concatenate(parts)


def main():
    rng = np.random.default_rng(20260)
    out = {}
    # ---- raw frame with stream gaps
    if 'ties-only' in sys.argv:
        return main_ties(rng, out)
    ref, tig = build_window(rng, [('F', 500), ('X', 120), ('F', 400), ('R', 300)])
    tig = tig.copy()
    tig[230:241] = ord('N')
    mg.density_case('few_informative_gaps', ref, tig)
    d = os.path.join(HERE, 'density', 'few_informative_gaps')
    meta = json.load(open(os.path.join(d, 'meta.json')))
    rc, df, _ = mg._run_density(os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), meta['refregion'], meta['tigregion'])
    assert rc == 0 and list(df.columns) == ['KMER', 'INDEX', 'STATE', 'STATE_MER'] and not (df.index.to_numpy() == df['INDEX'].to_numpy()).all()
    meta['index'] = [int(x) for x in df.index.to_numpy()]
    meta['index_name'] = df.index.name
    json.dump(meta, open(os.path.join(d, 'meta.json'), 'w'), indent=1)
    main_ties(rng, out)


def main_ties(rng, out):
    # ---- near ties
    seen = set()
    for i, (margin, a, b, c, j) in enumerate(search_argmax(rng)):
        if (a, b, c) in seen:      # both slopes of one block can make the list: one window is enough
            continue
        seen.add((a, b, c))
        ref, tig = build_window(rng, [('F', a + K - 1), ('R', b + K - 1), ('F', c + K - 1)])
        name = f'tie_argmax_{"abc"[i]}'
        mg.density_case(name, ref, tig)
        out[name] = {'runs': [a, b, c], 'row': j, 'predicted_relative_margin': margin}
    for i, (margin, m0, m2, a, dm) in enumerate(search_delta(rng)):
        ref, tig = build_window(rng, [('F', m0 + K - 1), ('R', m2 + K - 1)])
        name = f'tie_delta_{"abc"[i]}'
        mg.density_case(name, ref, tig)
        out[name] = {'runs': [m0, m2], 'gap_start': a, 'predicted_max_delta': dm, 'predicted_margin_to_0.005': margin}
    ref, tig = build_window(rng, [('F', 1100 + K - 1), ('R', 22 + K - 1), ('F', 1000 + K - 1)])
    mg.density_case('spike_near_one', ref, tig)
    with open(os.path.join(HERE, 'density_near_ties.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__' and 'homology-wrap' not in sys.argv and 'kern-exact' not in sys.argv:
    main()


def homology_wrap_cases():
    """pavlib.call.right_homology with a negative position (Python's negative indexing wraps the first |pos| reads to the end of the
    sequence, the limit len - pos lets the scan run on from its start; pavlib/call.py:623-647) and both functions with an empty SV
    sequence (`x % 0`), as the reference itself answers them."""
    import pavlib.call
    rng = np.random.default_rng(77)
    out = []
    seqs = ['ACGTACGTACGTACGT', 'AAAAAAAAAA', 'CAGCAGCAGCAGTT', 'GATTACANGATTACA']
    for _ in range(12):
        unit = synth.random_seq(rng, int(rng.integers(1, 5))).tobytes().decode()
        seqs.append(unit * int(rng.integers(3, 40)) + synth.random_seq(rng, int(rng.integers(0, 30))).tobytes().decode())
    for seq in seqs:
        for sv in (seq[:1], seq[:3], seq[-2:], seq[-4:] + seq[:2], 'ACGT', ''):
            for pos in sorted({-1, -2, -3, -len(seq), -len(seq) - 1, -len(seq) // 2, 0, len(seq) - 1, len(seq)}):
                rec = {'seq': seq, 'sv': sv, 'pos': pos}
                for name, fn in (('left', pavlib.call.left_homology), ('right', pavlib.call.right_homology)):
                    try:
                        rec[name] = int(fn(pos, seq, sv))
                    except Exception as ex:  # noqa: BLE001
                        rec[name] = {'error': type(ex).__name__}
                out.append(rec)
    with open(os.path.join(HERE, 'homology_wrap.json'), 'w') as fh:
        json.dump(out, fh)
    print('homology_wrap', len(out), 'cases,', sum(isinstance(r['right'], dict) or isinstance(r['left'], dict) for r in out), 'with an exception')


if __name__ == '__main__' and 'homology-wrap' in sys.argv:
    homology_wrap_cases()


def density_kern_exact():
    """density.tsv.gz stores KERN_* as decimal text, and pandas' to_csv keeps 16 decimal places: values around 1e-4 carry 12-13
    significant digits there, which caps any comparison at ~1e-12 relative (found in r02: the scalar C oracle and the GPU both sat at
    ~1e-12 against the TSV while scipy itself is within 1e-14 of the exact sum). The reference is run again on every smoothed case
    and its three float64 columns are stored bit for bit (kern.npy, 3 x N); the discrete columns must equal the stored table."""
    import pandas as pd
    from concurrent.futures import ThreadPoolExecutor
    base = os.path.join(HERE, 'density')

    def one(case):
        d = os.path.join(base, case)
        meta = json.load(open(os.path.join(d, 'meta.json')))
        if meta['returncode'] != 0 or 'KERN_FWD' not in (meta.get('columns') or []):
            return case, None
        rc, df, _ = mg._run_density(os.path.join(d, 'ref.fa'), os.path.join(d, 'tig.fa'), meta['refregion'], meta['tigregion'], meta['k'], meta['rev'],
                                    meta['srs'], meta.get('extra', ()))
        gold = pd.read_csv(os.path.join(d, 'density.tsv.gz'), sep='\t')
        assert rc == 0 and all((df[c].to_numpy() == gold[c].to_numpy()).all() for c in ('INDEX', 'STATE_MER', 'STATE', 'KMER')), case
        k = np.stack([df[c].to_numpy(dtype=np.float64) for c in ('KERN_FWD', 'KERN_FWDREV', 'KERN_REV')])
        np.save(os.path.join(d, 'kern.npy'), k)
        return case, float(np.max(np.abs(k - np.stack([gold[c].to_numpy() for c in ('KERN_FWD', 'KERN_FWDREV', 'KERN_REV')]))))
    with ThreadPoolExecutor(max_workers=6) as ex:
        for case, err in ex.map(one, sorted(os.listdir(base))):
            print('kern.npy', case, 'max |text - binary| =', err)


if __name__ == '__main__' and 'kern-exact' in sys.argv:
    density_kern_exact()
