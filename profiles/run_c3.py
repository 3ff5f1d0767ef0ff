#!/usr/bin/env python3
"""BASELINE configs[2] at full size on one B200: 2 haplotypes (h1/h2) of 10 Mbp contigs against a 3.1 Gbp hg38-shaped
reference (24 chromosomes, 50 % soft-masked, 5 % N), CIGAR walk + k=31 inversion density scan on the flagged windows
(one 50 kbp window per 300 kbp of contig), with the size-independent parity properties of oracle/properties.py and an
oracle comparison on a sample of records / windows.

    python profiles/run_c3.py [--scale 1.0] [--regime human|stress] [--out profiles/rNN_c3.json]

Not a bench line (bench.py measures configs[1]); this is the full-size parity + throughput record for configs[2].
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import properties as checks  # noqa: E402
from pav_b200 import _capi, device, synth  # noqa: E402
from pav_b200.pavlib import density  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def lift_ref_to_qry(ops_rec, ref_off):
    """Oriented query offset of reference offsets ``ref_off`` (relative to POS) inside one record, from the CIGAR prefix sums."""
    code, ln = ops_rec & 15, (ops_rec >> 4).astype(np.int64)
    adv_r = np.where(np.isin(code, (2, 7, 8)), ln, 0)
    adv_q = np.where(np.isin(code, (1, 4, 5, 7, 8)), ln, 0)
    cr, cq = np.concatenate(([0], np.cumsum(adv_r))), np.concatenate(([0], np.cumsum(adv_q)))
    k = np.clip(np.searchsorted(cr, ref_off, side='right') - 1, 0, len(ops_rec) - 1)
    inside = np.where(adv_q[k] > 0, np.minimum(ref_off - cr[k], adv_q[k]), 0)
    return cq[k] + inside


def run_walk(ctx, ref_store, ref_names, ref, hap, tigs, df, steps, roundtrip_every, oracle_records):
    names_t = list(tigs)
    t0 = time.perf_counter()
    tig_store = device.SeqStore(ctx, names_t, [tigs[n] for n in names_t], keep_host=False)
    t_store = time.perf_counter() - t0
    rid = np.array([ref_names.index(c) for c in df['#CHROM']], np.int32)
    tidx = {n: i for i, n in enumerate(names_t)}
    qid = np.array([tidx[c] for c in df['QRY_ID']], np.int32)
    t0 = time.perf_counter()
    ops, op_off, perr = device.parse_cigars(df['CIGAR'].tolist())
    t_parse = time.perf_counter() - t0
    assert perr.code == 0
    batch = device.CigarBatch(ctx, rid, qid, df['POS'].to_numpy(np.int32), df['REV'].to_numpy(np.uint8), ops, op_off)
    ms, st = [], None
    for i in range(2 + steps):
        ctx.l2_flush()
        st = batch.run(ref_store, tig_store)
        if i >= 2:
            ms.append((st.ms_kernels, st.ms_scan, st.ms_homology))
    t0 = time.perf_counter()
    snv, indel, cerr = batch.fetch()
    t_fetch = time.perf_counter() - t0
    assert cerr.code == 0
    batch.close()
    k_ms, walk_ms, hom_ms = (float(np.mean([m[i] for m in ms])) for i in range(3))
    n_rows = len(snv) + len(indel)
    log(f'[{hap}] {len(df)} records, {len(ops)} ops, {n_rows} rows: {k_ms:.3f} ms/step (walk {walk_ms:.3f}, homology {hom_ms:.3f}) '
        f'= {n_rows / k_ms * 1e3:.3e} rows/s; contig upload+pack {t_store:.2f}s, CIGAR tokenizer {t_parse:.2f}s, D2H {t_fetch:.2f}s')
    # ---- size-independent properties over the whole haplotype
    t0 = time.perf_counter()
    rows = list(zip(df['POS'], df['END'], df['REV']))
    prop = checks.check_walk(rows, ops, op_off, snv, indel, [ref[c] for c in df['#CHROM']], [tigs[q] for q in df['QRY_ID']],
                             roundtrip_every=roundtrip_every)
    prop['seconds'] = time.perf_counter() - t0
    log(f'[{hap}] properties ok: {prop}')
    # ---- oracle on a sample of records (those on the smallest chromosome, so the FASTA to write stays small)
    from oracle import pyoracle
    small = min(ref, key=lambda c: len(ref[c]))
    pick = np.flatnonzero((df['#CHROM'] == small).to_numpy())[:oracle_records]
    tmp = tempfile.mkdtemp(prefix='c3_oracle_')
    sub = df.iloc[pick]
    ref_fa, tig_fa, _ = synth.write_cigar_workload(tmp, {small: ref[small]}, {q: tigs[q] for q in sub['QRY_ID']}, sub)
    t0 = time.perf_counter()
    o_snv, o_indel, _ = pyoracle.walk_rows(sub, ref_fa, tig_fa)
    t_or = time.perf_counter() - t0
    g_snv, g_indel = snv[np.isin(snv['rec'], pick)], indel[np.isin(indel['rec'], pick)]
    same = (len(g_snv) == len(o_snv) and (g_snv['pos_ref'] == o_snv['pos_ref']).all() and (g_snv['qry_pos'] == o_snv['qry_pos']).all()
            and len(g_indel) == len(o_indel)
            and all((g_indel[c] == o_indel[c]).all() for c in ('pos', 'end', 'svlen', 'qry_pos', 'qry_end', 'left_shift', 'hom_ref_l', 'hom_ref_r',
                                                                 'hom_tig_l', 'hom_tig_r')))
    assert same, f'[{hap}] rows of the sampled records differ from the oracle'
    log(f'[{hap}] oracle: {len(pick)} records on {small} ({len(o_snv) + len(o_indel)} rows) identical; oracle walk {t_or:.2f}s '
        f'= {(len(o_snv) + len(o_indel)) / t_or:.3e} rows/s on one core')
    out = {'records': int(len(df)), 'ops': int(len(ops)), 'rows': int(n_rows), 'snv_rows': int(len(snv)), 'indel_rows': int(len(indel)),
           'ms_per_step': k_ms, 'ms_walk': walk_ms, 'ms_homology': hom_ms, 'rows_per_s': n_rows / k_ms * 1e3, 'steps': steps,
           'contig_bases': int(sum(len(v) for v in tigs.values())), 'seconds_contig_upload_pack': t_store, 'seconds_cigar_tokenizer': t_parse,
           'seconds_d2h': t_fetch, 'properties': prop,
           'oracle': {'records': int(len(pick)), 'rows': int(len(o_snv) + len(o_indel)), 'identical': True, 'rows_per_s_one_core': (len(o_snv) + len(o_indel)) / t_or}}
    return out, tig_store, names_t, ops, op_off, rid, qid


def run_density(ctx, ref_store, ref, hap, tigs, df, tig_store, names_t, ops, op_off, rid, qid, win_len, every, chunk, oracle_windows):
    k = 31
    wins, expect = [], []
    for r in range(len(df)):
        pos0, end0, rev = int(df['POS'].iloc[r]), int(df['END'].iloc[r]), bool(df['REV'].iloc[r])
        tlen = len(tigs[names_t[qid[r]]])
        starts = np.arange(every // 2, end0 - pos0 - win_len, every, dtype=np.int64)
        if not len(starts):
            continue
        q0 = lift_ref_to_qry(ops[op_off[r]:op_off[r + 1]], starts)
        for s, q in zip(starts.tolist(), q0.tolist()):
            q = min(q, tlen - win_len)
            if q < 0:
                continue
            tp = tlen - (q + win_len) if rev else q     # forward contig coordinates of the oriented window
            wins.append((rid[r], qid[r], pos0 + s, pos0 + s + win_len, tp, tp + win_len, int(rev), 20))
    win = np.array(wins, dtype=_capi.DENSITY_WINDOW)
    log(f'[{hap}] density: {len(win)} windows x {win_len} bp')
    params = density.default_params()
    tot_ms = tot_kmer = tot_kde = 0.0
    n_ok = n_fail = rows = 0
    checked = failed = None
    for a in range(0, len(win), chunk):
        sub = win[a:a + chunk]
        batch = density.DensityBatch(ctx, sub, params)
        ctx.l2_flush()
        st = batch.run(ref_store, tig_store)
        tot_ms += st.ms_kernels
        tot_kmer += st.ms_kmer
        tot_kde += st.ms_kde
        rows += int(st.rows)
        res, cols = batch.fetch()
        batch.close()
        for j, d in enumerate(density._split(res, cols)):
            ok = checks.check_density_window(d, win_len, k, expect_state=0, min_frac=0.5)
            n_ok += ok
            n_fail += not ok
            if checked is None and ok and d['smoothed']:
                checked = (sub[j], d)
            if failed is None and not ok:
                failed = sub[j]
    bases = len(win) * win_len
    out = {'windows': int(len(win)), 'window_bp': win_len, 'bases': int(bases), 'ms': tot_ms, 'ms_kmer': tot_kmer, 'ms_kde': tot_kde,
           'gbases_per_s': bases / (tot_ms * 1e-3) / 1e9 if tot_ms else None, 'rows': rows, 'windows_ok': int(n_ok), 'windows_status_125': int(n_fail),
           'chunk': chunk}
    log(f'[{hap}] density: {tot_ms:.2f} ms for {bases / 1e6:.0f} Mbases = {out["gbases_per_s"]:.2f} Gbases/s (k-mer {tot_kmer:.2f} ms, KDE {tot_kde:.2f} ms); '
        f'{n_ok} windows ok with FWD the dominant state, {n_fail} soft failures (exit 125 in the reference)')
    if checked is not None and oracle_windows:
        from oracle import pyoracle
        w, d = checked
        rseq = ref[list(ref)[w['ref_seq_id']]][w['ref_pos']:w['ref_end']]
        tseq = tigs[names_t[w['tig_seq_id']]][w['tig_pos']:w['tig_end']]
        t0 = time.perf_counter()
        rc, o = pyoracle.density_arrays(rseq.tobytes(), tseq.tobytes(), rev=bool(w['rev']))
        dt = time.perf_counter() - t0
        same = rc == 0 and all((d[c].astype(np.int64) == o[c].astype(np.int64)).all() for c in ('KMER', 'INDEX', 'STATE_MER', 'STATE'))
        assert same, f'[{hap}] density window differs from the oracle'
        out['oracle'] = {'windows': 1, 'identical': True, 'gbases_per_s_one_core': win_len / dt / 1e9}
        log(f'[{hap}] density oracle: 1 window identical ({dt:.2f}s on one core)')
        if failed is not None:   # a soft failure must be one in the oracle too (scripts/density.py:510-527 -> exit 125)
            w = failed
            rseq = ref[list(ref)[w['ref_seq_id']]][w['ref_pos']:w['ref_end']]
            tseq = tigs[names_t[w['tig_seq_id']]][w['tig_pos']:w['tig_end']]
            rc, _ = pyoracle.density_arrays(rseq.tobytes(), tseq.tobytes(), rev=bool(w['rev']))
            assert rc == 125, f'[{hap}] window reported as a soft failure, oracle says rc={rc}'
            out['oracle']['soft_failure_confirmed'] = True
            log(f'[{hap}] density oracle: first soft-failure window is exit 125 in the oracle too')
    return out


def run(scale=1.0, regime='human', steps=5, roundtrip_every=1, oracle_records=2, window_every=300_000, window_len=50_000,
        density_chunk=1024, do_density=True, contig_len=10_000_000):
    t0 = time.perf_counter()
    ref, trs = synth.config_c3_reference(scale=scale)
    haps = {h: synth.config_c3_haplotype(ref, trs, h, scale=scale, regime=regime, contig_len=contig_len) for h in ('h1', 'h2')}
    synth.config_c3_mask(ref)
    log(f'generated in {time.perf_counter() - t0:.1f}s: reference {sum(len(v) for v in ref.values()) / 1e9:.3f} Gbp, '
        + ', '.join(f'{h}: {len(d[1])} contigs' for h, d in haps.items()))
    ctx = device.get_context()
    ref_names = list(ref)
    t0 = time.perf_counter()
    ref_store = device.SeqStore(ctx, ref_names, [ref[n] for n in ref_names], keep_host=False)
    t_ref = time.perf_counter() - t0
    log(f'reference uploaded + packed in {t_ref:.2f}s')
    result = {'config': f'C3: 2 haplotypes x {int(contig_len * scale)} bp contigs vs {sum(len(v) for v in ref.values())} bp hg38-shaped reference '
                        f'(scale {scale}), regime {regime}', 'scale': scale, 'regime': regime, 'seconds_reference_upload_pack': t_ref, 'haplotypes': {}}
    for hap, (tigs, df) in haps.items():
        w, tig_store, names_t, ops, op_off, rid, qid = run_walk(ctx, ref_store, ref_names, ref, hap, tigs, df, steps, roundtrip_every, oracle_records)
        entry = {'walk': w}
        if do_density:
            entry['density'] = run_density(ctx, ref_store, ref, hap, tigs, df, tig_store, names_t, ops, op_off, rid, qid, window_len, window_every,
                                           density_chunk, True)
        tig_store.close()
        result['haplotypes'][hap] = entry
    ref_store.close()
    rows = sum(e['walk']['rows'] for e in result['haplotypes'].values())
    ms = sum(e['walk']['ms_per_step'] for e in result['haplotypes'].values())
    result['walk_rows_total'] = rows
    result['walk_rows_per_s'] = rows / ms * 1e3
    if do_density:
        b = sum(e['density']['bases'] for e in result['haplotypes'].values())
        m = sum(e['density']['ms'] for e in result['haplotypes'].values())
        result['density_gbases_per_s'] = b / (m * 1e-3) / 1e9 if m else None
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--regime', default='human', choices=list(synth.C3_REGIMES))
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--roundtrip-every', type=int, default=1)
    ap.add_argument('--oracle-records', type=int, default=2)
    ap.add_argument('--window-every', type=int, default=300_000)
    ap.add_argument('--window-len', type=int, default=50_000)
    ap.add_argument('--density-chunk', type=int, default=1024)
    ap.add_argument('--no-density', action='store_true')
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    from oracle import pyoracle
    from pav_b200 import build
    build.build()
    pyoracle.build()
    result = run(args.scale, args.regime, args.steps, args.roundtrip_every, args.oracle_records, args.window_every, args.window_len,
                 args.density_chunk, not args.no_density)
    print(json.dumps(result))
    if args.out:
        with open(args.out, 'w') as fh:
            json.dump(result, fh, indent=1)


if __name__ == '__main__':
    main()
