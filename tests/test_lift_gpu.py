"""Batched coordinate lifts on the device (`pavgpu_lift_index_create` / `pavgpu_lift_points` behind
`pavlib.lift.AlignLift.lift_points` / `lift_regions_to_qry`) against the answers of the reference's own `pavlib.align.AlignLift`
stored with the table (tests/golden/lift: 3,000 point lifts in both directions incl. gap=True and the positions the reference raises
on, 400 region lifts) and against the per-point host path on random tables."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from pav_b200 import synth

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plain(x):
    return None if x is None else [x[0], int(x[1]), bool(x[2]), int(x[3]), int(x[4]), [int(i) for i in x[5]]]


def test_device_lifts_match_reference_golden():
    from pav_b200.pavlib import lift, seq
    d = os.path.join(REPO, 'tests', 'golden', 'lift')
    df = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t')
    fai = pd.read_csv(os.path.join(d, 'tig.fai.tsv'), sep='\t', header=None, index_col=0)[1]
    al = lift.AlignLift(df, fai)
    queries = json.load(open(os.path.join(d, 'queries.json')))
    assert al._device_index() is not None
    n = 0
    for kind, to_qry, gap in (('to_qry', True, False), ('to_sub', False, False), ('to_sub', False, True)):
        qs = [q for q in queries if q['f'] == kind and (kind == 'to_qry' or q['gap'] == gap)]
        good = [q for q in qs if 'error' not in q]
        got = al.lift_points([q['id'] for q in good], [q['pos'] for q in good], to_qry, gap=gap)
        assert [_plain(g) for g in got] == [q['result'] for q in good]
        for q in qs:
            if 'error' in q:      # a position inside a record that no block covers: the reference raises
                with pytest.raises(RuntimeError):
                    al.lift_points([q['id']], [q['pos']], to_qry, gap=gap)
        n += len(qs)
    regions = [q for q in queries if q['f'] == 'region']
    got = al.lift_regions_to_qry([seq.Region(q['chrom'], q['pos'], q['end']) for q in regions])
    assert [None if r is None else [r.chrom, r.pos, r.end, bool(r.is_rev)] for r in got] == [q['qry'] for q in regions]
    assert n + len(regions) == 3400


def test_device_lifts_equal_host_lifts_on_random_tables():
    """Clipped, reverse, indel-rich records; one-base blocks; positions at block edges, in clips, outside every record, covered twice."""
    from pav_b200.pavlib import lift
    for seed, clip in ((5, (0, 0)), (6, (40, 25)), (7, (3, 0))):
        ref, tigs, df = synth.make_cigar_workload(seed, 2, 150_000, 9, 40_000, edit_rate=0.02, rev_frac=0.5, clip=clip)
        df = df.reset_index(drop=True)
        fai = pd.Series({k: len(v) for k, v in tigs.items()})
        host, dev = lift.AlignLift(df, fai), lift.AlignLift(df, fai)
        host._dev_index = False      # per-point host path
        rng = np.random.default_rng(seed)
        for to_qry in (True, False):
            ids, coords = [], []
            for _ in range(4000):
                row = df.iloc[int(rng.integers(0, df.shape[0]))]
                if to_qry:
                    ids.append(row['#CHROM']); coords.append(int(rng.integers(row['POS'] - 20, row['END'] + 20)))
                else:
                    ids.append(row['QRY_ID']); coords.append(int(rng.integers(max(row['QRY_POS'] - 60, 0), row['QRY_END'] + 60)))
            exp, bad = [], set()
            for k, (i, c) in enumerate(zip(ids, coords)):
                try:
                    exp.append(host.lift_to_qry(i, c) if to_qry else host.lift_to_sub(i, c))
                except RuntimeError:
                    exp.append('error'); bad.add(k)
            keep = [k for k in range(len(ids)) if k not in bad]
            got = dev.lift_points([ids[k] for k in keep], [coords[k] for k in keep], to_qry)
            assert [_plain(g) for g in got] == [_plain(exp[k]) for k in keep]
            for k in sorted(bad)[:25]:
                with pytest.raises(RuntimeError):
                    dev.lift_points([ids[k]], [coords[k]], to_qry)
            assert len(keep) > 3000
