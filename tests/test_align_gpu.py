"""SAM -> alignment table with the per-record CIGAR sums on the device: `pavgpu_cigar_record_stats` (one warp per
record over the packed ops of `pavgpu_cigar_parse`) against the reference's own table for a golden SAM (tests/golden/align/sam1,
written by the reference's get_align_bed on a stub SAM reader) and against the host sums on random CIGARs."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_get_align_bed_device_stats_matches_reference_golden():
    from pav_b200.pavlib import align, seq
    d = os.path.join(GOLDEN, 'align', 'sam1')
    meta = json.load(open(os.path.join(d, 'meta.json')))
    fai = seq.get_df_fai(os.path.join(d, 'tig.fa.fai'))
    dev = align.get_align_bed(os.path.join(d, 'align.sam'), fai, 'h1', min_mapq=meta['min_mapq'], device_stats=True)
    host = align.get_align_bed(os.path.join(d, 'align.sam'), fai, 'h1', min_mapq=meta['min_mapq'])
    assert dev.to_csv(sep='\t', index=False).encode() == open(os.path.join(d, 'align.bed'), 'rb').read()
    assert dev.equals(host) and [int(i) for i in dev.index] == meta['index'] and [str(t) for t in dev.dtypes] == meta['dtypes']
    with pytest.raises(RuntimeError, match='Found alignment match CIGAR operation'):
        align.get_align_bed(os.path.join(d, 'bad_m.sam'), fai, 'h1', device_stats=True)


def test_cigar_record_stats_equal_host_sums():
    """Random records: every op class, clips at the ends and (malformed) inside, clip-only records, one-op records, a 200,000-op
    record (several trips of the warp), an empty batch."""
    from pav_b200 import device
    from pav_b200.pavlib import align
    rng = np.random.default_rng(77)
    cigars = []
    for r in range(3000):
        n = int(rng.choice([1, 2, 3, 5, 17, 33, 64, 400]))
        body = ['%d%s' % (int(rng.integers(1, 5000)), 'MIDN=XP'[int(rng.choice([1, 2, 3, 4, 5, 4, 4, 5, 6, 0] if r % 50 == 0 else [1, 2, 4, 5, 4, 4]))])
                for _ in range(n)]
        lead = ['', '5H', '7S', '3H9S', '9S3H', '2S2S'][int(rng.integers(0, 6))]
        tail = ['', '4S', '6H', '8S1H', '1H8S'][int(rng.integers(0, 5))]
        mid = '11S' if r % 97 == 0 else ''
        cigars.append(lead + ''.join(body[:len(body) // 2]) + mid + ''.join(body[len(body) // 2:]) + tail)
    cigars += ['10H', '3S4H', '12=', '1X']
    cigars.append(''.join('%d%s' % (1 + i % 7, '=XID'[i % 4]) for i in range(200_000)))
    ops, op_off, perr = device.parse_cigars(cigars)
    assert perr.code == 0
    got = device.cigar_record_stats(ops, op_off)
    exp = align._record_stats_host((ops & 15).astype(np.int64), (ops >> 4).astype(np.int64), op_off)
    for f in got.dtype.names:
        assert np.array_equal(got[f], exp[f]), f
    assert len(device.cigar_record_stats(np.zeros(0, np.uint32), np.zeros(1, np.int64))) == 0
