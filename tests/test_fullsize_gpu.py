"""BASELINE configs[2] shape (2 haplotypes vs an hg38-shaped reference: 24 chromosomes of unequal length, soft-masked runs,
N blocks; CIGAR walk + density scan of the flagged windows) at 1/250 scale through the same driver that runs it at full size
(profiles/run_c3.py, result in profiles/): size-independent properties over every record (oracle/properties.py: row counts,
emission order, REF != ALT, decode(reference, rows) == contig) + oracle equality on sampled records and one window."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _driver():
    spec = importlib.util.spec_from_file_location('run_c3', os.path.join(REPO, 'profiles', 'run_c3.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('regime', ['human', 'stress'])
def test_c3_shape_scaled(regime):
    res = _driver().run(scale=0.004, regime=regime, steps=1, roundtrip_every=1, oracle_records=3, window_every=30_000, window_len=20_000,
                        density_chunk=256, do_density=True, contig_len=25_000_000)
    for hap in ('h1', 'h2'):
        w = res['haplotypes'][hap]['walk']
        assert w['properties']['roundtrip_records'] == w['records'] and w['oracle']['identical']
        assert w['rows'] > 1000
        d = res['haplotypes'][hap]['density']
        assert d['windows_ok'] > 0.9 * d['windows'] and d['oracle']['identical']
