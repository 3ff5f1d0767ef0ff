"""bench.py contract (CPU part): the reference arm prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--scale', '0.004', '--steps', '1', '--warmup', '0',
                          '--cpu-sample-records', '2', '--cpu-density-windows', '1'], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    lines = [ln for ln in out.stdout.decode().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['value'] > 0 and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] >= 1      # the unmodified reference (oracle/_ref or /root/reference)
    assert d['config']['workload'].startswith('C3') and d['scaling'] == 'strong'
    sec = d['secondary']                                                                       # Path B: scripts/density.py per window
    assert sec['metric'] == 'inv_kmer_density_gbases_per_sec' and sec['value'] > 0 and sec['cpu_baseline']['kind'] == 'reference'
    assert sec['cpu_baseline']['startup_seconds_per_process'] > 0
    assert d['cpu_baseline']['port']['kind'] == 'port' and d['cpu_baseline']['port']['value'] > 0      # the oracle port beside the reference
    assert sec['cpu_baseline']['port']['value'] > 0
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_bench_helpers():
    """Pure helpers of bench.py: the byte model per kernel, the gather floor from the committed microbenchmark, the LPT plan."""
    import importlib.util

    import numpy as np
    import pandas as pd
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(REPO, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    k = bench.kernel_table(0.02, 0.04, 0.07, 4, 1000, 4, 300, 50)
    assert k['homology_queue_kernel'] == (0.07, 256 * 50) and k['cigar_walk_kernel'][1] == 4 * 1000 + 32 * 4 + 16 * 300 + 64 * 50
    g = bench.gather_bound(370_433, 0.0727)
    assert g is not None and 40 < g['peak_gaccesses_per_s'] < 60 and 0.03 < g['floor_ms'] < 0.045 and 0.4 < g['frac'] < 0.7
    rng = np.random.default_rng(3)
    dfs = {h: pd.DataFrame({'POS': 0, 'END': rng.integers(10_000, 5_000_000, 40), 'CIGAR': ['10=' * int(n) for n in rng.integers(1, 400, 40)]}) for h in ('h1', 'h2')}
    plan = bench.shard_plan(dfs, 4)
    for h in dfs:
        assert sorted(np.concatenate(plan[h]).tolist()) == list(range(40))
