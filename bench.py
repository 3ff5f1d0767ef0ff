#!/usr/bin/env python3
"""Benchmark of the PAV hot path on B200 (driver contract: ONE JSON line on stdout from rank 0).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--metric walk|density]

Workloads (BASELINE.json configs, SURVEY 8(d) generators, all synthetic and seeded):

  C3  (top-level line, every N)  2 phased haplotypes (h1/h2) of 10 Mbp contigs against a 3.1 Gbp hg38-shaped reference
      (24 chromosomes, 50 % soft-masked, 5 % N), human-like edit regime (1 SNV / kbp + 1 indel / 5 kbp). With N GPUs the
      alignment records of both haplotypes are sharded over the ranks by longest-processing-time-first, rank 0 packs the
      reference and broadcasts the packed planes with one NCCL broadcast, every rank checks the planes it received against rank
      0's checksum -- BASELINE configs[2] at N = 1, configs[3] at N > 1: STRONG scaling (total work fixed).
        value  device-resident: packed planes, packed CIGAR ops and record descriptors already in HBM; one step = per-record
               row counts + record scan + CIGAR walk + homology for this rank's records of both haplotypes (one batch), replayed as one
               CUDA graph, timed with CUDA events on the library stream, L2 flushed (256 MB memset) between steps; rows stay in HBM.
               value = rows of all ranks / max over ranks of the mean step time.
        e2e    the call PAV makes, FASTA in -> DataFrames out, wall clock, max over ranks: pavlib.cigarcall.make_insdel_snv_calls per
               haplotype at N = 1; at N > 1 every rank makes that call on the chromosomes it owns
               (pav_b200.multigpu.make_insdel_snv_calls_shard: LPT over chromosomes, 1 / N of the reference read, uploaded and packed per
               rank, no communication, tables concatenate).
        e2e_dist (N > 1)  pav_b200.multigpu.make_insdel_snv_calls_dist: records LPT-sharded over ranks -> NCCL broadcast of the packed
               reference -> per-rank walk -> host gather -> ONE merged table formatted on rank 0.
  C5  (`secondary`, every N; top level with --metric density)  10,000 flagged 50 kbp windows, k = 31 (BASELINE configs[4]),
      split evenly over the ranks: inv k-mer Gbases/s device-resident and through pavlib.density.density_windows(lazy=True)
      (ASCII windows in host memory -> run lengths of STATE in host memory, the columns stay in HBM until a window becomes a call).
  C2  (`c2`, N = 1 only)  the round-1 line kept for continuity: 1 haplotype, 1,000 contigs x 200 kbp vs 200 Mbp, 1 edit / 100 bp;
      the per-kernel roofline with ncu traffic (captures under profiles/ are taken on this input) lives here.

--impl reference times the UNMODIFIED reference (PAV 2.4.6.0, staged byte for byte under oracle/_ref by oracle/stage_ref.py, run
with the stub third-party modules of oracle/ref_stubs) on the box's host cores: pavlib.cigarcall.make_insdel_snv_calls over a
seeded sample of the C3 records, one record per worker process per step, and scripts/density.py spawned per window exactly like
pavlib/inv.py:249-266 does (-t 1, one window per worker, all cores), start-up reported separately.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'cigar_walk_variant_records_per_sec'
UNIT = 'variant rows/s'
METRIC_B = 'inv_kmer_density_gbases_per_sec'
UNIT_B = 'Gbases/s'
WORKLOAD_C2 = 'C2: 1 haplotype, 1000 contigs x 200 kbp vs 200 Mbp reference (4 x 50 Mbp), 1 edit/100 bp, 50% REV'
WIN_LEN = 50_000
HOM_NAMES = ['homology_kernel', 'homology_tiled_kernel', 'homology_nbr_kernel', 'homology_bulk_kernel', 'homology_queue_kernel', 'homology_split_kernels']


def workload_c3(args, world):
    sc = '' if args.scale == 1.0 else f' [scale {args.scale}]'
    return (f'C3{sc}: 2 phased haplotypes (h1/h2) of 10 Mbp contigs vs 3.1 Gbp hg38-shaped reference (24 chromosomes, 50% soft-masked, 5% N), '
            f'{args.regime} edit regime, records sharded by alignment record (LPT) over {world} GPU(s)'
            + (', one NCCL broadcast of the packed reference' if world > 1 else ''))


def workload_c5(args, world):
    return f'C5: {args.c5_windows} flagged windows x 50 kbp, k=31, srs=20, split over {world} GPU(s)'


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--metric', default='walk', choices=['walk', 'density'], help='which path is the top-level line (the other one is `secondary`)')
    ap.add_argument('--scale', type=float, default=1.0, help='C3 size factor (1.0 = 3.1 Gbp; tests use 0.004)')
    ap.add_argument('--regime', default='human', choices=['human', 'stress'])
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--c5-windows', type=int, default=10_000, help='windows of the Path-B sweep over all ranks (0 = skip)')
    ap.add_argument('--c2', type=int, default=1, help='1: also run the C2 continuity leg at N = 1')
    ap.add_argument('--c2-contigs', type=int, default=1000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-records', type=int, default=0, help='C3 records in the CPU sample (0 = one per core)')
    ap.add_argument('--cpu-density-windows', type=int, default=0,
                    help='windows of the CPU density leg (0 = max(64, cores) for --impl reference, one per core otherwise)')
    ap.add_argument('--seed', type=int, default=1003)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


class Control:
    """Control plane: torch.distributed (gloo) when WORLD_SIZE > 1, no-ops otherwise. No tensor of the data path goes through it."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.dist = None
        if world > 1:
            import datetime

            import torch.distributed as dist
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('gloo', rank=rank, world_size=world, timeout=datetime.timedelta(minutes=30))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def gather(self, obj):
        """list with every rank's ``obj`` (on every rank)."""
        if not self.dist:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def max(self, v):
        return max(self.gather(float(v)))

    def sum(self, v):
        return sum(self.gather(float(v)))

    def bcast(self, obj):
        if not self.dist:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


@contextlib.contextmanager
def stdout_to_stderr():
    """Route C-level writes to fd 1 (e.g. NCCL's "NCCL version ..." banner) to stderr for the duration of the block."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix='clocks_', suffix='.csv')
        self.proc = None
        try:
            self.fh = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50'], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            tok = [t.strip() for t in line.split(',')]
            if len(tok) < 9:
                continue
            try:
                sm.append(float(tok[1]))
                mx.append(float(tok[2]))
            except ValueError:
                continue
            for nm, v in zip(names, tok[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak_gbs():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def read_align(path):
    import pandas as pd
    return pd.read_csv(path, sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)


# ----------------------------------------------------------------------------------------------------------
# Workload files (rank 0 writes, every rank reads)
# ----------------------------------------------------------------------------------------------------------
def make_c3_files(args, tmp, only_records=None):
    """Generate C3 and write ref.fa, {h1,h2}_tig.fa, {h1,h2}_align.bed under ``tmp``. ``only_records``: {hap: record indices} --
    write only those records, their contigs and their chromosomes (the CPU sample of --impl reference)."""
    from pav_b200 import synth
    t0 = time.perf_counter()
    ref, trs = synth.config_c3_reference(seed=args.seed, scale=args.scale)
    haps = {h: synth.config_c3_haplotype(ref, trs, h, seed=args.seed, scale=args.scale, regime=args.regime) for h in ('h1', 'h2')}
    synth.config_c3_mask(ref, seed=args.seed)
    t1 = time.perf_counter()
    keep_chrom = set()
    n_rec = {h: len(d[1]) for h, d in haps.items()}
    for h, (tigs, df) in haps.items():
        if only_records is not None:
            df = df.iloc[only_records[h]]
            tigs = {q: tigs[q] for q in df['QRY_ID']}
            keep_chrom.update(df['#CHROM'])
        synth.write_fasta(os.path.join(tmp, f'{h}_tig.fa'), tigs)
        df.to_csv(os.path.join(tmp, f'{h}_align.bed'), sep='\t', index=False)
    ref_bp = int(sum(len(v) for v in ref.values()))
    if only_records is not None:
        ref = {c: a for c, a in ref.items() if c in keep_chrom}
    synth.write_fasta(os.path.join(tmp, 'ref.fa'), ref)
    log(f'C3 generated in {t1 - t0:.1f}s, written in {time.perf_counter() - t1:.1f}s: reference {ref_bp / 1e9:.3f} Gbp, records {n_rec}')
    return {'reference_bp': ref_bp, 'records': n_rec, 'seconds_generate': t1 - t0, 'seconds_write': time.perf_counter() - t1}


def c3_record_counts(args):
    """Records per C3 haplotype without generating it (one per contig length of every chromosome, tails under 20 kbp dropped)."""
    from pav_b200 import synth
    contig_len = max(int(10_000_000 * args.scale), 40_000)
    n = 0
    for x in synth.HG38_LENGTHS:
        ln = max(int(x * args.scale), 40_000)
        for start in range(0, ln, contig_len):
            if min(contig_len, ln - start) < 20_000:
                break
            n += 1
    return {'h1': n, 'h2': n}


def c5_windows_range(seed, lo, hi):
    """Windows lo..hi-1 of the C5 sweep, each from its own seeded generator (so any rank can make its share): 30 % with 1-3 kbp
    inverted-repeat flanks, 0.5 % divergence, 10 % negative controls, inversion of 2-20 kbp."""
    from pav_b200 import synth
    out = []
    for w in range(lo, hi):
        rng = np.random.default_rng([seed, 0xC5, w])
        neg = bool(rng.random() < 0.1)
        flank = int(rng.integers(1000, 3001)) if rng.random() < 0.3 else 0
        r, t, iv = synth.make_inv_window(rng, WIN_LEN, None, flank, 0.005, neg)
        out.append((r, t, iv, neg))
    return out


# ----------------------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference
# ----------------------------------------------------------------------------------------------------------
_POOL = None


def _pool(cores):
    global _POOL
    if _POOL is None:
        import multiprocessing as mp
        _POOL = mp.get_context('fork').Pool(cores)
    return _POOL


def _ref_shard(task):
    """One worker = one Snakemake-style job: the reference's own make_insdel_snv_calls on its records (rules/call.snakefile:810)."""
    bed, rows, ref_fa, tig_fa, hap = task
    from oracle import refenv
    refenv.activate()
    import pavlib.cigarcall
    import pysam
    pysam._CACHE.clear()      # nothing cached between steps, like a fresh job
    df = read_align(bed).iloc[rows]
    a, b = pavlib.cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, hap, version_id=False)
    return len(a) + len(b)


def reference_walk_step(tasks, cores):
    pool = _pool(cores)
    t0 = time.perf_counter()
    counts = pool.map(_ref_shard, tasks, chunksize=1)
    dt = time.perf_counter() - t0
    return int(sum(counts)), dt


def _port_shard(task):
    """The oracle's C restatement of the walk + the reference's own container idiom for the frames (pd.Series per variant, concat):
    the `port` figure kept beside the reference's (round 1's CPU arm)."""
    bed, rows, ref_fa, tig_fa, hap = task
    from oracle import pyoracle
    df = read_align(bed).iloc[rows]
    a, b = pyoracle.make_insdel_snv_calls(df, ref_fa, tig_fa, hap, version_id=False, reference_containers=True)
    return len(a) + len(b)


def port_walk_step(tasks, cores):
    pool = _pool(cores)
    t0 = time.perf_counter()
    counts = pool.map(_port_shard, tasks, chunksize=1)
    dt = time.perf_counter() - t0
    return int(sum(counts)), dt


def _port_density(w):
    from oracle import pyoracle
    rc, _ = pyoracle.density_arrays(w[0].tobytes(), w[1].tobytes())
    return rc


def port_density(windows, cores):
    """The oracle's C restatement of scripts/density.py, one window per worker. Returns (Gbases/s, seconds)."""
    pool = _pool(cores)
    t0 = time.perf_counter()
    pool.map(_port_density, windows, chunksize=1)
    dt = time.perf_counter() - t0
    return sum(len(w[1]) for w in windows) / dt / 1e9, dt


def reference_density(windows, tmp, cores):
    """scripts/density.py per window, spawned like pavlib/inv.py:249-266 (-t 1), ``cores`` processes at a time. Returns
    (Gbases/s, seconds, start-up seconds per process, return codes)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import refenv
    from pav_b200 import synth
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join(refenv.pythonpath_entries())
    script = os.path.join(refenv.REF_ROOT, 'scripts', 'density.py')
    jobs = []
    for i, (r, t, _, _) in enumerate(windows):
        d = os.path.join(tmp, f'dw{i}')
        os.makedirs(d, exist_ok=True)
        synth.write_fasta(os.path.join(d, 'ref.fa'), {'chrW': r})
        synth.write_fasta(os.path.join(d, 'tig.fa'), {'tigW': t})
        jobs.append([sys.executable, script, '--tigregion', f'tigW:1-{len(t)}', '--refregion', f'chrW:1-{len(r)}', '--ref', os.path.join(d, 'ref.fa'),
                     '--tig', os.path.join(d, 'tig.fa'), '-k', '31', '-t', '1', '-r', 'false', '--staterunsmooth', '20'])

    def run(cmd):
        return subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env).returncode
    t0 = time.perf_counter()
    subprocess.run([sys.executable, '-c', 'import scipy.stats, pandas, numpy, pavlib, kanapy.util.kmer'], env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    startup = time.perf_counter() - t0
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        rcs = list(ex.map(run, jobs))
    dt = time.perf_counter() - t0
    bases = sum(len(w[1]) for w in windows)
    return bases / dt / 1e9, dt, startup, rcs


def cpu_sample(args, cores, n_rec):
    """{hap: record indices}: a seeded sample of the C3 records, one per core by default, alternating haplotypes."""
    n = args.cpu_sample_records or cores
    rng = np.random.default_rng([args.seed, 0x5A])
    pick = {'h1': [], 'h2': []}
    order = {h: rng.permutation(n_rec[h]).tolist() for h in ('h1', 'h2')}
    for i in range(n):
        h = 'h1' if i % 2 == 0 else 'h2'
        if order[h]:
            pick[h].append(order[h].pop())
    return {h: sorted(v) for h, v in pick.items()}


def run_reference(args, rank, world):
    """--impl reference: the unmodified reference on the host cores, rank 0 only."""
    if rank != 0:
        return
    from oracle import refenv
    if not refenv.available():
        print(json.dumps({'impl': 'reference', 'unavailable': 'neither /root/reference nor the staged copy oracle/_ref is present '
                                                              '(oracle/stage_ref.py stages it in the build container)'}), flush=True)
        return
    cores = len(os.sched_getaffinity(0))
    tmp = tempfile.mkdtemp(prefix='pavbench_ref_')
    n_rec = c3_record_counts(args)
    pick = cpu_sample(args, cores, n_rec)
    info = make_c3_files(args, tmp, only_records=pick)     # only the sample's contigs and chromosomes are written
    tasks = []
    for h in ('h1', 'h2'):
        bed = os.path.join(tmp, f'{h}_align.bed')
        tasks += [(bed, [i], os.path.join(tmp, 'ref.fa'), os.path.join(tmp, f'{h}_tig.fa'), h) for i in range(len(pick[h]))]
    vals, rows = [], 0
    for i in range(args.warmup + args.steps):
        rows, dt = reference_walk_step(tasks, cores)
        if i >= args.warmup:
            vals.append((rows / dt, dt))
        log(f'[reference] step {i}: {rows} rows in {dt:.2f}s -> {rows / dt:.0f} rows/s on {min(cores, len(tasks))} processes')
        if i >= args.warmup and sum(d for _, d in vals) > 120 and len(vals) >= 3:
            log('[reference] time budget reached: stopping the timed steps early')
            break
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals]) * 1e3)
    sample = (f'{len(tasks)} of {sum(info["records"].values())} C3 records (seeded sample, one per worker process; {rows} variant rows per step): FASTA read '
              f'by .fai offset + the reference\'s own make_insdel_snv_calls (pavlib/cigarcall.py:24-362, unmodified, oracle/_ref) per step; '
              f'{len(vals)} timed steps')
    # the port figure beside it: the oracle's C walk + reference-style frame assembly on the same sample, same pool
    from oracle import pyoracle
    pyoracle.build()
    p_rows, p_dt = port_walk_step(tasks, cores)
    port = {'value': p_rows / p_dt, 'unit': UNIT, 'kind': 'port', 'seconds': p_dt,
            'what': 'oracle/pav_oracle.c walk + the reference\'s container idiom for the frames (pd.Series per variant + concat), same records, same pool'}
    log(f'[reference] port on the same sample: {p_rows} rows in {p_dt:.2f}s -> {p_rows / p_dt:.0f} rows/s')
    sec = None
    if args.c5_windows > 0:
        n_w = min(args.cpu_density_windows or max(64, cores), args.c5_windows)
        wins = c5_windows_range(1005, 0, n_w)
        gb, dt, startup, rcs = reference_density(wins, tmp, cores)
        pgb, pdt = port_density(wins, cores)
        sec = {'impl': 'reference', 'metric': METRIC_B, 'value': gb, 'unit': UNIT_B, 'seconds': dt, 'cores': cores,
               'config': {'workload': workload_c5(args, args.gpus)},
               'cpu_baseline': {'value': gb, 'unit': UNIT_B, 'cores': cores, 'kind': 'reference',
                                'sample': f'first {n_w} of {args.c5_windows} C5 windows, one `python3 scripts/density.py ... -t 1` process per window '
                                          f'(spawned like pavlib/inv.py:249-266), {cores} at a time; interpreter + import start-up ({startup:.2f}s per '
                                          f'process, measured separately) is inside the time',
                                'startup_seconds_per_process': startup, 'return_codes': {str(c): rcs.count(c) for c in set(rcs)},
                                'port': {'value': pgb, 'unit': UNIT_B, 'kind': 'port', 'seconds': pdt,
                                         'what': 'oracle/pav_oracle.c restatement of scripts/density.py, same windows, one per worker, in-process'}}}
        log(f'[reference] density: {n_w} windows in {dt:.1f}s = {gb:.3e} Gbases/s on {cores} cores (start-up {startup:.2f}s per process)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(vals), 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic',
        'config': {'workload': workload_c3(args, args.gpus), 'reference_bp': info['reference_bp'], 'records': info['records']},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': min(cores, len(tasks)), 'kind': 'reference', 'sample': sample, 'port': port},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'secondary': sec,
    }
    if args.metric == 'density' and sec is not None:
        top = dict(sec)
        top.update({'n_gpus': args.gpus, 'steps': 1, 'warmup': 0, 'ms_per_step': sec['seconds'] * 1e3, 'higher_is_better': True, 'scaling': 'strong',
                    'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                    'e2e': {'value': sec['value'], 'unit': UNIT_B, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                    'secondary': {k: v for k, v in line.items() if k != 'secondary'}})
        line = top
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# Our arm
# ----------------------------------------------------------------------------------------------------------
def shard_plan(dfs, world):
    """{hap: list of index arrays, one per rank}: LPT over the records of both haplotypes together (one cost model, one balance)."""
    from pav_b200 import multigpu
    costs, owner = [], []
    for h, df in dfs.items():
        span = (df['END'] - df['POS']).to_numpy(np.int64)
        costs.append(multigpu.record_costs(df['CIGAR'].tolist(), span))
        owner += [(h, i) for i in range(len(df))]
    shards = multigpu.lpt_shards(np.concatenate(costs), world)
    plan = {h: [[] for _ in range(world)] for h in dfs}
    for r, idx in enumerate(shards):
        for j in idx.tolist():
            h, i = owner[j]
            plan[h][r].append(i)
    return {h: [np.array(sorted(v), dtype=np.int64) for v in plan[h]] for h in plan}


class ResidentShard:
    """This rank's records of BOTH haplotypes resident in HBM as one batch (contig planes of both haplotypes in one store + packed
    ops): one step of the rank is one replay of one CUDA graph."""

    def __init__(self, ctx, dfs, plan_rank, tig_fa, ref_names):
        import pandas as pd

        from pav_b200 import device
        from pav_b200 import fasta as fasta_mod
        parts = []
        for h in ('h1', 'h2'):
            d = dfs[h].iloc[plan_rank[h]].copy()
            d['_HAP'] = h
            parts.append(d)
        self.df = pd.concat(parts, ignore_index=True)
        self.tig_fa = tig_fa
        names_t, arrays = [], []
        for h in ('h1', 'h2'):
            fa = fasta_mod.open_fasta(tig_fa[h])
            for n in dict.fromkeys(self.df.loc[self.df['_HAP'] == h, 'QRY_ID']):
                names_t.append(n)
                arrays.append(fa.fetch_array(n))
        self.tig_store = device.SeqStore(ctx, names_t, arrays, keep_host=False)
        tidx = {n: i for i, n in enumerate(names_t)}
        ridx = {n: i for i, n in enumerate(ref_names)}
        rid = np.array([ridx[c] for c in self.df['#CHROM']], np.int32)
        qid = np.array([tidx[c] for c in self.df['QRY_ID']], np.int32)
        ops, op_off, perr = device.parse_cigars(self.df['CIGAR'].tolist())
        assert perr.code == 0
        self.batch = None
        if len(self.df):
            self.batch = device.CigarBatch(ctx, rid, qid, self.df['POS'].to_numpy(np.int32), self.df['REV'].to_numpy(np.uint8), ops, op_off)

    def run(self, ref_store):
        return self.batch.run(ref_store, self.tig_store) if self.batch is not None else None

    def close(self):
        if self.batch is not None:
            self.batch.close()
        self.tig_store.close()


def oracle_check_records(rh, tmp, ref_fa, tag, n_chk=2):
    """Rows of up to ``n_chk`` of this rank's records (on its smallest chromosome) against the oracle. True / False / None."""
    if rh.batch is None or len(rh.df) == 0:
        return None
    try:
        from oracle import pyoracle
        from pav_b200 import fasta as fasta_mod
        from pav_b200 import synth
        fa_r = fasta_mod.open_fasta(ref_fa)
        small = min(set(rh.df['#CHROM']), key=fa_r.length)
        pick = np.flatnonzero((rh.df['#CHROM'] == small).to_numpy())[:n_chk]
        sub = rh.df.iloc[pick]
        tig_seqs = {q: fasta_mod.open_fasta(rh.tig_fa[h]).fetch_array(q) for q, h in zip(sub['QRY_ID'], sub['_HAP'])}
        ref_p, tig_p, _ = synth.write_cigar_workload(os.path.join(tmp, f'chk_{tag}'), {small: fa_r.fetch_array(small)}, tig_seqs, sub)
        o_snv, o_indel, _ = pyoracle.walk_rows(sub, ref_p, tig_p)
        snv, indel, cerr = rh.batch.fetch()
        assert cerr.code == 0
        g_snv, g_indel = snv[np.isin(snv['rec'], pick)], indel[np.isin(indel['rec'], pick)]
        return bool(len(g_snv) == len(o_snv) and (g_snv['pos_ref'] == o_snv['pos_ref']).all() and (g_snv['qry_pos'] == o_snv['qry_pos']).all()
                    and len(g_indel) == len(o_indel)
                    and all((g_indel[c] == o_indel[c]).all() for c in ('pos', 'end', 'svlen', 'qry_pos', 'qry_end', 'left_shift', 'hom_ref_l', 'hom_ref_r',
                                                                         'hom_tig_l', 'hom_tig_r')))
    except Exception as ex:  # noqa: BLE001
        log(f'[{tag}] oracle spot check failed to run: {ex!r}')
        return None


def kernel_table(count_ms, walk_ms, hom_ms, hom_kernel, n_ops, n_chunks, n_snv, n_indel):
    """kernel -> (ms, algorithmic bytes per launch): DESIGN.md section 3."""
    return {
        'cigar_count_kernel': (count_ms, 4 * n_ops + 4 * n_chunks),     # per-record row counts + (its last CTA) the record scan
        'cigar_walk_kernel': (walk_ms, 4 * n_ops + 2 * 16 * n_chunks + 16 * n_snv + 64 * n_indel),
        HOM_NAMES[hom_kernel]: (hom_ms, (64 + 64 + 128) * n_indel),
    }


def gather_bound(n_indel, hom_ms):
    """Second, tighter bound of the homology kernel: its sequence reads are SCATTERED 32-byte sectors, and the memory system delivers
    those at a fraction of the streaming rate. The ceiling is measured by profiles/microbench/gather_rate.cu on a B200 of this pool
    (independent random 8-byte loads over 4 GiB, saturated: profiles/r02_gather_rate.jsonl); the floor of the kernel is four such
    accesses per indel (one per plane) plus its 128 B of streamed stub + row at the copy rate."""
    try:
        rows = [json.loads(ln) for ln in open(os.path.join(REPO, 'profiles', 'r02_gather_rate.jsonl')) if ln.strip()]
        peak = max(r['gsectors_per_s'] for r in rows if r.get('pattern') == 'random')
    except Exception:  # noqa: BLE001
        return None
    hbm, _ = measured_peak_gbs()
    floor_ms = 4 * n_indel / (peak * 1e9) * 1e3 + 128 * n_indel / (hbm * 1e9) * 1e3
    return {'peak_gaccesses_per_s': peak, 'source': 'profiles/r02_gather_rate.jsonl (profiles/microbench/gather_rate.cu)', 'accesses_per_launch': 4 * n_indel,
            'floor_ms': floor_ms, 'frac': floor_ms / hom_ms if hom_ms > 0 else None,
            'note': 'homology kernel against the measured rate of scattered sector reads (48.9 G/s = 1.57 TB/s as 32-byte sectors, a quarter of the copy rate)'}


BYTES_MODEL = ('4 B/op + 4 B/chunk (count); 4 B/op + 32 B/chunk descriptors + 16 B/SNV row + 64 B/indel stub (walk); 64 B stub + 64 B row + 128 B of '
               'sequence = one 32-byte DRAM sector from each of the four planes an indel touches (homology); DESIGN.md section 3')


def leg_c3(args, ctl, ctx, tmp, rank, world):
    """C3 sharded over the ranks: device-resident value + end-to-end call. Returns the fields of the top-level line."""
    from pav_b200 import device, multigpu
    from pav_b200 import fasta as fasta_mod
    from pav_b200.pavlib import cigarcall
    info = make_c3_files(args, tmp) if rank == 0 else None
    ctl.barrier()
    info = ctl.bcast(info)
    ref_fa = os.path.join(tmp, 'ref.fa')
    tig_fa = {h: os.path.join(tmp, f'{h}_tig.fa') for h in ('h1', 'h2')}
    dfs = {h: read_align(os.path.join(tmp, f'{h}_align.bed')) for h in ('h1', 'h2')}
    plan = shard_plan(dfs, world)

    # ---- reference planes: rank 0 packs, everyone else receives them over NCCL; every rank verifies what it holds
    fa_r = fasta_mod.open_fasta(ref_fa)
    ref_names = [str(n) for n in fa_r.index]
    t0 = time.perf_counter()
    if rank == 0:
        ref_store = device.SeqStore(ctx, ref_names, [fa_r.fetch_array(n) for n in ref_names], keep_host=False)
    else:
        ref_store = device.SeqStore.from_packed(ctx, ref_names, [fa_r.length(n) for n in ref_names], None, None)
    t_pack = time.perf_counter() - t0
    bcast_ms = 0.0
    if world > 1:
        uid = ctl.bcast(device.nccl_unique_id() if rank == 0 else None)
        ctl.barrier()
        with stdout_to_stderr():   # NCCL prints its version banner on stdout; the driver wants one JSON line there
            bcast_ms = ref_store.broadcast(uid, rank, world)
    sums = ctl.gather(ref_store.checksum())
    planes_ok = [s == sums[0] for s in sums]
    if not all(planes_ok):
        raise RuntimeError(f'reference planes differ between ranks after the broadcast: {sums}')
    bcast_ms = ctl.max(bcast_ms)
    _, b2, _, bm = ref_store.plane_sizes()

    # ---- this rank's records -> HBM
    haps = [ResidentShard(ctx, dfs, {h: plan[h][rank] for h in ('h1', 'h2')}, tig_fa, ref_names)]
    ctl.barrier()

    def step():
        ctx.l2_flush()
        return [s for s in (rh.run(ref_store) for rh in haps) if s is not None]

    clocks = ClockSampler(ctx.device)   # spans warm-up, the timed steps and identical untimed steps (the timed region is milliseconds)
    t_clk = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        step()
    ctl.barrier()
    step_ms, launches, graph_used = [], 0, True
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        sts = step()
        step_ms.append(sum(s.ms_kernels for s in sts))
        launches += sum(int(s.kernel_launches) for s in sts)
        graph_used = graph_used and all(int(s.graph) == 1 for s in sts)
    wall_ms = (time.perf_counter() - wall0) * 1e3 / max(args.steps, 1)
    while time.perf_counter() - t_clk < 1.2:
        step()
    clk = clocks.stop()
    clk['window'] = 'warm-up + timed steps + identical untimed steps, >= 1.2 s in total, nvidia-smi -lms 50'
    ctl.barrier()
    # per-kernel split: the same steps once more without the graph (events between the launches)
    os.environ['PAVGPU_NO_GRAPH'] = '1'
    parts = []
    for _ in range(6):
        sts = step()
        parts.append((sum(s.ms_count for s in sts), sum(s.ms_scan for s in sts), sum(s.ms_homology for s in sts), sum(s.ms_kernels for s in sts)))
    os.environ.pop('PAVGPU_NO_GRAPH', None)
    count_ms, walk_ms, hom_ms, nograph_ms = [float(np.mean([p[i] for p in parts[2:]])) for i in range(4)]
    sts = step()
    n_snv, n_indel = sum(int(s.n_snv) for s in sts), sum(int(s.n_indel) for s in sts)
    n_ops, n_chunks = sum(int(s.n_ops) for s in sts), sum(int(s.n_chunks) for s in sts)
    hom_kernel = max((int(s.homology_tiled) for s in sts), default=0)
    my_rows = n_snv + n_indel
    my_ms = float(np.mean(step_ms)) if step_ms else 0.0
    parity = [oracle_check_records(rh, tmp, ref_fa, f'r{rank}', n_chk=3) for rh in haps]
    per_rank = ctl.gather({'rank': rank, 'records': int(sum(len(rh.df) for rh in haps)), 'ops': n_ops, 'rows': my_rows, 'ms_per_step': my_ms,
                           'ms_count_scan': count_ms, 'ms_walk': walk_ms, 'ms_homology': hom_ms, 'ms_per_step_without_graph': nograph_ms,
                           'planes_checksum_equal_rank0': planes_ok[rank], 'oracle_spot_check': parity, 'graph': graph_used,
                           'sm_mhz': clk.get('sm_mhz'), 'clock_reasons': clk.get('reasons')})
    ms_per_step = max(p['ms_per_step'] for p in per_rank)
    total_rows = sum(p['rows'] for p in per_rank)
    value = total_rows / (ms_per_step * 1e-3) if ms_per_step > 0 else 0.0
    span = int(sum((rh.df['END'] - rh.df['POS']).sum() for rh in haps))
    for rh in haps:
        rh.close()
    checks = [x for p in per_rank for x in p['oracle_spot_check']]

    # ---- end to end: FASTA in -> DataFrames out.
    # N = 1: the public call per haplotype. N > 1: every rank makes the public call on the chromosomes it owns
    # (multigpu.make_insdel_snv_calls_shard: LPT over chromosomes, no communication -- each rank reads, uploads and packs 1 / N of the
    # reference and formats its own tables; the reference's own model of one job per batch of records, rules/align.snakefile:163, with
    # the chromosome as batch key so that the tables concatenate). The other multi-GPU call, make_insdel_snv_calls_dist (records over
    # ranks, NCCL broadcast of the packed reference, ONE merged table formatted on rank 0), is timed beside it as `e2e_dist`.
    e2e_s, e2e_rows, phases = [], 0, None
    E2E_WARMUP = 1
    os.environ.setdefault('PAVGPU_TUNE_ALLOC', '1')    # opt-in allocator tuning of the frame builder, declared in `config`
    for i in range((E2E_WARMUP + args.e2e_steps) if args.e2e_steps > 0 else 0):
        fasta_mod._CACHE.clear()   # every step re-opens and re-reads the FASTA files, like a fresh Snakemake job would
        out = None                 # the previous step's frames are released before the clock starts (the CPU arm's workers exit with theirs)
        ctl.barrier()
        t0 = time.perf_counter()
        rows_step, ph = 0, {}
        for h in ('h1', 'h2'):
            if world > 1:
                out = multigpu.make_insdel_snv_calls_shard(dfs[h], ref_fa, tig_fa[h], h, rank, world, version_id=False)
            else:
                out = cigarcall.make_insdel_snv_calls(dfs[h], ref_fa, tig_fa[h], h, version_id=False)
            ph[h] = cigarcall.last_phase_seconds
            rows_step += len(out[0]) + len(out[1])
            out = None             # (holding h1's frames while h2's are built was measured: no faster -- the allocator cannot reuse their memory)
        dt = ctl.max(time.perf_counter() - t0)
        rows_all = int(ctl.sum(rows_step))
        if i >= E2E_WARMUP:
            e2e_s.append(dt)
            e2e_rows, phases = rows_all, ph
        log(f'[rank {rank}] e2e step {i}: {dt:.2f}s ({rows_step} rows formatted on this rank) {ph}')
    if e2e_s:
        assert e2e_rows == total_rows, (e2e_rows, total_rows)
    dist_s, dist_ph = [], None
    if world > 1 and args.e2e_steps > 0:
        for i in range(E2E_WARMUP + args.e2e_steps):
            fasta_mod._CACHE.clear()
            out = None
            ctl.barrier()
            t0 = time.perf_counter()
            rows_d, ph = 0, {}
            for h in ('h1', 'h2'):
                with stdout_to_stderr():
                    out = multigpu.make_insdel_snv_calls_dist(dfs[h], ref_fa, tig_fa[h], h, version_id=False)
                ph[h] = dict(multigpu.last_dist_stats['seconds'], nccl_comm_reused=multigpu.last_dist_stats.get('nccl_comm_reused'))
                if out is not None:
                    rows_d += len(out[0]) + len(out[1])
                out = None
            dt = ctl.max(time.perf_counter() - t0)
            if rank == 0:
                assert rows_d == total_rows, (rows_d, total_rows)
            if i >= E2E_WARMUP:
                dist_s.append(dt)
                dist_ph = ph
            log(f'[rank {rank}] e2e_dist step {i}: {dt:.2f}s ({rows_d} rows formatted on this rank) {ph}')
    contig_bytes = int(sum(fasta_mod.open_fasta(tig_fa[h]).length(n) for h in ('h1', 'h2') for n in set(dfs[h]['QRY_ID'])))
    # per e2e step: each of the two haplotype calls uploads the reference (ASCII, packing rank) and every rank its contigs and ops
    h2d = int(2 * info['reference_bp'] + contig_bytes + 4 * ctl.sum(n_ops))
    d2h = int(16 * ctl.sum(n_snv) + 64 * ctl.sum(n_indel))
    e2e_val = total_rows / float(np.mean(e2e_s)) if e2e_s else None

    # ---- roofline of the dominant kernel on rank 0's shard (SURVEY 8(d) model beside this design's own)
    peak, peak_src = measured_peak_gbs()
    kernels = kernel_table(count_ms, walk_ms, hom_ms, hom_kernel, n_ops, n_chunks, n_snv, n_indel)
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bytes = kernels[dom]
    step_bytes = sum(b for _, b in kernels.values())
    survey_bytes = 4 * n_ops + -(-2 * span // 4) + -(-2 * span // 8) + 32 * n_snv + 64 * n_indel
    roofline = {
        'bound': 'hbm', 'kernel': dom, 'achieved': dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0, 'peak': peak, 'unit': 'GB/s',
        'frac': dom_bytes / (dom_ms * 1e-3) / 1e9 / peak if dom_ms > 0 else 0.0, 'traffic': None, 'peak_source': peak_src,
        'algorithmic_bytes_per_launch': int(dom_bytes), 'kernel_ms': dom_ms, 'per_kernel_ms': {k: v[0] for k, v in kernels.items()},
        'per_kernel_ms_source': 'rank 0, the timed steps repeated without the CUDA graph (events between the launches); the timed steps themselves '
                                'replay one graph (this rank\'s records of both haplotypes are one batch)',
        'step': {'algorithmic_bytes': int(step_bytes), 'achieved': step_bytes / (my_ms * 1e-3) / 1e9 if my_ms else None,
                 'frac': step_bytes / (my_ms * 1e-3) / 1e9 / peak if my_ms else None},
        'survey_8d_model': {'algorithmic_bytes': int(survey_bytes), 'achieved': survey_bytes / (my_ms * 1e-3) / 1e9 if my_ms else None,
                            'frac': survey_bytes / (my_ms * 1e-3) / 1e9 / peak if my_ms else None, 'per': 'whole step of rank 0 (count + walk + homology)'},
        'bytes_model': BYTES_MODEL, 'gather': gather_bound(n_indel, hom_ms),
        'note': 'traffic (DRAM bytes from ncu) is reported for the C2 leg (`c2.roofline`), the input the ncu captures under profiles/ are taken on',
    }
    ref_store.close()
    return {
        'value': value, 'ms_per_step': ms_per_step, 'per_rank': per_rank, 'rows': total_rows, 'gpu_launches': int(ctl.sum(launches)),
        'wall_ms_per_step_incl_flush': wall_ms, 'roofline': roofline, 'clocks': clk, 'ref_broadcast_ms': bcast_ms,
        'ref_broadcast_bytes': int(b2 + bm) if world > 1 else 0, 'ref_pack_seconds_rank0': ctl.bcast(t_pack),
        'planes_verified_on_every_rank': all(planes_ok), 'oracle_spot_check': bool(checks) and all(x is not False for x in checks) and any(x for x in checks),
        'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': args.e2e_steps, 'warmup': E2E_WARMUP,
                'ms_per_step': float(np.mean(e2e_s)) * 1e3 if e2e_s else None,
                'api': ('pav_b200.multigpu.make_insdel_snv_calls_shard per haplotype on every rank: the public call on the chromosomes the rank owns '
                        '(LPT over chromosomes; FASTA in, DataFrames out, no communication; tables concatenate with merge_shard_frames)' if world > 1 else
                        'pav_b200.pavlib.cigarcall.make_insdel_snv_calls per haplotype (FASTA in, DataFrames out)'),
                'phase_seconds_last_step_rank0': phases},
        'e2e_dist': ({'value': total_rows / float(np.mean(dist_s)), 'unit': UNIT, 'ms_per_step': float(np.mean(dist_s)) * 1e3,
                      'api': 'pav_b200.multigpu.make_insdel_snv_calls_dist per haplotype (records LPT-sharded over ranks, NCCL broadcast of the packed '
                             'reference, plane checksum on every rank, per-rank walk, host gather, ONE merged table formatted on rank 0)',
                      'phase_seconds_last_step_rank0': dist_ph}
                     if dist_s else None),
        'info': info,
    }


def leg_c5(args, ctl, ctx, rank, world):
    """C5 sweep split over the ranks: device-resident Gbases/s + the public call with lazy columns + oracle spot check."""
    from pav_b200 import _capi, device
    from pav_b200.pavlib import density
    n_total = args.c5_windows
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    t0 = time.perf_counter()
    wins = c5_windows_range(1005, lo, hi)
    t_gen = time.perf_counter() - t0
    CHUNK = 1024
    params = density.default_params()
    first_cols = None
    for rep in range(2):      # pass 0 warms the arena / pools, pass 1 is timed
        ms_tot = ms_kmer = ms_kde = 0.0
        rows = launches = n_eval = pairs = 0
        status = {}
        for a in range(0, len(wins), CHUNK):
            sub = wins[a:a + CHUNK]
            names = [f'w{i}' for i in range(len(sub))]
            rs = device.SeqStore(ctx, names, [w[0] for w in sub], keep_host=False)
            ts = device.SeqStore(ctx, names, [w[1] for w in sub], keep_host=False)
            win = np.zeros(len(sub), dtype=_capi.DENSITY_WINDOW)
            for i in range(len(sub)):
                win[i] = (i, i, 0, WIN_LEN, 0, WIN_LEN, 0, 20)
            batch = density.DensityBatch(ctx, win, params)
            ctx.l2_flush()
            st = batch.run(rs, ts)
            ms_tot += st.ms_kernels
            ms_kmer += st.ms_kmer
            ms_kde += st.ms_kde
            rows += int(st.rows)
            launches += int(st.kernel_launches)
            pairs += int(st.kde_pairs)
            res, runs, run_off = batch.fetch_runs()
            n_eval += int(res['n_eval'].sum())
            for s in res['status'].tolist():
                status[s] = status.get(s, 0) + 1
            if rep == 1 and a == 0 and len(sub):
                first_cols = batch.fetch_window(0, int(res['n_rows'][0])) if int(res['status'][0]) == 0 else {}
                first_cols['_status'] = int(res['status'][0])
            batch.close()
            rs.close()
            ts.close()
    # ---- the public call: ASCII windows in host memory -> run lengths of STATE in host memory (columns on demand)
    e2e_s, n_runs, e2e_phases = [], 0, {}
    for i in range(3):
        out = None
        ctl.barrier()
        t0 = time.perf_counter()
        out = density.density_windows([(w[0], w[1], False, 20) for w in wins], lazy=True)
        e2e_s.append(ctl.max(time.perf_counter() - t0))
        n_runs = sum(len(d['runs']) for d in out)
        e2e_phases = dict(density.last_stats.get('seconds_all_batches') or density.last_stats.get('seconds') or {})
    out = None
    e2e_t = float(np.mean(e2e_s[1:]))
    ok = None
    if wins and first_cols is not None:
        try:
            from oracle import pyoracle
            rc, o = pyoracle.density_arrays(wins[0][0].tobytes(), wins[0][1].tobytes())
            ok = bool(rc == first_cols['_status'] and (rc != 0 or all((first_cols[c].astype(np.int64) == o[c].astype(np.int64)).all()
                                                                   for c in ('KMER', 'INDEX', 'STATE_MER', 'STATE'))))
        except Exception as ex:  # noqa: BLE001
            log('density oracle spot check failed to run:', repr(ex))
    per_rank = ctl.gather({'rank': rank, 'windows': len(wins), 'ms': ms_tot, 'ms_kmer': ms_kmer, 'ms_kde': ms_kde, 'rows': rows, 'oracle_spot_check': ok,
                           'seconds_generate': t_gen})
    ms_max = max(p['ms'] for p in per_rank)
    total_bases = n_total * WIN_LEN
    W, K = WIN_LEN, 31
    kmer_bytes = len(wins) * (-(-2 * W // 4) + 8 * W + 16 * (W - K + 1) + 13 * (W - K + 1))
    peak, peak_src = measured_peak_gbs()
    roofs = {
        'kmer_part': {'bound': 'hbm', 'kernels': 'kmer_window (reference k-mer table in shared memory; ref_insert + tig_state for windows it does not take) + compact',
                      'bytes_model': 'per window: packed planes of both windows + 8 B per reference k-mer inserted + 2 x 8 B probed per contig k-mer + 13 B per row '
                                     'compacted -- the bytes of a table in HBM (r01 model, kept so that fractions compare across rounds); with the table in shared '
                                     'memory the 24 B per position never leave the SM (ncu: 9 MB of DRAM reads per 296 windows for the k-mer kernel)',
                      'algorithmic_bytes': int(kmer_bytes), 'ms': ms_kmer,
                      'achieved': kmer_bytes / (ms_kmer * 1e-3) / 1e9 if ms_kmer > 0 else None, 'peak': peak, 'unit': 'GB/s',
                      'frac': kmer_bytes / (ms_kmer * 1e-3) / 1e9 / peak if ms_kmer > 0 else None, 'peak_source': peak_src, 'per': f'rank {rank}'},
        'kde_part': {'bound': 'float64 CUDA-core arithmetic (exp + FMA), not HBM',
                     'kernels': 'runs_stats + kde_table + kde_eval x2 + gap_classify + finish_rows + state_rle',
                     'ms': ms_kde, 'rows_N': rows, 'evaluated_points_E': n_eval, 'eval_fraction_E_over_N': n_eval / rows if rows else None,
                     'logical_pairs': int(pairs), 'logical_pairs_per_sec': pairs / (ms_kde * 1e-3) if ms_kde > 0 else None},
    }
    checks = [p['oracle_spot_check'] for p in per_rank]
    return {
        'metric': METRIC_B, 'unit': UNIT_B, 'value': total_bases / (ms_max * 1e-3) / 1e9 if ms_max > 0 else None, 'ms_per_step': ms_max,
        'scaling': 'strong', 'config': {'workload': workload_c5(args, world), 'chunk_windows': CHUNK, 'l2': 'flushed before every chunk'},
        'e2e': {'value': total_bases / e2e_t / 1e9, 'unit': UNIT_B, 'ms_per_step': e2e_t * 1e3, 'phase_seconds_last_step_rank0': e2e_phases, 'h2d_bytes_per_step': int(2 * total_bases),
                'd2h_bytes_per_step': int(16 * ctl.sum(n_runs) + 32 * n_total),
                'api': 'pav_b200.pavlib.density.density_windows(lazy=True): ASCII windows in host memory -> status + run lengths of STATE in host memory; '
                       'the 38 B/row columns stay in HBM until a window becomes a call (pavgpu_density_batch_fetch_runs / _fetch_window)'},
        'per_rank': per_rank, 'roofline': roofs, 'rows': int(ctl.sum(rows)), 'gpu_launches': int(ctl.sum(launches)),
        'status_counts_this_rank': {str(k): v for k, v in status.items()},
        'oracle_spot_check': all(x is not False for x in checks) and any(x for x in checks),
    }


def leg_c2(args, ctx, tmp):
    """Round-1 line (N = 1): C2 device-resident walk, C-ABI host-buffer call, public call, per-kernel roofline."""
    from oracle import pyoracle
    from pav_b200 import device, synth
    from pav_b200 import fasta as fasta_mod
    from pav_b200.pavlib import cigarcall
    n_contigs, contig_len = args.c2_contigs, 200_000
    chrom_len = n_contigs * contig_len // 4
    ref, trs = synth.make_reference(1002, 4, chrom_len)
    tigs, df = synth.make_contigs(ref, trs, 1002, n_contigs, contig_len)
    ref_fa, tig_fa, _ = synth.write_cigar_workload(os.path.join(tmp, 'c2'), ref, tigs, df)
    names_r, names_t = list(ref), list(tigs)
    ref_store = device.SeqStore(ctx, names_r, [ref[n] for n in names_r])
    tig_arrays = [tigs[n] for n in names_t]
    tig_store = device.SeqStore(ctx, names_t, tig_arrays)
    rid = np.array([names_r.index(c) for c in df['#CHROM']], np.int32)
    tidx = {n: i for i, n in enumerate(names_t)}
    qid = np.array([tidx[c] for c in df['QRY_ID']], np.int32)
    pos, rev = df['POS'].to_numpy(np.int32), df['REV'].to_numpy(np.uint8)
    ops, op_off, perr = device.parse_cigars(df['CIGAR'].tolist())
    assert perr.code == 0
    batch = device.CigarBatch(ctx, rid, qid, pos, rev, ops, op_off)
    for _ in range(max(args.warmup, 3)):
        ctx.l2_flush()
        batch.run(ref_store, tig_store)
    step_ms = []
    for _ in range(args.steps):
        ctx.l2_flush()
        st = batch.run(ref_store, tig_store)
        step_ms.append(st.ms_kernels)
    graph = int(st.graph)
    os.environ['PAVGPU_NO_GRAPH'] = '1'
    parts = []
    for _ in range(7):
        ctx.l2_flush()
        s2 = batch.run(ref_store, tig_store)
        parts.append((s2.ms_count, s2.ms_scan, s2.ms_homology, s2.ms_kernels))
    os.environ.pop('PAVGPU_NO_GRAPH', None)
    count_ms, walk_ms, hom_ms, nograph_ms = [float(np.mean([p[i] for p in parts[2:]])) for i in range(4)]
    n_rows = int(st.n_snv + st.n_indel)
    my_ms = float(np.mean(step_ms))
    snv, indel, cerr = batch.fetch()
    parity = None
    try:
        n_chk = min(8, len(df))
        o_snv, o_indel, _ = pyoracle.walk_rows(df.iloc[:n_chk], ref_fa, tig_fa)
        g_snv, g_indel = snv[snv['rec'] < n_chk], indel[indel['rec'] < n_chk]
        parity = bool(len(g_snv) == len(o_snv) and (g_snv['pos_ref'] == o_snv['pos_ref']).all() and (g_snv['qry_pos'] == o_snv['qry_pos']).all()
                      and len(g_indel) == len(o_indel)
                      and all((g_indel[c] == o_indel[c]).all() for c in ('pos', 'end', 'svlen', 'qry_pos', 'qry_end', 'left_shift', 'hom_ref_l', 'hom_ref_r',
                                                                         'hom_tig_l', 'hom_tig_r')))
    except Exception as ex:  # noqa: BLE001
        log('C2 oracle spot check failed to run:', ex)
    cabi_s = []
    for i in range(1 + 3):
        t0 = time.perf_counter()
        ts2 = device.SeqStore(ctx, names_t, tig_arrays, keep_host=False)
        device.cigar_call(ctx, ref_store, ts2, rid, qid, pos, rev, ops, op_off)
        ts2.close()
        if i >= 1:
            cabi_s.append(time.perf_counter() - t0)
    h2d_cabi = int(sum(len(a) for a in tig_arrays) + ops.nbytes + op_off.nbytes + rid.nbytes * 3 + rev.nbytes)
    d2h_cabi = int(snv.nbytes + indel.nbytes)
    snv = indel = None
    batch.close()
    tig_store.close()
    e2e_s = []
    for i in range(2 + 3):
        fasta_mod._CACHE.clear()
        out = None
        t0 = time.perf_counter()
        out = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=False)
        if i >= 2:
            e2e_s.append(time.perf_counter() - t0)
    assert len(out[0]) + len(out[1]) == n_rows
    out = None
    ref_store.close()
    n_ops, n_snv, n_indel, n_chunks = int(st.n_ops), int(st.n_snv), int(st.n_indel), int(st.n_chunks)
    peak, peak_src = measured_peak_gbs()
    kernels = kernel_table(count_ms, walk_ms, hom_ms, int(st.homology_tiled), n_ops, n_chunks, n_snv, n_indel)
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bytes = kernels[dom]
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel on the same input, from this round's committed ncu capture
        tj = json.load(open(os.path.join(REPO, 'profiles', 'ncu_traffic.json')))
        if tj.get('n_ops') == n_ops and dom in tj.get('kernels', {}):
            traffic = tj['kernels'][dom]
    except Exception:  # noqa: BLE001
        pass
    step_bytes = sum(b for _, b in kernels.values())
    span = int((df['END'] - df['POS']).sum())
    survey_bytes = 4 * n_ops + -(-2 * span // 4) + -(-2 * span // 8) + 32 * n_snv + 64 * n_indel
    return {
        'config': {'workload': WORKLOAD_C2, 'ops': n_ops, 'rows': n_rows, 'snv_rows': n_snv, 'indel_rows': n_indel},
        'metric': METRIC, 'unit': UNIT, 'value': n_rows / (my_ms * 1e-3), 'ms_per_step': my_ms, 'ms_per_step_without_graph': nograph_ms, 'graph': graph,
        'oracle_spot_check': parity,
        'e2e': {'value': n_rows / float(np.mean(e2e_s)), 'unit': UNIT, 'ms_per_step': float(np.mean(e2e_s)) * 1e3,
                'h2d_bytes_per_step': int(sum(len(ref[n]) for n in names_r) + h2d_cabi), 'd2h_bytes_per_step': d2h_cabi,
                'api': 'pav_b200.pavlib.cigarcall.make_insdel_snv_calls (FASTA in, DataFrames out)', 'phase_seconds_last_step': cigarcall.last_phase_seconds},
        'e2e_cabi': {'value': n_rows / float(np.mean(cabi_s)), 'unit': UNIT, 'ms_per_step': float(np.mean(cabi_s)) * 1e3, 'h2d_bytes_per_step': h2d_cabi,
                     'd2h_bytes_per_step': d2h_cabi, 'api': 'pavgpu_seqstore_create(contigs) + pavgpu_cigar_call (host buffers)'},
        'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms else 0.0, 'peak': peak, 'unit': 'GB/s',
                     'frac': dom_bytes / (dom_ms * 1e-3) / 1e9 / peak if dom_ms else 0.0, 'traffic': traffic, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': int(dom_bytes), 'kernel_ms': dom_ms, 'per_kernel_ms': {k: v[0] for k, v in kernels.items()},
                     'step': {'algorithmic_bytes': int(step_bytes), 'achieved': step_bytes / (my_ms * 1e-3) / 1e9, 'frac': step_bytes / (my_ms * 1e-3) / 1e9 / peak},
                     'survey_8d_model': {'algorithmic_bytes': int(survey_bytes), 'achieved': survey_bytes / (my_ms * 1e-3) / 1e9,
                                         'frac': survey_bytes / (my_ms * 1e-3) / 1e9 / peak},
                     'bytes_model': BYTES_MODEL, 'gather': gather_bound(n_indel, hom_ms)},
    }


def cpu_baseline_ours(args, tmp, info):
    """N = 1 only: the unmodified reference on a bounded sample, both paths (the code --impl reference runs for K steps)."""
    from oracle import refenv
    if not refenv.available():
        return None, None
    cores = len(os.sched_getaffinity(0))
    pick = cpu_sample(args, cores, info['records'])
    tasks = []
    for h in ('h1', 'h2'):
        bed = os.path.join(tmp, f'{h}_align.bed')
        tasks += [(bed, [i], os.path.join(tmp, 'ref.fa'), os.path.join(tmp, f'{h}_tig.fa'), h) for i in pick[h]]
    reference_walk_step(tasks, cores)        # untimed: starts the worker pool and imports the reference in every worker
    rows, dt = reference_walk_step(tasks, cores)
    a = {'value': rows / dt, 'unit': UNIT, 'cores': min(cores, len(tasks)), 'kind': 'reference',
         'sample': f'{len(tasks)} of {sum(info["records"].values())} C3 records (seeded sample, one per worker process; {rows} variant rows, {dt:.1f}s, second of two passes): the unmodified '
                   'reference\'s make_insdel_snv_calls (oracle/_ref/pavlib/cigarcall.py), FASTA read by .fai offset'}
    try:      # the port figure beside it (oracle C walk + reference-style frame assembly, same records, same pool)
        p_rows, p_dt = port_walk_step(tasks, cores)
        a['port'] = {'value': p_rows / p_dt, 'unit': UNIT, 'kind': 'port', 'seconds': p_dt}
    except Exception as ex:  # noqa: BLE001
        log('port walk leg failed:', repr(ex))
    b = None
    if args.c5_windows > 0:
        n_w = min(args.cpu_density_windows or cores, args.c5_windows)
        gb, dt, startup, rcs = reference_density(c5_windows_range(1005, 0, n_w), tmp, cores)
        b = {'value': gb, 'unit': UNIT_B, 'cores': cores, 'kind': 'reference',
             'sample': f'first {n_w} of {args.c5_windows} C5 windows, one `python3 scripts/density.py ... -t 1` process per window (pavlib/inv.py:249-266), '
                       f'{cores} at a time, {dt:.1f}s; start-up ({startup:.2f}s per process, measured separately) is inside the time',
             'startup_seconds_per_process': startup}
        try:
            pgb, pdt = port_density(c5_windows_range(1005, 0, n_w), cores)
            b['port'] = {'value': pgb, 'unit': UNIT_B, 'kind': 'port', 'seconds': pdt}
        except Exception as ex:  # noqa: BLE001
            log('port density leg failed:', repr(ex))
    return a, b


def run_ours(args, rank, world, local):
    from pav_b200 import build as pbuild
    if rank == 0:
        pbuild.build()
    ctl = Control(rank, world)
    ctl.barrier()
    from pav_b200 import device
    os.environ['PAVGPU_DEVICE_INDEX'] = str(local)
    ctx = device.get_context(local if device._capi.lib().pavgpu_device_count() > local else 0)
    tmp = ctl.bcast(tempfile.mkdtemp(prefix='pavbench_') if rank == 0 else None)

    a = leg_c3(args, ctl, ctx, tmp, rank, world)
    b = None
    if args.c5_windows > 0:
        try:
            b = leg_c5(args, ctl, ctx, rank, world)
        except Exception as ex:  # noqa: BLE001
            if world > 1:
                raise            # ranks must stay in step
            log('C5 leg failed:', repr(ex))
            b = {'error': repr(ex)}
    c2 = None
    if world == 1 and args.c2:
        try:
            c2 = leg_c2(args, ctx, tmp)
        except Exception as ex:  # noqa: BLE001
            log('C2 leg failed:', repr(ex))
            c2 = {'error': repr(ex)}
    cpu_a = cpu_b = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_a, cpu_b = cpu_baseline_ours(args, tmp, a['info'])
    if rank == 0:
        info = a.pop('info')
        line = {
            'metric': METRIC, 'value': a['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': a['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic',
            'config': {'workload': workload_c3(args, world), 'reference_bp': info['reference_bp'], 'records': info['records'], 'rows': a['rows'],
                       'l2': 'flushed (256 MB memset) between iterations', 'parallelism': f'alignment records sharded over {world} GPU(s), LPT',
                       'alloc_tuning': 'PAVGPU_TUNE_ALLOC=1 for the end-to-end leg (opt-in glibc / pymalloc settings of the frame builder)',
                       'seconds_generate': info['seconds_generate'], 'seconds_write': info['seconds_write']},
            'e2e': a['e2e'], 'e2e_dist': a['e2e_dist'], 'gpu_launches': a['gpu_launches'], 'wall_ms_per_step_incl_flush': a['wall_ms_per_step_incl_flush'],
            'roofline': a['roofline'], 'cpu_baseline': cpu_a, 'clocks': a['clocks'], 'per_rank': a['per_rank'],
            'ref_broadcast_ms': a['ref_broadcast_ms'], 'ref_broadcast_bytes': a['ref_broadcast_bytes'], 'ref_pack_seconds_rank0': a['ref_pack_seconds_rank0'],
            'planes_verified_on_every_rank': a['planes_verified_on_every_rank'], 'oracle_spot_check': a['oracle_spot_check'],
            'secondary': b, 'c2': c2,
        }
        if b is not None and 'error' not in b:
            b['cpu_baseline'] = cpu_b
        if args.metric == 'density' and b is not None and 'error' not in b:
            top = dict(b)
            rf = b['roofline']['kmer_part']
            top.update({'n_gpus': world, 'steps': 1, 'warmup': 1, 'higher_is_better': True, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                        'roofline': {'bound': 'hbm', 'kernel': rf['kernels'], 'achieved': rf['achieved'], 'peak': rf['peak'], 'unit': 'GB/s', 'frac': rf['frac'],
                                     'traffic': None, 'kde_part': b['roofline']['kde_part']},
                        'clocks': a['clocks'], 'secondary': {k: v for k, v in line.items() if k not in ('secondary', 'c2')}})
            line = top
        print(json.dumps(line), flush=True)
    ctl.barrier()
    ctl.close()


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local)


if __name__ == '__main__':
    main()
