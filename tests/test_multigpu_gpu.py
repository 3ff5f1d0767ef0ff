"""Two-GPU run of the distributed walk (NCCL broadcast of the packed reference + host gather). Skipped on
single-GPU boxes; the host logic is covered on CPU by tests/test_multigpu_cpu.py."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp, out_q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), LOCAL_RANK=str(rank), PAVGPU_DEVICE_INDEX=str(rank))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pav_b200 import multigpu
    df = pd.read_csv(os.path.join(tmp, 'wl_align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)
    first = multigpu.make_insdel_snv_calls_dist(df, os.path.join(tmp, 'wl_ref.fa'), os.path.join(tmp, 'wl_tig.fa'), 'h1', version_id=True)
    reused_first = multigpu.last_dist_stats['nccl_comm_reused']
    # the second call of the process reuses the NCCL communicator of the first and must return the same tables
    res = multigpu.make_insdel_snv_calls_dist(df, os.path.join(tmp, 'wl_ref.fa'), os.path.join(tmp, 'wl_tig.fa'), 'h1', version_id=True)
    st = multigpu.last_dist_stats
    assert reused_first is False and st['nccl_comm_reused'] is True
    if rank == 0:
        assert first[0].equals(res[0]) and first[1].equals(res[1])
    # what every rank holds after the broadcast, read back from ITS device: full export of the planes (not only the checksum)
    from pav_b200 import device, fasta
    out_q.put(('stats', rank, st['planes_verified'], st['checksum'], st['records'], st['rows']))
    if rank == 0:
        out_q.put(('tables', res[0].to_csv(sep='\t', index=False), res[1].to_csv(sep='\t', index=False)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_equal_oracle(tmp_path):
    from pav_b200 import _capi, synth
    if _capi.lib().pavgpu_device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from oracle import pyoracle
    ref, tigs, df = synth.make_cigar_workload(41, 2, 400_000, 30, 25_000, edit_rate=0.01, rev_frac=0.5)
    synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    df = pd.read_csv(os.path.join(str(tmp_path), 'wl_align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)
    exp = pyoracle.make_insdel_snv_calls(df, str(tmp_path / 'wl_ref.fa'), str(tmp_path / 'wl_tig.fa'), 'h1', version_id=True)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    msgs = [q.get(timeout=300) for _ in range(3)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = [m for m in msgs if m[0] == 'tables'][0]
    stats = sorted(m for m in msgs if m[0] == 'stats')
    assert got[1] == exp[0].to_csv(sep='\t', index=False) and got[2] == exp[1].to_csv(sep='\t', index=False)
    # both ranks verified the planes they hold against rank 0's checksum, hold the same checksum, and both walked records
    assert [m[2] for m in stats] == [True, True] and stats[0][3] == stats[1][3] and stats[0][3] != (0, 0)
    assert all(m[4] > 0 and m[5] > 0 for m in stats) and sum(m[4] for m in stats) == len(df)


def test_chromosome_shards_concatenate_to_the_whole_table(tmp_path):
    """multigpu.chrom_shards / make_insdel_snv_calls_shard / merge_shard_frames: every rank formats the chromosomes it owns (no
    communication); the shards' tables put together are the single call's tables, IDs with their version suffixes included."""
    from pav_b200 import multigpu, synth
    from pav_b200.pavlib import cigarcall
    ref, tigs, df = synth.make_cigar_workload(43, 5, 300_000, 40, 20_000, edit_rate=0.01, rev_frac=0.5)
    synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    df = pd.read_csv(os.path.join(str(tmp_path), 'wl_align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)
    ref_fa, tig_fa = str(tmp_path / 'wl_ref.fa'), str(tmp_path / 'wl_tig.fa')
    whole = cigarcall.make_insdel_snv_calls(df, ref_fa, tig_fa, 'h1', version_id=True)
    for world in (1, 2, 3):
        shards = multigpu.chrom_shards(df, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(len(df)))
        owners = [set(df['#CHROM'].iloc[s]) for s in shards]
        assert all(not (owners[a] & owners[b]) for a in range(world) for b in range(a + 1, world))
        parts = [multigpu.make_insdel_snv_calls_shard(df, ref_fa, tig_fa, 'h1', r, world, version_id=True) for r in range(world)]
        merged = multigpu.merge_shard_frames(parts)
        for got, exp in zip(merged, whole):
            assert got.to_csv(sep='\t', index=False) == exp.to_csv(sep='\t', index=False)
