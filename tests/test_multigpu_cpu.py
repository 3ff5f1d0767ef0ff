"""World-size-2 gloo test (CPU) of the multi-GPU host logic: LPT sharding, gather, merge back into the
reference's emission order, DataFrame assembly on rank 0. The device walk is replaced by the CPU oracle
(test infrastructure) through ``walk_fn`` so the test runs without a GPU; the GPU variant lives in
tests/test_multigpu_gpu.py."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_walk(table, ref_arr, tig_arr):
    """Rows in the product's layout, computed by the CPU oracle (checker only)."""
    from oracle import pyoracle
    from pav_b200 import _capi
    L = pyoracle.lib()
    w = L.orc_walk_new()
    import ctypes
    try:
        for rec in range(table.n_rec):
            ref = ref_arr[table.ref_id[rec]].tobytes()
            tig = tig_arr[table.qry_id[rec]].tobytes()
            if table.rev[rec]:
                tig = pyoracle.revcomp_bytes(tig)
            cg = table.cigars[rec].encode()
            rc = L.orc_walk_record(w, cg, len(cg), int(table.pos[rec]), ref, ref.upper(), len(ref), tig, tig.upper(), len(tig),
                                   int(table.rev[rec]), rec)
            assert rc == 0
        ns, ni = L.orc_walk_n_snv(w), L.orc_walk_n_indel(w)
        o_snv = np.zeros(ns, pyoracle.SNV_DTYPE)
        o_ind = np.zeros(ni, pyoracle.INDEL_DTYPE)
        if ns:
            ctypes.memmove(o_snv.ctypes.data, L.orc_walk_snv(w), ns * pyoracle.SNV_DTYPE.itemsize)
        if ni:
            ctypes.memmove(o_ind.ctypes.data, L.orc_walk_indel(w), ni * pyoracle.INDEL_DTYPE.itemsize)
    finally:
        L.orc_walk_free(w)
    snv = np.zeros(ns, _capi.SNV_ROW)
    snv['pos_ref'], snv['qry_pos'], snv['rec'] = o_snv['pos_ref'], o_snv['qry_pos'], o_snv['rec']
    # op_idx: rows of one record are already in op order; a running counter per record is a valid key
    snv['op_idx'] = np.arange(ns)
    indel = np.zeros(ni, _capi.INDEL_ROW)
    for c in ('rec', 'svtype', 'svlen', 'pos', 'end', 'qry_pos', 'qry_end', 'left_shift', 'hom_ref_l', 'hom_ref_r', 'hom_tig_l', 'hom_tig_r'):
        indel[c] = o_ind[c]
    indel['op_idx'] = np.arange(ni)
    return snv, indel


def _worker(rank, world, port, tmp, out_q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pav_b200 import multigpu
    df = pd.read_csv(os.path.join(tmp, 'wl_align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)
    res = multigpu.make_insdel_snv_calls_dist(df, os.path.join(tmp, 'wl_ref.fa'), os.path.join(tmp, 'wl_tig.fa'), 'h1',
                                              version_id=True, walk_fn=_oracle_walk)
    if rank == 0:
        out_q.put((res[0].to_csv(sep='\t', index=False), res[1].to_csv(sep='\t', index=False),
                   [int(i) for i in res[0].index], [int(i) for i in res[1].index]))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_shards_balanced_and_complete():
    from pav_b200 import multigpu
    rng = np.random.default_rng(0)
    costs = rng.pareto(1.5, 500) + 1
    shards = multigpu.lpt_shards(costs, 8)
    allidx = np.sort(np.concatenate(shards))
    assert (allidx == np.arange(500)).all()
    loads = np.array([costs[s].sum() for s in shards])
    assert loads.max() <= loads.mean() + costs.max()
    assert all((np.diff(s) > 0).all() for s in shards if len(s) > 1)


def test_dist_two_ranks_equal_single_process(tmp_path):
    from oracle import pyoracle
    from pav_b200 import synth
    ref, tigs, df = synth.make_cigar_workload(31, 2, 150_000, 13, 20_000, edit_rate=0.01, rev_frac=0.5, clip=(3, 4))
    synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    df = pd.read_csv(os.path.join(str(tmp_path), 'wl_align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)
    exp_snv, exp_indel = pyoracle.make_insdel_snv_calls(df, str(tmp_path / 'wl_ref.fa'), str(tmp_path / 'wl_tig.fa'), 'h1', version_id=True)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == exp_snv.to_csv(sep='\t', index=False)
    assert got[1] == exp_indel.to_csv(sep='\t', index=False)
    assert got[2] == [int(i) for i in exp_snv.index] and got[3] == [int(i) for i in exp_indel.index]


def test_chrom_shards_partition_by_reference_sequence():
    """multigpu.chrom_shards: every record in exactly one shard, a chromosome never in two, loads balanced by LPT over chromosomes."""
    import pandas as pd
    from pav_b200 import multigpu
    rng = np.random.default_rng(5)
    chroms = [f'chr{i}' for i in range(1, 25)]
    weight = np.linspace(8.0, 1.5, 24)
    rows = []
    for c, wgt in zip(chroms, weight):
        for _ in range(int(wgt * 4)):
            n_ops = int(rng.integers(200, 400))
            rows.append((c, 0, 1000 * n_ops, '10=' * n_ops))
    order = rng.permutation(len(rows))
    df = pd.DataFrame([rows[i] for i in order], columns=['#CHROM', 'POS', 'END', 'CIGAR'])
    for world in (1, 2, 4, 8):
        shards = multigpu.chrom_shards(df, world)
        assert len(shards) == world and sorted(np.concatenate(shards).tolist()) == list(range(len(df)))
        owners = [set(df['#CHROM'].iloc[s]) for s in shards]
        assert sum(len(o) for o in owners) == 24
        cost = multigpu.record_costs(df['CIGAR'].tolist(), (df['END'] - df['POS']).to_numpy())
        load = np.array([cost[s].sum() for s in shards])
        assert load.max() <= 1.15 * load.mean(), (world, load)
    # more ranks than chromosomes: the surplus ranks get empty shards, nothing is lost
    small = df.loc[df['#CHROM'].isin(['chr1', 'chr2', 'chr3'])].reset_index(drop=True)
    shards = multigpu.chrom_shards(small, 8)
    assert sum(1 for s_ in shards if len(s_)) == 3 and sorted(np.concatenate(shards).tolist()) == list(range(len(small)))
    import pandas as pd
    empty = (pd.DataFrame({'#CHROM': []}), pd.DataFrame({'#CHROM': []}))
    part = (pd.DataFrame({'#CHROM': ['b', 'a', 'a'], 'POS': [5, 1, 9]}), pd.DataFrame({'#CHROM': ['a'], 'POS': [3]}))
    snv, ins = multigpu.merge_shard_frames([empty, part, empty])
    assert snv['#CHROM'].tolist() == ['a', 'a', 'b'] and snv['POS'].tolist() == [1, 9, 5] and ins['POS'].tolist() == [3]
    assert multigpu.merge_shard_frames([empty, empty])[0].shape[0] == 0
