"""Multi-GPU host layer: alignment records shard across ranks, the packed reference is
broadcast once from rank 0 over NCCL, variant rows are gathered on the host and formatted on rank 0.

One process per GPU. The control plane is any ``torch.distributed`` process group (gloo or nccl); the
only data-path collective is ``pavgpu_seqstore_broadcast``. Row order is independent of the number of
ranks: rows carry (record, op) keys and are merged back into the reference's emission order before the
stable sort.
"""
import numpy as np

from . import device, fasta
from .pavlib import cigarcall


def record_costs(cigars, ref_span=None):
    """Cheap per-record work estimate for partitioning: CIGAR text length (~ #ops) + aligned span / 64."""
    c = np.array([len(x) for x in cigars], dtype=np.float64) / 3.0
    if ref_span is not None:
        c = c + np.asarray(ref_span, dtype=np.float64) / 64.0
    return c


def lpt_shards(costs, n_shards):
    """Longest-processing-time-first partition. Returns ``n_shards`` sorted index arrays."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind='stable')
    load = np.zeros(n_shards)
    bins = [[] for _ in range(n_shards)]
    for i in order.tolist():
        k = int(np.argmin(load))
        bins[k].append(i)
        load[k] += costs[i]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def merge_rows(parts, shard_index):
    """Per-rank ``(snv, indel)`` row arrays with shard-local ``rec`` -> one pair in global emission order
    (record, op, base), ``rec`` rewritten to the row number in the full table."""
    snvs, indels = [], []
    for (snv, indel), idx in zip(parts, shard_index):
        snv, indel = snv.copy(), indel.copy()
        if len(snv):
            snv['rec'] = idx[snv['rec']]
        if len(indel):
            indel['rec'] = idx[indel['rec']]
        snvs.append(snv)
        indels.append(indel)
    snv = np.concatenate(snvs) if snvs else np.zeros(0, cigarcall.device._capi.SNV_ROW)
    indel = np.concatenate(indels) if indels else np.zeros(0, cigarcall.device._capi.INDEL_ROW)
    # inside a shard the rows of a record are contiguous and already in (op, base) order, and the shards' records are ascending:
    # a stable sort on the record number alone restores the global emission order (a few sorted runs: cheap for the merge sort)
    snv = snv[np.argsort(snv['rec'], kind='stable')]
    indel = indel[np.argsort(indel['rec'], kind='stable')]
    return snv, indel


def chrom_shards(df_align, n_shards):
    """Partition the records by REFERENCE SEQUENCE: longest-processing-time-first over the chromosomes (cost = the summed
    ``record_costs`` of a chromosome's records). Returns ``n_shards`` sorted index arrays into ``df_align``.

    A chromosome lives in exactly one shard, so a shard is a self-contained ``make_insdel_snv_calls`` input: it reads, uploads and
    packs only its own chromosomes (1 / n of the reference per rank, no broadcast, no gather), variant IDs and their ``version_id``
    suffixes never cross shards, and the per-shard tables concatenate into the whole table (``merge_shard_frames``). This is the
    split to use when every rank formats its own rows -- the reference's model of one job per batch of records
    (CALL_BATCH, rules/align.snakefile:163; tables merged by rule call_cigar_merge) with the batch key changed from INDEX % 10 to
    the chromosome. ``make_insdel_snv_calls_dist`` (records over ranks, one merged table on rank 0) is the other one.
    With fewer chromosomes than ranks the surplus ranks get empty shards (and return empty tables): a human reference keeps 8 GPUs
    busy (24 chromosomes, chr1 is 8 % of the genome), a 4-chromosome input does not -- shard that one by record."""
    chrom = df_align['#CHROM'].to_numpy(dtype=object)
    span = (df_align['END'].to_numpy(dtype=np.int64) - df_align['POS'].to_numpy(dtype=np.int64)) if 'END' in df_align.columns else None
    cost = record_costs(df_align['CIGAR'].tolist(), span)
    names, inv = np.unique(chrom.astype(str), return_inverse=True)
    per_chrom = np.bincount(inv, weights=cost, minlength=len(names))
    bins = lpt_shards(per_chrom, n_shards)
    owner = np.empty(len(names), dtype=np.int64)
    for r, b in enumerate(bins):
        owner[b] = r
    rec_owner = owner[inv]
    return [np.flatnonzero(rec_owner == r) for r in range(n_shards)]


def make_insdel_snv_calls_shard(df_align, ref_fa_name, tig_fa_name, hap, rank, world, version_id=True):
    """This rank's part of ``make_insdel_snv_calls`` under ``chrom_shards``: the two tables of the chromosomes it owns (columns and
    row order as in the whole table restricted to those chromosomes; row labels count the shard's own rows). No communication."""
    idx = chrom_shards(df_align, world)[rank]
    return cigarcall.make_insdel_snv_calls(df_align.iloc[idx], ref_fa_name, tig_fa_name, hap, version_id=version_id)


def merge_shard_frames(frames):
    """``[(df_snv, df_insdel), ...]`` of all shards -> the whole pair of tables: the shards' rows in chromosome order (each shard is
    already sorted inside its chromosomes, and no chromosome is in two shards), row labels renumbered."""
    import pandas as pd
    out = []
    for k in (0, 1):
        parts = [f[k] for f in frames if f is not None and len(f[k])]
        if not parts:
            out.append(frames[0][k])
            continue
        df = pd.concat(parts, axis=0)
        order = np.argsort(df['#CHROM'].to_numpy(dtype=object).astype(str), kind='stable')
        out.append(df.iloc[order].reset_index(drop=True))
    return tuple(out)


last_dist_stats = None   # per-call record of the last make_insdel_snv_calls_dist on this rank (timings, checksums)


def make_insdel_snv_calls_dist(df_align, ref_fa_name, tig_fa_name, hap, version_id=True, group=None, walk_fn=None, verify_planes=True):
    """Distributed ``make_insdel_snv_calls``: call on every rank with the same arguments; rank 0 returns
    ``(df_snv, df_insdel)`` identical to the single-GPU result, the other ranks return ``None``.

    Records are split over the ranks by longest-processing-time-first (the reference's analogue is the file-level split
    ``CALL_BATCH = INDEX % 10``, rules/align.snakefile:163 / pavlib/cigarcall.py:21); rank 0 packs the reference and broadcasts the
    packed planes with one NCCL broadcast; every rank then checks the planes in its HBM against rank 0's checksum
    (``verify_planes``), walks its shard, and the rows are gathered on the host of rank 0 and formatted there.

    ``walk_fn(table, ref_arr, tig_arr) -> (snv, indel)`` replaces the device walk (CPU tests of the host
    logic inject a checker here).

    Failures before or during the collectives (no device, out of memory while packing, NCCL) are agreed on by all ranks first, so
    every rank raises instead of some of them waiting in a broadcast forever; only errors of a rank's own walk (CIGAR errors) are
    captured per rank and re-raised on rank 0 in table order.
    """
    import time

    import torch.distributed as dist
    global last_dist_stats
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_rec = df_align.shape[0]
    if n_rec == 0:
        return (cigarcall._empty(cigarcall.SNV_COLUMNS), cigarcall._empty(cigarcall.INSDEL_COLUMNS)) if rank == 0 else None
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    full = cigarcall.AlignTable(df_align)
    span = None
    if 'END' in df_align.columns:
        span = df_align['END'].to_numpy(dtype=np.int64) - full.pos
    shards = lpt_shards(record_costs(full.cigars, span), world)
    mine = df_align.iloc[shards[rank]]
    marks = [('plan', time.perf_counter())]
    ref_fa, tig_fa = fasta.open_fasta(ref_fa_name), fasta.open_fasta(tig_fa_name)
    part = (np.zeros(0, device._capi.SNV_ROW), np.zeros(0, device._capi.INDEL_ROW))
    stats = {'rank': rank, 'world': world, 'records': int(len(mine)), 'bcast_ms': 0.0, 'checksum': None, 'planes_verified': None}
    err = None
    ref_store = ctx = None
    ref_arr = tig_all = None
    pool = ThreadPoolExecutor(max_workers=2 * cigarcall._READERS + 1)
    try:
        t_mine = cigarcall.AlignTable(mine) if len(mine) else None
        if walk_fn is None:
            # ---- the reference planes: every rank takes part in the collectives or all of them raise together
            setup_err, uid, have_comm = None, [None], False
            tig_futs = []
            try:
                ctx = device.get_context()
                names = list(full.ref_names)  # every rank holds the whole reference of this table, same order
                # sequences are read by a few threads each (cigarcall.read_sequences): rank 0 the reference and, for the frames it
                # builds at the end, every contig of the table; the other ranks the contigs of their records
                pinned = cigarcall._CALLS >= 1
                cigarcall._CALLS += 1
                if rank == 0:
                    ref_arr, ref_futs = cigarcall.read_sequences(ref_fa, names, pool, ctx, pinned)
                    tig_all, tig_futs = cigarcall.read_sequences(tig_fa, list(full.tig_names), pool, ctx, pinned)
                    for f in ref_futs:
                        f.result()
                    marks.append(('read_reference', time.perf_counter()))
                    ref_store = device.SeqStore(ctx, names, ref_arr, keep_host=False)
                else:
                    if t_mine is not None:
                        tig_all, tig_futs = cigarcall.read_sequences(tig_fa, list(t_mine.tig_names), pool, ctx, pinned)
                    ref_store = device.SeqStore.from_packed(ctx, names, [ref_fa.length(n) for n in names], None, None)
                have_comm = device.nccl_comm_cached(ctx, rank, world)
                marks.append(('store', time.perf_counter()))
            except Exception as ex:  # noqa: BLE001
                setup_err = f'rank {rank}: {type(ex).__name__}: {ex}'
                have_comm = False
            flags = [None] * world
            dist.all_gather_object(flags, (setup_err, have_comm), group=group)
            if any(e for e, _ in flags):
                if ref_store is not None:
                    ref_store.close()
                raise RuntimeError('make_insdel_snv_calls_dist: reference set-up failed: ' + '; '.join(e for e, _ in flags if e))
            # the NCCL communicator of an earlier call is reused when every rank still has it (setting one up costs seconds);
            # otherwise rank 0 draws a new id and all ranks join it
            if all(h for _, h in flags):
                uid = [None]
            else:
                if rank == 0:
                    uid = [device.nccl_unique_id()]
                dist.broadcast_object_list(uid, src=0, group=group)
            stats['nccl_comm_reused'] = uid[0] is None
            marks.append(('agree', time.perf_counter()))
            bc_err = None
            try:
                stats['bcast_ms'] = ref_store.broadcast(uid[0], rank, world)
                marks.append(('broadcast', time.perf_counter()))
                stats['checksum'] = ref_store.checksum() if verify_planes else None
                marks.append(('checksum', time.perf_counter()))
            except Exception as ex:  # noqa: BLE001
                bc_err = f'rank {rank}: {type(ex).__name__}: {ex}'
            sums = [None] * world
            dist.all_gather_object(sums, (bc_err, stats['checksum']), group=group)
            bad = [e for e, _ in sums if e]
            if not bad and verify_planes:
                bad = [f'rank {r}: planes {c} differ from rank 0 {sums[0][1]}' for r, (_, c) in enumerate(sums) if c != sums[0][1]]
                stats['planes_verified'] = not bad
            if bad:
                ref_store.close()
                raise RuntimeError('make_insdel_snv_calls_dist: reference broadcast failed: ' + '; '.join(bad))
        t1 = time.perf_counter()
        # ---- this rank's shard
        try:
            if t_mine is not None:
                t = t_mine
                if walk_fn is not None:
                    part = walk_fn(t, [ref_fa.fetch_array(n) for n in t.ref_names], [tig_fa.fetch_array(n) for n in t.tig_names])
                else:
                    for f in tig_futs:
                        f.result()
                    # ids must index the broadcast store, not the shard-local name table
                    t.ref_id = np.array([full.ref_names[str(c)] for c in t.chrom.tolist()], dtype=np.int32)
                    tig_arr_mine = [tig_all[full.tig_names[n]] for n in t.tig_names] if rank == 0 else tig_all
                    part = cigarcall.walk_rows(t, None, tig_arr_mine, ctx=ctx, ref_store=ref_store)
                    stats['walk'] = dict(cigarcall.last_stats) if cigarcall.last_stats else None
        except (RuntimeError, IndexError) as ex:  # CIGAR errors: first one in table order wins on rank 0
            err = (type(ex).__name__, str(ex))
        finally:
            if ref_store is not None:
                ref_store.close()
    finally:
        pool.shutdown(wait=True)
    t2 = time.perf_counter()
    stats['rows'] = int(len(part[0]) + len(part[1]))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((part, err), gathered, dst=0, group=group)
    t3 = time.perf_counter()
    stats['seconds'] = {'shard_plan_reference_broadcast': t1 - t0, 'walk_incl_contig_read': t2 - t1, 'gather': t3 - t2}
    prev = t0
    for name, t in marks:      # where the first phase goes (each mark: seconds since the previous one)
        stats['seconds']['setup_' + name] = t - prev
        prev = t
    last_dist_stats = stats
    if rank != 0:
        return None
    errs = [(int(shards[r][0]) if len(shards[r]) else n_rec, e) for r, (_, e) in enumerate(gathered) if e is not None]
    if errs:
        # the reference would have stopped at the earliest offending record; report the error of the shard that
        # contains the smallest table row (exact when a single record is malformed)
        _, (kind, msg) = min(errs, key=lambda x: x[0])
        raise (IndexError if kind == 'IndexError' else RuntimeError)(msg)
    snv, indel = merge_rows([p for p, _ in gathered], shards)
    if ref_arr is None:        # (walk_fn runs: nothing was read up front)
        ref_arr = [ref_fa.fetch_array(n) for n in full.ref_names]
        tig_all = [tig_fa.fetch_array(n) for n in full.tig_names]
    tig_arr = tig_all
    stats['seconds']['merge_rank0'] = time.perf_counter() - t3
    t3 = time.perf_counter()
    out = cigarcall.build_frames(snv, indel, full.chrom, full.qry, full.rev, full.align_index, ref_arr, tig_arr, full.ref_id,
                                 full.qry_id, hap, version_id)
    stats['seconds']['frames_rank0'] = time.perf_counter() - t3
    return out
