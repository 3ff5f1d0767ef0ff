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
    snv = snv[np.lexsort((snv['pos_ref'], snv['op_idx'], snv['rec']))]
    indel = indel[np.lexsort((indel['op_idx'], indel['rec']))]
    return snv, indel


def make_insdel_snv_calls_dist(df_align, ref_fa_name, tig_fa_name, hap, version_id=True, group=None, walk_fn=None):
    """Distributed ``make_insdel_snv_calls``: call on every rank with the same arguments; rank 0 returns
    ``(df_snv, df_insdel)`` identical to the single-GPU result, the other ranks return ``None``.

    ``walk_fn(table, ref_arr, tig_arr) -> (snv, indel)`` replaces the device walk (CPU tests of the host
    logic inject a checker here); by default the walk runs on this rank's GPU against the reference planes
    broadcast from rank 0.
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_rec = df_align.shape[0]
    if n_rec == 0:
        return (cigarcall._empty(cigarcall.SNV_COLUMNS), cigarcall._empty(cigarcall.INSDEL_COLUMNS)) if rank == 0 else None
    full = cigarcall.AlignTable(df_align)
    span = None
    if 'END' in df_align.columns:
        span = df_align['END'].to_numpy(dtype=np.int64) - full.pos
    shards = lpt_shards(record_costs(full.cigars, span), world)
    mine = df_align.iloc[shards[rank]]
    ref_fa, tig_fa = fasta.open_fasta(ref_fa_name), fasta.open_fasta(tig_fa_name)
    part = (np.zeros(0, device._capi.SNV_ROW), np.zeros(0, device._capi.INDEL_ROW))
    err = None
    try:
        if walk_fn is not None:
            if len(mine):
                t = cigarcall.AlignTable(mine)
                part = walk_fn(t, [ref_fa.fetch_array(n) for n in t.ref_names], [tig_fa.fetch_array(n) for n in t.tig_names])
        else:
            ctx = device.get_context()
            names = list(full.ref_names)  # every rank holds the whole reference of this table, same order
            if rank == 0:
                ref_store = device.SeqStore(ctx, names, [ref_fa.fetch_array(n) for n in names], keep_host=False)
                uid = [device.nccl_unique_id()]
            else:
                ref_store = device.SeqStore.from_packed(ctx, names, [ref_fa.length(n) for n in names], None, None)
                uid = [None]
            dist.broadcast_object_list(uid, src=0, group=group)
            ref_store.broadcast(uid[0], rank, world)
            try:
                if len(mine):
                    t = cigarcall.AlignTable(mine)
                    # ids must index the broadcast store, not the shard-local name table
                    t.ref_id = np.array([full.ref_names[str(c)] for c in t.chrom.tolist()], dtype=np.int32)
                    part = cigarcall.walk_rows(t, None, [tig_fa.fetch_array(n) for n in t.tig_names], ctx=ctx, ref_store=ref_store)
            finally:
                ref_store.close()
    except (RuntimeError, IndexError) as ex:  # CIGAR errors: first one in table order wins on rank 0
        err = (type(ex).__name__, str(ex))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((part, err), gathered, dst=0, group=group)
    if rank != 0:
        return None
    errs = [(int(shards[r][0]) if len(shards[r]) else n_rec, e) for r, (_, e) in enumerate(gathered) if e is not None]
    if errs:
        # the reference would have stopped at the earliest offending record; report the error of the shard that
        # contains the smallest table row (exact when a single record is malformed)
        _, (kind, msg) = min(errs, key=lambda x: x[0])
        raise (IndexError if kind == 'IndexError' else RuntimeError)(msg)
    snv, indel = merge_rows([p for p, _ in gathered], shards)
    ref_arr = [ref_fa.fetch_array(n) for n in full.ref_names]
    tig_arr = [tig_fa.fetch_array(n) for n in full.tig_names]
    return cigarcall.build_frames(snv, indel, full.chrom, full.qry, full.rev, full.align_index, ref_arr, tig_arr, full.ref_id,
                                  full.qry_id, hap, version_id)
