"""Flagging rules between the CIGAR walk and the inversion scan, as sorted scans over columns.

The reference implements these inside Snakemake ``run:`` blocks as ``DataFrame.iterrows()`` loops that build one
``pd.Series`` per output row -- over the multi-million-row tables the CIGAR walk now returns in milliseconds:

    cigar_filter          rules/call.snakefile:813-846        FILTER = PASS / TRIM against the trimmed alignments
    call_cigar(_to_files) rules/call.snakefile:800-846        the whole rule body (walk on the GPU + FILTER [+ the two bed.gz files])
    call_cigar_merge      rules/call.snakefile:755-790        concatenate + sort the per-batch tables
    cluster_variants      rules/call_inv.snakefile:616-690    clusters of SNVs / indels
    flag_insdel_cluster   rules/call_inv.snakefile:493-583    INS matched to nearby DELs, merged
    merge_flagged_loci    rules/call_inv.snakefile:331-474    merge of the four flag tables, TRY_INV, BATCH
    call_inv_batch        rules/call_inv.snakefile:127-311    flagged regions of a batch -> inversion calls (all loci per GPU batch)

Every loop there carries only "the previous row" (or a running maximum), so each is a break-point mask + segment
reduction here. The reference's behaviours that look accidental are kept, because the tables must come out identical
(tests/golden/flag/, produced by executing the reference's own rule bodies): the cluster rule ignores
``cluster_win_min`` (it reads ``cluster_win`` twice, :617-618), clusters variants by midpoint in POS order, and the
INS/DEL merge never records its last open interval (:551-571).

Host-side numpy: these tables are already pandas frames on the host and a pass over them is cheaper than a round trip to
the device; the GPU work of this stage is the walk that produced them.
"""
import numpy as np
import pandas as pd

BATCH_COUNT_DEFAULT = 60   # rules/call_inv.snakefile:81


# ------------------------------------------------------------------------------------------------ call_cigar FILTER
def _as_int64(col):
    """Column of ints (object dtype in the call tables) -> int64 array, converted in C where numpy can."""
    a = col.to_numpy()
    try:
        return a.astype(np.int64)
    except (TypeError, ValueError):
        return np.array([int(x) for x in a.tolist()], dtype=np.int64)


def cigar_filter(df, df_trim):
    """FILTER column of a call table: PASS when the variant lies strictly inside the trimmed alignment of its
    ``ALIGN_INDEX``, TRIM otherwise (also when the record disappeared in trimming). ``df_trim``: POS, END indexed by INDEX."""
    if df.shape[0] == 0:
        return pd.Series([], index=df.index, dtype=object)
    idx = df_trim.index.to_numpy()
    order = np.argsort(idx, kind='stable')
    idx_s = idx[order]
    ai = _as_int64(df['ALIGN_INDEX'])
    k = np.searchsorted(idx_s, ai)
    k_ok = np.minimum(k, max(len(idx_s) - 1, 0))
    found = (k < len(idx_s)) & (idx_s[k_ok] == ai) if len(idx_s) else np.zeros(len(ai), dtype=bool)
    t_pos = np.where(found, df_trim['POS'].to_numpy().astype(np.int64)[order][k_ok], -1) if len(idx_s) else np.full(len(ai), -1)
    t_end = np.where(found, df_trim['END'].to_numpy().astype(np.int64)[order][k_ok], -1) if len(idx_s) else np.full(len(ai), -1)
    pos, end = _as_int64(df['POS']), _as_int64(df['END'])
    ok = (pos > t_pos) & (end < t_end)
    return pd.Series(np.where(ok, 'PASS', 'TRIM').astype(object), index=df.index)


def call_cigar(df_align, batch, ref_fa_name, tig_fa_name, hap, df_trim):
    """Body of ``rule call_cigar`` (rules/call.snakefile:800-846): walk the records of one CALL_BATCH on the GPU, add FILTER."""
    from . import cigarcall
    df_align = df_align.loc[df_align['CALL_BATCH'] == batch]
    df_snv, df_insdel = cigarcall.make_insdel_snv_calls(df_align, ref_fa_name, tig_fa_name, hap, version_id=False)
    df_snv['FILTER'] = cigar_filter(df_snv, df_trim)
    df_insdel['FILTER'] = cigar_filter(df_insdel, df_trim)
    return df_snv, df_insdel


def _trim_bounds(table, df_trim):
    """Per alignment record: (POS, END) of its trimmed alignment, (-1, -1) when the record did not survive trimming."""
    idx = df_trim.index.to_numpy()
    order = np.argsort(idx, kind='stable')
    idx_s = idx[order]
    ai = np.array([int(x) for x in table.align_index.tolist()], dtype=np.int64)
    if not len(idx_s):
        return np.full(len(ai), -1, dtype=np.int64), np.full(len(ai), -1, dtype=np.int64)
    k = np.minimum(np.searchsorted(idx_s, ai), len(idx_s) - 1)
    found = idx_s[k] == ai
    return (np.where(found, df_trim['POS'].to_numpy().astype(np.int64)[order][k], -1),
            np.where(found, df_trim['END'].to_numpy().astype(np.int64)[order][k], -1))


def write_gzip_members(path, data, threads=None, block=8 << 20, level=6):
    """``data`` (bytes) -> gzip file made of independently compressed members (cut at line ends), compressed by a thread pool
    (zlib releases the GIL). Any gzip reader, ``pd.read_csv`` included, reads the concatenation as one stream."""
    import os
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    cuts, p = [0], 0
    while len(data) - p > block:
        q = data.find(b'\n', p + block)
        if q < 0:
            break
        p = q + 1
        cuts.append(p)
    cuts.append(len(data))

    def comp(i):
        c = zlib.compressobj(level, zlib.DEFLATED, 31)
        return c.compress(data[cuts[i]:cuts[i + 1]]) + c.flush()
    n = len(cuts) - 1
    with ThreadPoolExecutor(max_workers=max(1, min(threads or (os.cpu_count() or 4), n))) as pool, open(path, 'wb') as fh:
        for part in pool.map(comp, range(n)):
            fh.write(part)


def call_cigar_to_files(df_align, batch, ref_fa_name, tig_fa_name, hap, df_trim, bed_insdel, bed_snv, threads=None):
    """Body of ``rule call_cigar`` including its two ``to_csv(..., compression='gzip')`` calls (rules/call.snakefile:800-846): the
    tables go from device rows to TSV text in C and to disk through a parallel gzip writer -- no DataFrame, no per-row Python
    object. The files decompress to exactly what the reference's rule writes. Returns ``(n_snv, n_insdel)``."""
    from . import cigarcall
    df_align = df_align.loc[df_align['CALL_BATCH'] == batch]
    text = None
    if df_align.shape[0] > 0:
        table, snv, indel, ref_arr, tig_arr, _ = cigarcall.call_rows(df_align, ref_fa_name, tig_fa_name)
        t_pos, t_end = _trim_bounds(table, df_trim)
        pass_snv = ((snv['pos_ref'] > t_pos[snv['rec']]) & (snv['pos_ref'].astype(np.int64) + 1 < t_end[snv['rec']])).astype(np.uint8)
        pass_indel = ((indel['pos'] > t_pos[indel['rec']]) & (indel['end'] < t_end[indel['rec']])).astype(np.uint8)
        text = cigarcall.tables_tsv(snv, indel, table.chrom, table.qry, table.rev, table.align_index, ref_arr, tig_arr, table.ref_id, table.qry_id,
                                    hap, False, pass_snv, pass_indel)
        n = (len(snv), len(indel))
    if text is None:   # empty batch, or names that need CSV quoting: the frames + pandas
        df_snv, df_insdel = call_cigar(df_align, batch, ref_fa_name, tig_fa_name, hap, df_trim) if df_align.shape[0] else \
            (cigarcall._empty(cigarcall.SNV_COLUMNS), cigarcall._empty(cigarcall.INSDEL_COLUMNS))
        if df_align.shape[0] == 0:
            df_snv['FILTER'] = pd.Series([], dtype=object)
            df_insdel['FILTER'] = pd.Series([], dtype=object)
        df_insdel.to_csv(bed_insdel, sep='\t', index=False, compression='gzip')
        df_snv.to_csv(bed_snv, sep='\t', index=False, compression='gzip')
        return len(df_snv), len(df_insdel)
    write_gzip_members(bed_snv, text[0], threads)
    write_gzip_members(bed_insdel, text[1], threads)
    return n


def call_cigar_merge(bed_insdel_files, bed_snv_files, bed_insdel_out, bed_snv_out, threads=None):
    """Body of ``rule call_cigar_merge`` (rules/call.snakefile:755-790): concatenate the per-batch call tables and sort them
    (INS/DEL by #CHROM, POS, END, ID; SNVs by #CHROM, POS). The reference parses every column of every file into a DataFrame and
    formats all of them again; a row of the output is a line of the input, so only the key columns go through pandas here (read
    and sorted exactly as the reference does, type inference included) and the lines themselves are re-ordered as bytes and
    written through the parallel gzip writer."""
    import gzip
    for files, out, keys in ((bed_insdel_files, bed_insdel_out, ['#CHROM', 'POS', 'END', 'ID']), (bed_snv_files, bed_snv_out, ['#CHROM', 'POS'])):
        headers, lines, key_frames = [], [], []
        for fn in files:
            with gzip.open(fn, 'rb') as fh:
                data = fh.read()
            first, _, body = data.partition(b'\n')
            headers.append(first)
            rows = body.split(b'\n')
            if rows and rows[-1] == b'':
                rows.pop()
            lines.extend(rows)
            key_frames.append(pd.read_csv(fn, sep='\t', keep_default_na=False, usecols=keys))
        if len(set(headers)) > 1:
            raise RuntimeError('call_cigar_merge: batch tables have different columns')
        df_key = pd.concat(key_frames, axis=0).reset_index(drop=True)
        if df_key.shape[0] != len(lines):
            raise RuntimeError('call_cigar_merge: a batch table has fields with embedded newlines; merge it through pandas')
        order = df_key.sort_values(keys).index.to_numpy()
        text = headers[0] + b'\n' + b''.join(lines[i] + b'\n' for i in order.tolist()) if headers else b''
        write_gzip_members(out, text, threads)


# ------------------------------------------------------------------------------------------------ call_inv_batch
INV_BED_COLUMNS = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'HAP', 'QRY_REGION', 'QRY_STRAND', 'CI', 'RGN_REF_INNER', 'RGN_QRY_INNER',
                   'RGN_REF_DISC', 'RGN_QRY_DISC', 'FLAG_ID', 'FLAG_TYPE', 'ALIGN_INDEX', 'CALL_SOURCE', 'FILTER', 'SEQ']


def _collapse_to_set(values, to_type=None):
    """pavlib/util.py:107-122."""
    stack, out = list(values), set()
    while stack:
        v = stack.pop()
        if isinstance(v, (tuple, list)):
            stack.extend(v)
        else:
            out.add(to_type(v) if to_type is not None else v)
    return out


class _KUtil:
    def __init__(self, k_size):
        self.k_size = k_size


def call_inv_batch(df_flag, batch, ref_fa_name, tig_fa_name, df_aln, df_fai, hap, k_size=31, inv_region_limit=None, inv_min_expand=None,
                   srs_list=None, log=None, density_out_dir=None, threads=1):
    """Body of ``rule call_inv_batch`` (rules/call_inv.snakefile:127-311): resolve the flagged regions of one batch into inversion
    calls. Every expansion round of every region still open is scored in one GPU batch (``scan_for_inv_batch``) instead of one
    ``scripts/density.py`` process per region and expansion; the table is the reference's, row for row (duplicates of an inversion
    found through a second flagged region are dropped, columns and their order are kept -- including the missing FILTER column
    of the reference's empty table when a non-empty batch yields no call).

    ``df_flag``: flagged regions (``merge_flagged_loci``); ``df_aln`` / ``df_fai``: trimmed alignments and contig lengths for the
    coordinate lift; ``log``: open text file or None; ``density_out_dir``: where to write the per-call density tables, or None."""
    import os

    from . import inv, lift, seq
    df_flag = df_flag.loc[df_flag['BATCH'] == batch]
    if df_flag.shape[0] == 0:
        return pd.DataFrame([], columns=INV_BED_COLUMNS)
    srs_tree = inv.get_srs_tree(srs_list)
    align_lift = lift.AlignLift(df_aln, df_fai)
    # (columns as lists: a Series per row through iterrows() was a tenth of the host time of this rule)
    f_chrom, f_pos, f_end, f_type = (df_flag[c].tolist() for c in ('#CHROM', 'POS', 'END', 'TYPE'))
    regions = [seq.Region(c, p, e) for c, p, e in zip(f_chrom, f_pos, f_end)]
    calls = inv.scan_for_inv_batch(regions, ref_fa_name, tig_fa_name, align_lift, _KUtil(k_size), max_region_size=inv_region_limit, log=log,
                                   srs_tree=srs_tree, min_exp_count=inv_min_expand, catch=True)
    id_set, rows = set(), []
    for flag_type, inv_call in zip(f_type, calls):
        if inv_call is None or isinstance(inv_call, RuntimeError) or inv_call.id in id_set:   # errors are logged and skipped (call_inv.snakefile:198-200)
            continue
        sequence = seq.region_seq_fasta(inv_call.region_tig_outer, tig_fa_name, rev_compl=inv_call.region_tig_outer.is_rev)
        align_index = ','.join(sorted(_collapse_to_set((inv_call.region_ref_outer.pos_aln_index, inv_call.region_ref_outer.end_aln_index,
                                                         inv_call.region_ref_inner.pos_aln_index, inv_call.region_ref_inner.end_aln_index), to_type=str)))
        rows.append([
            inv_call.region_ref_outer.chrom, inv_call.region_ref_outer.pos, inv_call.region_ref_outer.end, inv_call.id, 'INV', inv_call.svlen, hap,
            inv_call.region_tig_outer.to_base1_string(), '-' if inv_call.region_tig_outer.is_rev else '+', 0,
            inv_call.region_ref_inner.to_base1_string(), inv_call.region_tig_inner.to_base1_string(),
            inv_call.region_ref_discovery.to_base1_string(), inv_call.region_tig_discovery.to_base1_string(),
            inv_call.region_flag.region_id(), flag_type, align_index, inv.CALL_SOURCE, 'PASS', sequence])
        id_set.add(inv_call.id)
        if density_out_dir is not None:
            os.makedirs(density_out_dir, exist_ok=True)
            inv_call.df.to_csv(os.path.join(density_out_dir, 'density_{}_{}.tsv.gz'.format(inv_call.id, hap)), sep='\t', index=False, compression='gzip')
    if not rows:
        return pd.DataFrame([], columns=[c for c in INV_BED_COLUMNS if c != 'FILTER'])   # sic (call_inv.snakefile:293-307)
    df_bed = pd.DataFrame(np.array(rows, dtype=object), columns=INV_BED_COLUMNS)
    return df_bed.sort_values(['#CHROM', 'POS', 'END', 'ID'])


# ------------------------------------------------------------------------------------------------ helpers
def _chrom_codes(chrom):
    """Integer codes that sort like the strings do (one hash pass over the column, then a sort of the few distinct names)."""
    codes, uniq = pd.factorize(np.asarray(chrom, dtype=object), sort=False)
    uniq = np.array([str(u) for u in uniq], dtype=object)
    rank = np.empty(len(uniq), dtype=np.int64)
    order = np.argsort(uniq.astype(str), kind='stable')
    rank[order] = np.arange(len(uniq))
    return rank[codes] if len(uniq) else codes.astype(np.int64), uniq[order].astype(str) if len(uniq) else np.zeros(0, dtype=str)


def _sorted_by_chrom_pos(df):
    """``df.sort_values(['#CHROM', 'POS'])`` (stable) as a positional permutation."""
    codes, _ = _chrom_codes(df['#CHROM'].to_numpy())
    return np.lexsort((df['POS'].to_numpy().astype(np.int64), codes))


def _segments(breaks):
    """Start / end (exclusive) positions of the runs delimited by ``breaks`` (True = a new run starts here)."""
    starts = np.flatnonzero(breaks)
    ends = np.concatenate((starts[1:], [len(breaks)]))
    return starts, ends


# ------------------------------------------------------------------------------------------------ call_inv_cluster
def cluster_variants(tables, vartype, cluster_win=200, cluster_win_min=500, cluster_min_snv=20, cluster_min_indel=10):
    """Clusters of SNVs (``vartype='snv'``) or indels (``'indel'``). ``tables``: call tables with #CHROM POS END SVTYPE SVLEN FILTER
    (the rule reads the INS/DEL table for indels and the SNV table for SNVs). -> #CHROM POS END COUNT."""
    if vartype == 'indel':
        cluster_min = cluster_min_indel
    elif vartype == 'snv':
        cluster_min = cluster_min_snv
    else:
        raise RuntimeError('Bad variant type {}: Expected "indel" or "snv"')
    win_min = cluster_win   # sic: the reference assigns params.cluster_win to cluster_win_min (call_inv.snakefile:617-618)
    df = pd.concat([t[['#CHROM', 'POS', 'END', 'SVTYPE', 'SVLEN', 'FILTER']] for t in tables], axis=0)
    df = df.iloc[_sorted_by_chrom_pos(df)]
    df = df.loc[df['FILTER'] == 'PASS']
    if vartype == 'indel':
        df = df.loc[df['SVLEN'] < 50]
    empty = pd.DataFrame([], columns=['#CHROM', 'POS', 'END', 'COUNT'])
    if df.shape[0] == 0:
        return empty
    codes, names = _chrom_codes(df['#CHROM'].to_numpy())
    mid = (df['END'].to_numpy().astype(np.int64) + df['POS'].to_numpy().astype(np.int64)) // 2   # DEL to midpoint, in POS order
    brk = np.ones(len(mid), dtype=bool)
    brk[1:] = ~((mid[1:] < mid[:-1] + cluster_win) & (codes[1:] == codes[:-1]))
    s, e = _segments(brk)
    c_pos, c_end, count = mid[s], mid[e - 1], e - s
    keep = (count >= cluster_min) & (c_end - c_pos >= win_min)
    if not keep.any():
        return empty
    return pd.DataFrame({'#CHROM': names[codes[s][keep]].astype(object), 'POS': c_pos[keep], 'END': c_end[keep], 'COUNT': count[keep]},
                        columns=['#CHROM', 'POS', 'END', 'COUNT']).reset_index(drop=True)


# ------------------------------------------------------------------------------------------------ call_inv_flag_insdel_cluster
def flag_insdel_cluster(df, vartype, flank_cluster=2, flank_merge=2000, cluster_min_svlen=4):
    """Loci where an insertion sits next to deletions (``vartype`` 'sv': SVLEN >= 50, 'indel': cluster_min_svlen <= SVLEN < 50).
    ``df``: merged INS/DEL calls with #CHROM POS END ID SVTYPE SVLEN FILTER. -> #CHROM POS END."""
    svlen_min = cluster_min_svlen if vartype == 'indel' else 50
    empty = pd.DataFrame([], columns=['#CHROM', 'POS', 'END'])
    df = df.loc[df['FILTER'] == 'PASS']
    df = df.loc[df['SVLEN'] >= svlen_min]
    if vartype == 'indel':
        df = df.loc[df['SVLEN'] < 50]
    if df.shape[0] == 0:
        return empty
    codes, names = _chrom_codes(df['#CHROM'].to_numpy())
    svtype = df['SVTYPE'].to_numpy()
    pos, end = df['POS'].to_numpy().astype(np.int64), df['END'].to_numpy().astype(np.int64)
    svlen = df['SVLEN'].to_numpy().astype(np.int64)
    is_ins, is_del = svtype == 'INS', svtype == 'DEL'
    m_chrom, m_pos, m_end = [], [], []
    # Every INS against the DELs of its chromosome that overlap [POS - flank, POS + flank): with the DELs sorted by POS and a
    # running maximum of END, the candidates are a prefix range; the exact overlap test runs on that range only.
    for c in np.unique(codes[is_ins]).tolist():
        d = np.flatnonzero(is_del & (codes == c))
        if not len(d):
            continue
        d = d[np.argsort(pos[d], kind='stable')]
        d_pos, d_end = pos[d], end[d]
        run_end = np.maximum.accumulate(d_end)
        ins = np.flatnonzero(is_ins & (codes == c))
        lo_q, hi_q = pos[ins] - svlen[ins] * flank_cluster, pos[ins] + svlen[ins] * flank_cluster
        hi_i = np.searchsorted(d_pos, hi_q, side='left')            # DELs with POS < hi
        lo_i = np.searchsorted(run_end, lo_q, side='right')          # first DEL whose running END exceeds lo
        for a, b, lo in zip(lo_i.tolist(), hi_i.tolist(), lo_q.tolist()):
            if b <= a:
                continue
            sel = d_end[a:b] > lo
            if sel.any():
                m_chrom.append(c)
                m_pos.append(int(d_pos[a:b][sel].min()))
                m_end.append(int(d_end[a:b][sel].max()))
    if not m_chrom:
        return empty
    m_chrom, m_pos, m_end = np.array(m_chrom), np.array(m_pos, dtype=np.int64), np.array(m_end, dtype=np.int64)
    order = np.lexsort((m_pos, m_chrom))
    m_chrom, m_pos, m_end = m_chrom[order], m_pos[order], m_end[order]
    # merge intervals closer than flank_merge (running END per chromosome)
    brk = np.ones(len(m_pos), dtype=bool)
    run_end = m_end.copy()
    for c in np.unique(m_chrom).tolist():
        w = np.flatnonzero(m_chrom == c)
        run_end[w] = np.maximum.accumulate(m_end[w])
        brk[w[1:]] = m_pos[w[1:]] - flank_merge > run_end[w[:-1]]
    s, e = _segments(brk)
    s, e = s[:-1], e[:-1]       # sic: the reference never records the interval still open when its loop ends (call_inv.snakefile:551-571)
    if not len(s):
        return empty
    out = pd.DataFrame({'#CHROM': names[m_chrom[s]].astype(object), 'POS': m_pos[s], 'END': run_end[e - 1]}, columns=['#CHROM', 'POS', 'END'])
    return out.reset_index(drop=True)


# ------------------------------------------------------------------------------------------------ call_inv_merge_flagged_loci
def accept_flagged_region(type_set, allow_single_cluster=False, match_any=frozenset()):
    """rules/call_inv.snakefile:56-79."""
    if not allow_single_cluster and (type_set == {'CLUSTER_SNV'} or type_set == {'CLUSTER_INDEL'}):
        return False
    if match_any and not type_set & match_any:
        return False
    return True


_TYPE_BITS = {'MATCH_SV': 1, 'MATCH_INDEL': 2, 'CLUSTER_INDEL': 4, 'CLUSTER_SNV': 8}


def merge_flagged_loci(df_insdel_sv, df_insdel_indel, df_cluster_indel, df_cluster_snv, flank=500, batch_count=BATCH_COUNT_DEFAULT,
                       inv_sig_filter='svindel'):
    """Merge the four flag tables into candidate regions, decide which ones the inversion caller tries (TRY_INV) and deal them
    round-robin into ``batch_count`` batches. -> #CHROM POS END ID SVTYPE SVLEN TYPE COUNT_INDEL COUNT_SNV TRY_INV BATCH."""
    allow_single_cluster, match_any = False, set()
    if inv_sig_filter is not None:
        if inv_sig_filter == 'single_cluster':
            allow_single_cluster = True
        elif inv_sig_filter == 'svindel':
            match_any = {'MATCH_SV', 'MATCH_INDEL'}
        elif inv_sig_filter == 'sv':
            match_any = {'MATCH_SV'}
        else:
            raise RuntimeError(f'Unrecognized region filter: {inv_sig_filter} (must be "single_cluster", "svindel", or "sv")')
    columns = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'TYPE', 'COUNT_INDEL', 'COUNT_SNV', 'TRY_INV', 'BATCH']
    chrom, pos, end, c_indel, c_snv, bits = [], [], [], [], [], []
    for df, name, col in ((df_insdel_sv, 'MATCH_SV', None), (df_insdel_indel, 'MATCH_INDEL', None),
                          (df_cluster_indel, 'CLUSTER_INDEL', 'COUNT_INDEL'), (df_cluster_snv, 'CLUSTER_SNV', 'COUNT_SNV')):
        n = df.shape[0]
        if n == 0:
            continue
        chrom.append(df['#CHROM'].to_numpy().astype(str))
        pos.append(df['POS'].to_numpy().astype(np.int64))
        end.append(df['END'].to_numpy().astype(np.int64))
        cnt = df['COUNT'].to_numpy().astype(np.int64) if col else np.zeros(n, dtype=np.int64)
        c_indel.append(cnt if col == 'COUNT_INDEL' else np.zeros(n, dtype=np.int64))
        c_snv.append(cnt if col == 'COUNT_SNV' else np.zeros(n, dtype=np.int64))
        bits.append(np.full(n, _TYPE_BITS[name], dtype=np.int64))
    if not chrom:
        return pd.DataFrame([], columns=columns)
    chrom, pos, end = np.concatenate(chrom), np.concatenate(pos), np.concatenate(end)
    c_indel, c_snv, bits = np.concatenate(c_indel), np.concatenate(c_snv), np.concatenate(bits)
    codes, names = _chrom_codes(chrom)
    order = np.lexsort((pos, codes))       # concat order is kept among equal (#CHROM, POS), like the stable sort of the reference
    codes, pos, end, c_indel, c_snv, bits = codes[order], pos[order], end[order], c_indel[order], c_snv[order], bits[order]
    # a row joins the open region when it starts before (END of the *previous row*) + flank -- the reference overwrites the region
    # end with every row it adds (call_inv.snakefile:404-412)
    brk = np.ones(len(pos), dtype=bool)
    brk[1:] = ~((pos[1:] < end[:-1] + flank) & (codes[1:] == codes[:-1]))
    s, e = _segments(brk)
    r_pos, r_end = pos[s], end[e - 1]
    r_indel = np.add.reduceat(c_indel, s)
    r_snv = np.add.reduceat(c_snv, s)
    r_bits = np.bitwise_or.reduceat(bits, s)
    r_chrom = names[codes[s]]
    # pd.concat(...).T.sort_values(['#CHROM', 'POS']) of the regions: already in that order except that regions of one chromosome
    # may repeat a POS; the stable sort keeps emission order for those
    order = np.lexsort((r_pos, codes[s]))
    r_chrom, r_pos, r_end, r_indel, r_snv, r_bits = r_chrom[order], r_pos[order], r_end[order], r_indel[order], r_snv[order], r_bits[order]
    type_sets = [{t for t, b in _TYPE_BITS.items() if v & b} for v in r_bits.tolist()]
    try_inv = np.array([accept_flagged_region(t, allow_single_cluster, match_any) for t in type_sets], dtype=bool)
    batch = np.full(len(r_pos), -1, dtype=np.int64)
    batch[try_inv] = np.arange(int(try_inv.sum())) % int(batch_count)
    return pd.DataFrame({
        '#CHROM': r_chrom.astype(object), 'POS': r_pos, 'END': r_end,
        'ID': [f'{c}-{p}-RGN-{e_ - p}' for c, p, e_ in zip(r_chrom.tolist(), r_pos.tolist(), r_end.tolist())],
        'SVTYPE': 'RGN', 'SVLEN': r_end - r_pos, 'TYPE': [','.join(sorted(t)) for t in type_sets],
        'COUNT_INDEL': r_indel, 'COUNT_SNV': r_snv, 'TRY_INV': try_inv, 'BATCH': batch,
    }, columns=columns)
