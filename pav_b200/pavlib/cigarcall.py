"""Call SNVs and INS/DEL variants from CIGAR strings on the GPU.

Drop-in for the reference's ``pavlib.cigarcall`` (pavlib/cigarcall.py): same function name,
arguments, return value ``(df_snv, df_insdel)``, column order, dtypes (all ``object``), row order
(stable sort on ``#CHROM, POS, END, ID``), index values and exception messages. The per-record
Python loop of the reference (:50-311) is replaced by one ``pavgpu_cigar_call``; this module only
moves bytes in (FASTA -> HBM, CIGAR text -> packed ops) and formats the rows that come back.
"""
import os
import time

import numpy as np
import pandas as pd

from .. import device, fasta, sidecar
from . import variant

CALL_SOURCE = 'CIGAR'          # pavlib/cigarcall.py:19
CALL_CIGAR_BATCH_COUNT = 10    # pavlib/cigarcall.py:21 (read by rules/align.snakefile:163)

SNV_COLUMNS = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'REF', 'ALT', 'HAP', 'QRY_REGION', 'QRY_STRAND', 'CI',
               'ALIGN_INDEX', 'CALL_SOURCE']
INSDEL_COLUMNS = ['#CHROM', 'POS', 'END', 'ID', 'SVTYPE', 'SVLEN', 'HAP', 'QRY_REGION', 'QRY_STRAND', 'CI', 'ALIGN_INDEX',
                  'LEFT_SHIFT', 'HOM_REF', 'HOM_TIG', 'CALL_SOURCE', 'SEQ']

_CHR = np.array([chr(i) for i in range(256)], dtype=object)
_OP_CHAR = 'MIDNSHP=X'

# statistics of the last call (device timings etc.), for bench.py
last_stats = None
last_phase_seconds = None   # host-side phase breakdown of the last make_insdel_snv_calls
last_walk_seconds = None    # breakdown of the device phase (stores, CIGAR tokenizer, pavgpu_cigar_call)


def _first_seen(values):
    seen = {}
    for v in values:
        if v not in seen:
            seen[v] = len(seen)
    return seen


def _sort_order(chrom_codes, pos, end, id_of, end_is_pos_plus_1=False):
    """Permutation equal to pandas' stable ``sort_values(['#CHROM','POS','END','ID'])``. ``id_of(i)`` gives the ID of
    emission row ``i``; it is only asked for rows that tie on (chrom, pos, end)."""
    if end_is_pos_plus_1 and len(pos) and int(pos.max()) < (1 << 40) and int(chrom_codes.max()) < (1 << 22):
        # SNV rows: END = POS + 1, so one composite key orders them. Rows arrive in (record, op) order and the records are sorted
        # by #CHROM, POS: without overlapping records the rows are sorted already (one linear check instead of a sort)
        key = (chrom_codes << 40) | pos
        if bool((key[1:] >= key[:-1]).all()):
            order = np.arange(len(key), dtype=np.int64)
        else:
            order = np.argsort(key, kind='stable')
    elif len(pos) and int(pos.max()) < (1 << 40) and int(chrom_codes.max()) < (1 << 22) and \
            bool((((chrom_codes[1:] << 40) | pos[1:]) > ((chrom_codes[:-1] << 40) | pos[:-1])).all()):
        return np.arange(len(pos), dtype=np.int64)      # strictly increasing (#CHROM, POS): sorted, and no ties to break
    else:
        order = np.lexsort((end, pos, chrom_codes))
    if len(order) > 1:
        c, p, e = chrom_codes[order], pos[order], end[order]
        tie = (c[1:] == c[:-1]) & (p[1:] == p[:-1]) & (e[1:] == e[:-1])
        if tie.any():
            # groups of equal (chrom, pos, end): order them by ID, stable
            starts = np.flatnonzero(tie & ~np.concatenate(([False], tie[:-1])))
            for s in starts.tolist():
                t = s + 1
                while t < len(tie) and tie[t]:
                    t += 1
                grp = order[s:t + 1]
                grp_ids = [id_of(i) for i in grp.tolist()]
                order[s:t + 1] = grp[np.array(sorted(range(len(grp)), key=grp_ids.__getitem__), dtype=np.int64)]
    return order


def _empty(columns):
    return pd.DataFrame([], columns=columns)


def _raise_illegal(err, chrom, qry, align_index):
    rec = err.rec
    op = _OP_CHAR[err.opcode] if err.opcode < len(_OP_CHAR) else '?'
    if op == 'M':
        raise RuntimeError((
            'Illegal operation code in CIGAR string at operation {}: '
            'Alignments must be generated with =/X (not M): '
            'opcode={}, subject={}:{}, query={}:{}, align-index={}'
        ).format(err.op_index + 1, op, chrom[rec], err.pos_ref, qry[rec], err.pos_qry, align_index[rec]))
    raise RuntimeError((
        'Illegal operation code in CIGAR string at operation {}: '
        'opcode={}, subject={}:{} , query={}:{}, align-index={}'
    ).format(err.op_index + 1, op, chrom[rec], err.pos_ref, qry[rec], err.pos_qry, align_index[rec]))


def _raise_parse(perr, chrom, qry, pos):
    r = perr.rec
    if perr.code == 2:
        raise RuntimeError('Missing length in CIGAR string for contig {} alignment starting at {}:{}: CIGAR index {}'.format(
            qry[r], chrom[r], pos[r], perr.text_pos))
    if perr.code == 3:
        raise RuntimeError('Unknown CIGAR operation for contig {} alignment starting at {}:{}: CIGAR operation {}'.format(
            qry[r], chrom[r], pos[r], chr(perr.ch)))
    raise IndexError('string index out of range')


class AlignTable:
    """Columns of the alignment table the walk needs, extracted once (the caller's frame is never modified)."""

    def __init__(self, df_align):
        self.n_rec = df_align.shape[0]
        self.chrom = df_align['#CHROM'].to_numpy(dtype=object)
        self.qry = df_align['QRY_ID'].to_numpy(dtype=object)
        self.rev = np.array([bool(x) for x in df_align['REV'].tolist()], dtype=bool)
        self.pos = np.array([int(x) for x in df_align['POS'].tolist()], dtype=np.int64)
        self.align_index = df_align['INDEX'].to_numpy(dtype=object)
        self.cigars = df_align['CIGAR'].tolist()
        self.ref_names = _first_seen(str(c) for c in self.chrom.tolist())
        self.tig_names = _first_seen(str(q) for q in self.qry.tolist())
        self.ref_id = np.array([self.ref_names[str(c)] for c in self.chrom.tolist()], dtype=np.int32)
        self.qry_id = np.array([self.tig_names[str(q)] for q in self.qry.tolist()], dtype=np.int32)


class _Awaited:
    """``value`` once ``futures`` are done (future-like for walk_rows). A module-level class on purpose: a class defined inside
    the calling function would sit in a reference cycle and keep the 200 MB staging buffers alive until the cyclic GC runs."""

    def __init__(self, futures, value):
        self.futures, self.value = futures, value

    def result(self):
        for f in self.futures:
            f.result()
        return self.value


def _resolve(x):
    """Inputs of walk_rows may be ``concurrent.futures.Future`` objects (FASTA read / CIGAR tokenizer running beside the uploads)."""
    return x.result() if hasattr(x, 'result') else x


def walk_rows(table, ref_arr, tig_arr, ctx=None, ref_store=None, parsed=None):
    """Run the CIGAR walk for ``table`` on the GPU. ``ref_arr`` / ``tig_arr``: uint8 arrays in the order of
    ``table.ref_names`` / ``table.tig_names`` (or futures resolving to them). ``ref_store`` may be a resident store
    (multi-GPU: the broadcast reference) whose sequence order matches ``table.ref_names``. ``parsed``: result of
    ``device.parse_cigars(table.cigars)`` (or a future) when the caller tokenised the CIGARs already.

    Returns ``(snv rows, indel rows)`` in emission order; raises the reference's exceptions for bad CIGARs."""
    global last_stats, last_walk_seconds
    ctx = ctx or device.get_context()
    own_ref = ref_store is None
    tig_store = None
    t0 = time.perf_counter()
    try:
        if own_ref:
            ref_store = device.SeqStore(ctx, list(table.ref_names), _resolve(ref_arr), keep_host=False)
        t1 = time.perf_counter()
        tig_store = device.SeqStore(ctx, list(table.tig_names), _resolve(tig_arr), keep_host=False)
        t2 = time.perf_counter()
        ops, op_off, perr = _resolve(parsed) if parsed is not None else device.parse_cigars(table.cigars)
        t3 = time.perf_counter()
        snv, indel, cerr, stats = device.cigar_call(ctx, ref_store, tig_store, table.ref_id, table.qry_id,
                                                    table.pos.astype(np.int32), table.rev.astype(np.uint8), ops, op_off)
        t4 = time.perf_counter()
    finally:
        if own_ref and ref_store is not None:
            ref_store.close()
        if tig_store is not None:
            tig_store.close()
    last_walk_seconds = {'ref_store': t1 - t0, 'tig_store_incl_wait_for_read': t2 - t1, 'cigar_parse_wait': t3 - t2, 'cigar_call': t4 - t3,
                         'close': time.perf_counter() - t4}
    last_stats = stats.as_dict()
    # Errors surface in walk order (the reference raises lazily while iterating records and ops)
    if cerr.code == 1 and (perr.code == 0 or (cerr.rec, cerr.op_index) < (perr.rec, perr.op_index)):
        _raise_illegal(cerr, table.chrom, table.qry, table.align_index)
    if perr.code != 0:
        _raise_parse(perr, table.chrom, table.qry, table.pos)
    return snv, indel


def make_insdel_snv_calls(df_align, ref_fa_name, tig_fa_name, hap, version_id=True):
    """
    Parse variants from CIGAR strings.

    :param df_align: Post-cut BED of read alignments (needs ``#CHROM POS INDEX QRY_ID REV CIGAR``).
    :param ref_fa_name: Reference FASTA file name.
    :param tig_fa_name: Contig FASTA file name.
    :param hap: String identifying the haplotype ("h1", "h2").
    :param version_id: Version duplicate variant IDs if `True`.

    :return: ``(df_snv, df_insdel)`` -- the order the reference actually returns (pavlib/cigarcall.py:362).
    """
    global last_phase_seconds
    if df_align.shape[0] == 0:
        return _empty(SNV_COLUMNS), _empty(INSDEL_COLUMNS)
    table, snv, indel, ref_arr, tig_arr, phases = call_rows(df_align, ref_fa_name, tig_fa_name)
    t2 = time.perf_counter()
    frames = build_frames(snv, indel, table.chrom, table.qry, table.rev, table.align_index, ref_arr, tig_arr, table.ref_id,
                          table.qry_id, hap, version_id)
    phases['frames'] = time.perf_counter() - t2
    last_phase_seconds = phases
    return frames


def read_sequences(fa, names, pool, ctx=None, pinned=False):
    """-> (list of uint8 arrays in the order of ``names``, futures): the records are read by up to ``_READERS`` jobs of about equal
    size submitted to ``pool`` (file reads, numpy copies and the C calls release the GIL); with ``pinned`` the bases land in one
    pinned staging buffer of the context's pool (H2D at PCIe speed) when they fit ``PAVGPU_PINNED_STAGING_MAX_MB``."""
    names = list(names)
    for nm in names:
        if str(nm) not in fa.index:
            raise KeyError(f'sequence {str(nm)!r} not found in {fa.path}')
    lens = [fa.length(nm) for nm in names]
    offs = np.concatenate(([0], np.cumsum([(ln + 63) // 64 * 64 for ln in lens]))).astype(np.int64)
    buf = None
    if pinned and ctx is not None and int(offs[-1]) <= _PINNED_STAGING_MAX:
        try:
            buf = device.pinned_empty(ctx, int(offs[-1]))
        except RuntimeError:   # no pinned memory to be had: ordinary memory works, only slower
            buf = None
    if buf is None:
        buf = np.empty(int(offs[-1]), dtype=np.uint8)
    out = [buf[offs[i]:offs[i] + lens[i]] for i in range(len(names))]
    # jobs of about equal size: small records are grouped, large ones are cut into runs of lines (a chromosome read by one thread
    # bounded the call at N > 1 GPUs, where a rank owns three chromosomes: 0.1 s for 250 Mbp through one strided copy)
    total = int(offs[-1])
    target = max(total // (4 * _READERS), 1 << 20)
    jobs, cur, acc = [], [], 0
    for i in range(len(names)):
        chunks = fa.line_chunks(names[i], target) if lens[i] > 2 * target else None
        if chunks is None:
            cur.append(i)
            acc += lens[i]
            if acc >= target:
                jobs.append(('whole', cur))
                cur, acc = [], 0
        else:
            jobs.extend(('lines', (i, l0, l1)) for l0, l1 in chunks)
    if cur:
        jobs.append(('whole', cur))

    def job(kind, arg):
        if kind == 'whole':
            for i in arg:
                fa.fetch_into(names[i], out[i])
        else:
            i, l0, l1 = arg
            fa.fetch_lines_into(names[i], out[i], l0, l1)
    return out, [pool.submit(job, kind, arg) for kind, arg in jobs]


def call_rows(df_align, ref_fa_name, tig_fa_name):
    """First half of ``make_insdel_snv_calls`` (``df_align`` not empty): read the sequences, run the walk on the GPU.

    Returns ``(table, snv rows, indel rows, ref_arr, tig_arr, phase seconds)``; ``ref_arr`` / ``tig_arr`` are indexed by
    ``table.ref_id`` / ``table.qry_id`` (with a sidecar ``ref_arr`` holds every sequence of the reference, in file order)."""
    from concurrent.futures import ThreadPoolExecutor
    global _CALLS
    t0 = time.perf_counter()
    table = AlignTable(df_align)
    ref_fa = fasta.open_fasta(ref_fa_name)
    tig_fa = fasta.open_fasta(tig_fa_name)
    sc = sidecar.find(ref_fa_name)    # packed-reference sidecar next to the FASTA (pav_b200/sidecar.py), when there is a fresh one
    ref_store, own_store = None, False
    # Sequences are read by a few threads (file reads, numpy copies and the C calls release the GIL) while the main thread
    # uploads what is ready; the CIGAR tokenizer runs beside them. From the second call of a process on, the bases are read
    # straight into pinned staging buffers of the context's pool (H2D at PCIe speed; the first call would only pay for pinning).
    ctx = device.get_context()
    pinned = _CALLS >= 1 if os.environ.get('PAVGPU_PINNED_STAGING') is None else os.environ['PAVGPU_PINNED_STAGING'] == '1'
    _CALLS += 1

    def read_all(fa, names):
        return read_sequences(fa, names, pool, ctx, pinned)

    with ThreadPoolExecutor(max_workers=2 * _READERS + 1) as pool:
        futs = []
        try:
            if sc is not None:
                # host arrays are views of the mapped file, the upload is the packed planes; records index the sidecar's order
                missing = [nm for nm in table.ref_names if nm not in sc.ids]
                if missing:
                    raise KeyError(f'sequence {missing[0]!r} not found in {sc.path}')
                ref_arr, ref_futs = sc.arrays(), []
                ref_id = np.array([sc.ids[nm] for nm in table.ref_names], dtype=np.int32)[table.ref_id]
                table.ref_names, table.ref_id = {nm: i for i, nm in enumerate(sc.names)}, ref_id
            else:
                ref_arr, ref_futs = read_all(ref_fa, table.ref_names)
            tig_arr, tig_futs = read_all(tig_fa, table.tig_names)
            f_ops = pool.submit(device.parse_cigars, table.cigars)
            futs = ref_futs + tig_futs + [f_ops]
            if sc is not None:
                ref_store, own_store = sidecar.reference_store(sc)
            for f in ref_futs:
                f.result()
            t1 = time.perf_counter()

            # walk_rows resolves its inputs lazily: the contigs are awaited after the reference is uploaded
            snv, indel = walk_rows(table, ref_arr, _Awaited(tig_futs, tig_arr), ctx=ctx, ref_store=ref_store, parsed=f_ops)
        except BaseException:
            for f in futs:
                f.cancel()
            raise
        finally:
            if own_store:
                ref_store.close()
    t2 = time.perf_counter()
    return table, snv, indel, ref_arr, tig_arr, {'fasta_reference': t1 - t0, 'device_walk_incl_contig_read_h2d_d2h': t2 - t1, 'sidecar': sc is not None}


_MALLOC_TUNED = False
_CALLS = 0        # make_insdel_snv_calls calls in this process (pinned staging starts with the second one)
_READERS = int(os.environ.get('PAVGPU_FASTA_READERS', str(min(12, max(3, len(os.sched_getaffinity(0)) // 2)))))   # reader threads per FASTA
#   (C3, one haplotype, B200 box with 16 cores: 3 readers + 1 GB pinned cap 1.63 s per call, 12 readers 1.52 s, 12 readers + pinned staging
#   for the whole 3.1 GB of each side 1.05 s -- profiles/r02_c3_e2e_sweep.log)
_PINNED_STAGING_MAX = int(os.environ.get('PAVGPU_PINNED_STAGING_MAX_MB', '4096')) << 20   # per FASTA side; larger inputs stay in ordinary memory


def _tune_malloc():
    """Large tables create millions of small objects and a few dozen multi-MB pointer arrays per call; by default every byte of
    them is freshly mapped memory (first-touch page faults) that goes straight back to the OS when the result dies:
      * glibc: keep freed heap (no trim), grow it in big steps, and serve arrays up to 32 MiB from the heap instead of mmap
        (measured: first call 2.0 s -> 1.1 s for 2 M rows before the C formatter existed);
      * pymalloc: recycle released arenas (``_pyrows.keep_arenas``) instead of unmapping and re-faulting them every call.
    These are process-wide settings that outlive the call and keep up to ``PAVGPU_ARENA_CACHE_MB`` (default 512) of freed arenas
    plus untrimmed heap resident, so they are OPT-IN: nothing is touched unless ``PAVGPU_TUNE_ALLOC=1`` is set (bench.py sets it
    for its end-to-end leg and says so in its JSON line). With it the first large result applies them."""
    global _MALLOC_TUNED
    if _MALLOC_TUNED:
        return
    _MALLOC_TUNED = True
    import os
    if os.environ.get('PAVGPU_TUNE_ALLOC', '0') != '1':
        return
    try:
        import ctypes
        libc = ctypes.CDLL('libc.so.6')
        libc.mallopt(-1, 2 ** 31 - 1)      # M_TRIM_THRESHOLD: keep freed heap
        libc.mallopt(-2, 256 << 20)        # M_TOP_PAD: grow the heap 256 MiB at a time
        libc.mallopt(-3, 32 << 20)         # M_MMAP_THRESHOLD: column arrays (8 B x rows) come from the heap and are reused
    except Exception:  # noqa: BLE001  (non-glibc platforms)
        pass
    try:
        from .. import _pyrows
        _pyrows.keep_arenas(int(os.environ.get('PAVGPU_ARENA_CACHE_MB', '512')))
    except Exception:  # noqa: BLE001
        pass


def _obj(values):
    """list -> object ndarray (one pass, no type inference)."""
    if isinstance(values, list):
        a = np.empty(len(values), dtype=object)
        a[:] = values
        return a
    return values


def _frame(cols, columns, index):
    n = len(index)
    data = {}
    for name in columns:
        v = cols[name]
        if isinstance(v, (list, np.ndarray)):
            data[name] = _obj(v)
        else:  # scalar column
            a = np.empty(n, dtype=object)
            a[:] = v
            data[name] = a
    # copy=False keeps one block per column (no 2-D consolidation copy); dtypes stay object
    return pd.DataFrame(data, columns=columns, index=pd.Index(index, dtype=np.int64), dtype=object, copy=False)


def build_frames(snv, indel, chrom, qry, rev, align_index, ref_arr, tig_arr, ref_id, qry_id, hap, version_id):
    """Rows from the device (``pavgpu_snv_row`` / ``pavgpu_indel_row`` arrays in emission order) -> the two
    DataFrames of the reference. Host-side string formatting only (pav_b200/csrc/pyrows.c); every coordinate
    comes from the GPU. IDs are formatted in emission order (version_id and the sort's tie-break need them),
    every other column directly in the final row order."""
    import gc
    if len(snv) + len(indel) > 100_000:
        _tune_malloc()
    gc_was_on = gc.isenabled()
    gc.disable()   # tens of millions of str / int objects are created below; none of them can form cycles
    try:
        return _build_frames(snv, indel, chrom, qry, rev, align_index, ref_arr, tig_arr, ref_id, qry_id, hap, version_id)
    finally:
        if gc_was_on:
            gc.enable()


def tables_tsv(snv, indel, chrom, qry, rev, align_index, ref_arr, tig_arr, ref_id, qry_id, hap, version_id, pass_snv=None, pass_indel=None):
    """The two call tables as TSV text (bytes, header included) exactly as ``DataFrame.to_csv(sep='\t', index=False)`` writes the
    frames of ``build_frames`` -- without building them (no per-row Python objects). ``pass_snv`` / ``pass_indel``: uint8 per
    emission row, 1 = PASS / 0 = TRIM, adds the FILTER column of ``rule call_cigar``. Returns ``None`` when a name needs CSV
    quoting or is not ASCII (the caller then goes through the frames)."""
    from .. import _pyrows
    R = _Records(chrom, qry, rev, align_index, ref_id, qry_id)
    ai_l = [f'{x}' for x in R.ai_objs]
    special = ('\t', '\n', '\r', '"')
    if not R.ascii_names or not hap.isascii() or any(ch in x for x in R.chrom_l + R.qry_l + ai_l + [hap] for ch in special) or \
            not all(x.isascii() for x in ai_l) or (len(indel) and int(indel['svlen'].min()) <= 0):
        return None
    seqs = list(ref_arr) + list(tig_arr)
    comp = fasta.COMPLEMENT.tobytes()
    out = []
    for rows, cols, passes, is_snv in ((snv, SNV_COLUMNS, pass_snv, True), (indel, INSDEL_COLUMNS, pass_indel, False)):
        header = ('\t'.join(cols + (['FILTER'] if passes is not None else [])) + '\n').encode()
        if len(rows) == 0:
            out.append(header)
            continue
        rows = np.ascontiguousarray(rows)
        if is_snv:
            ids, order, _ = _snv_ids_order(rows, R, ref_arr, tig_arr, version_id)
            fn = _pyrows.snv_tsv
        else:
            ids, order = _indel_ids_order(rows, R, version_id)
            fn = _pyrows.indel_tsv
        if ids is not None and any(ch in x for x in ids.tolist() for ch in special):
            return None
        out.append(fn(rows.view(np.uint8), order, ids, R.chrom_l, R.qry_l, R.strand_objs, ai_l, R.ref_id32, R.qry_id32, R.rev8, seqs, len(ref_arr), comp,
                      hap, CALL_SOURCE, None if passes is None else np.ascontiguousarray(passes, dtype=np.uint8), header))
    return out[0], out[1]


class _Records:
    """Per-record columns in the forms the formatters want (objects for frame cells, ASCII strings for text)."""

    def __init__(self, chrom, qry, rev, align_index, ref_id, qry_id):
        self.n_rec = len(chrom)
        self.chrom, self.align_index = chrom, align_index
        self.chrom_objs, self.ai_objs = chrom.tolist(), align_index.tolist()
        self.chrom_l = [f'{c}' for c in self.chrom_objs]
        self.qry_l = [f'{q}' for q in qry.tolist()]
        chrom_rank = {c: i for i, c in enumerate(sorted(set(self.chrom_l)))}
        self.chrom_code_rec = np.array([chrom_rank[c] for c in self.chrom_l], dtype=np.int64)
        self.strand_objs = ['-' if r else '+' for r in rev.tolist()]
        self.rev, self.ref_id, self.qry_id = rev, ref_id, qry_id
        self.ref_id32 = np.ascontiguousarray(ref_id, dtype=np.int32)
        self.qry_id32 = np.ascontiguousarray(qry_id, dtype=np.int32)
        self.rev8 = np.ascontiguousarray(rev, dtype=np.uint8)
        self.ascii_names = all(x.isascii() for x in self.chrom_l) and all(x.isascii() for x in self.qry_l)


def _any_repeat(*keys):
    """True when two rows agree on every key array (most significant first). Rows usually arrive sorted by the first key, so a
    linear pass decides; otherwise one lexsort."""
    n = len(keys[0])
    if n < 2:
        return False
    k0 = keys[0]
    if bool((k0[1:] > k0[:-1]).all()):     # strictly increasing first key: no two rows can agree
        return False
    order = np.lexsort(keys[::-1])
    same = np.ones(n - 1, dtype=bool)
    for k in keys:
        ks = k[order]
        same &= ks[1:] == ks[:-1]
    return bool(same.any())


def _snv_ids_order(snv, R, ref_arr, tig_arr, version_id):
    """``(ids in emission order or None, final row order, bases_of)`` for the SNV rows: IDs are only materialised when they are
    versioned (or names are not ASCII); otherwise the sort asks for the ID of a row only inside tie groups."""
    from .. import _pyrows
    n = len(snv)
    rec = snv['rec']
    pos = snv['pos_ref'].astype(np.int64)
    qp = snv['qry_pos']

    def bases_of(idx):
        """(REF, ALT) bytes of emission rows ``idx`` (whole table for ID versioning, tie groups of the sort otherwise)."""
        ref_b = np.empty(len(idx), dtype=np.uint8)
        alt_b = np.empty(len(idx), dtype=np.uint8)
        r_of, p_of, q_of = rec[idx], pos[idx], qp[idx]
        # one gather per run of rows of the same record (rows arrive in record order: as many runs as records)
        cut = np.flatnonzero(r_of[1:] != r_of[:-1]) + 1
        for a, b in zip([0] + cut.tolist(), cut.tolist() + [len(idx)]):
            if a == b:
                continue
            r = int(r_of[a])
            ref_b[a:b] = ref_arr[R.ref_id[r]][p_of[a:b]]
            t = tig_arr[R.qry_id[r]][q_of[a:b]]
            alt_b[a:b] = fasta.COMPLEMENT[t] if R.rev[r] else t
        return ref_b, alt_b

    ids = None
    if version_id or not R.ascii_names:
        ref_b, alt_b = bases_of(np.arange(n))
        ids = _pyrows.format(n, [('l', R.chrom_l, rec.astype(np.int64)), ('s', '-'), ('i', pos + 1), ('s', '-SNV-'),
                                 ('c', fasta.UPPER[ref_b]), ('c', fasta.UPPER[alt_b])])
        # an ID is '{chrom}-{pos+1}-SNV-{REF}{ALT}': two rows share one exactly when they agree on these four values, so the
        # string pass of version_id (a Counter over every ID) only runs when such rows exist
        if version_id and _any_repeat((R.chrom_code_rec[rec] << 40) | pos, (fasta.UPPER[ref_b].astype(np.int64) << 8) | fasta.UPPER[alt_b]):
            ids = np.ascontiguousarray(variant.version_id(pd.Series(ids, dtype=object)).to_numpy(dtype=object))
        id_of = ids.__getitem__
    else:
        def id_of(i):
            rb, ab = bases_of(np.array([i]))
            return f'{R.chrom_l[rec[i]]}-{pos[i] + 1}-SNV-{chr(fasta.UPPER[rb[0]])}{chr(fasta.UPPER[ab[0]])}'
    order = _sort_order(R.chrom_code_rec[rec], pos, pos + 1, id_of, end_is_pos_plus_1=True)
    return ids, order, bases_of


def _indel_ids_order(indel, R, version_id):
    """``(ids in emission order or None, final row order)`` for the INS / DEL rows."""
    from .. import _pyrows
    n = len(indel)
    rec = indel['rec']
    pos = indel['pos'].astype(np.int64)
    end = indel['end'].astype(np.int64)
    svlen = indel['svlen']
    svt = indel['svtype']
    ids = None
    if version_id or not R.ascii_names:
        ids = _pyrows.format(n, [('l', R.chrom_l, rec.astype(np.int64)), ('s', '-'), ('i', pos + 1), ('s', '-'),
                                 ('l', ['INS', 'DEL'], (svt == 1).astype(np.int64)), ('s', '-'), ('i', svlen.astype(np.int64))])
        if version_id and _any_repeat((R.chrom_code_rec[rec] << 40) | pos, ((svt == 1).astype(np.int64) << 40) | svlen.astype(np.int64)):
            ids = np.ascontiguousarray(variant.version_id(pd.Series(ids, dtype=object)).to_numpy(dtype=object))   # same argument as for SNVs
        id_of = ids.__getitem__
    else:
        def id_of(i):
            return f'{R.chrom_l[rec[i]]}-{pos[i] + 1}-{("INS", "DEL")[int(svt[i] == 1)]}-{svlen[i]}'
    return ids, _sort_order(R.chrom_code_rec[rec], pos, end, id_of)


def _build_frames(snv, indel, chrom, qry, rev, align_index, ref_arr, tig_arr, ref_id, qry_id, hap, version_id):
    from .. import _pyrows
    R = _Records(chrom, qry, rev, align_index, ref_id, qry_id)
    chrom_l, qry_l, chrom_objs, ai_objs, strand_objs = R.chrom_l, R.qry_l, R.chrom_objs, R.ai_objs, R.strand_objs
    ref_id32, qry_id32, rev8, ascii_names = R.ref_id32, R.qry_id32, R.rev8, R.ascii_names

    # ------------------------------------------------------------------ SNV rows (cigarcall.py:98-135)
    seqs = list(ref_arr) + list(tig_arr)
    comp = fasta.COMPLEMENT.tobytes()
    if len(snv):
        n = len(snv)
        snv = np.ascontiguousarray(snv)
        ids, order, bases_of = _snv_ids_order(snv, R, ref_arr, tig_arr, version_id)
        if ascii_names:
            cols = _pyrows.snv_frame(snv.view(np.uint8), order, ids, chrom_objs, chrom_l, qry_l, strand_objs, ai_objs, ref_id32, qry_id32, rev8,
                                     seqs, len(ref_arr), comp, ('SNV', 1, hap, 0, CALL_SOURCE))
            df_snv = _frame(dict(zip(SNV_COLUMNS, cols)), SNV_COLUMNS, order)
        else:   # names outside ASCII: generic formatter
            rec, pos, qp = snv['rec'], snv['pos_ref'].astype(np.int64), snv['qry_pos']
            ref_b, alt_b = bases_of(np.arange(n))
            rec64, pos_o, qp1 = rec[order].astype(np.int64), pos[order], qp[order].astype(np.int64) + 1
            cols = {
                '#CHROM': chrom[rec64], 'POS': _pyrows.ints(pos_o), 'END': _pyrows.ints(pos_o + 1), 'ID': ids[order],
                'SVTYPE': 'SNV', 'SVLEN': 1, 'REF': _CHR[ref_b[order]], 'ALT': _CHR[alt_b[order]], 'HAP': hap,
                'QRY_REGION': _pyrows.format(n, [('l', qry_l, rec64), ('s', ':'), ('i', qp1), ('s', '-'), ('i', qp1)]),
                'QRY_STRAND': np.array(strand_objs, dtype=object)[rec64], 'CI': 0, 'ALIGN_INDEX': align_index[rec64],
                'CALL_SOURCE': CALL_SOURCE,
            }
            df_snv = _frame(cols, SNV_COLUMNS, order)
    else:
        df_snv = _empty(SNV_COLUMNS)

    # ------------------------------------------------------------------ INS / DEL rows (cigarcall.py:141-282)
    if len(indel):
        n = len(indel)
        indel = np.ascontiguousarray(indel)
        rec = indel['rec']
        pos = indel['pos'].astype(np.int64)
        end = indel['end'].astype(np.int64)
        svlen = indel['svlen']
        svtype_l = ['INS', 'DEL']
        svt = indel['svtype']
        ids, order = _indel_ids_order(indel, R, version_id)
        if ascii_names:
            cols = _pyrows.indel_frame(indel.view(np.uint8), order, ids, chrom_objs, chrom_l, qry_l, strand_objs, ai_objs, ref_id32, qry_id32, rev8,
                                       seqs, len(ref_arr), comp, ('INS', 'DEL', hap, 0, CALL_SOURCE))
            df_insdel = _frame(dict(zip(INSDEL_COLUMNS, cols)), INSDEL_COLUMNS, order)
        else:
            indel_o = indel[order]
            rec64, pos_o, end_o = rec[order].astype(np.int64), pos[order], end[order]
            svlen_o, is_del = svlen[order].astype(np.int64), svt[order] == 1
            qp = indel_o['qry_pos'].astype(np.int64)
            qe = indel_o['qry_end'].astype(np.int64)
            n_ref = len(ref_arr)
            which = np.where(is_del, ref_id[rec64].astype(np.int64), n_ref + qry_id[rec64].astype(np.int64))
            start = np.where(is_del, pos_o, qp)
            rc = (~is_del & rev[rec64]).astype(np.uint8)
            cols = {
                '#CHROM': chrom[rec64], 'POS': _pyrows.ints(pos_o), 'END': _pyrows.ints(end_o), 'ID': ids[order],
                'SVTYPE': np.array(svtype_l, dtype=object)[is_del.astype(np.int64)], 'SVLEN': _pyrows.ints(svlen_o), 'HAP': hap,
                'QRY_REGION': _pyrows.format(n, [('l', qry_l, rec64), ('s', ':'), ('i', qp + 1), ('s', '-'), ('i', np.where(is_del, qp + 1, qe))]),
                'QRY_STRAND': np.array(strand_objs, dtype=object)[rec64], 'CI': 0, 'ALIGN_INDEX': align_index[rec64],
                'LEFT_SHIFT': _pyrows.ints(indel_o['left_shift'].astype(np.int64)),
                'HOM_REF': _pyrows.format(n, [('i', indel_o['hom_ref_l'].astype(np.int64)), ('s', ','), ('i', indel_o['hom_ref_r'].astype(np.int64))]),
                'HOM_TIG': _pyrows.format(n, [('i', indel_o['hom_tig_l'].astype(np.int64)), ('s', ','), ('i', indel_o['hom_tig_r'].astype(np.int64))]),
                'CALL_SOURCE': CALL_SOURCE,
                'SEQ': _pyrows.slices(list(ref_arr) + list(tig_arr), which, start, svlen_o, rc, fasta.COMPLEMENT.tobytes()),
            }
            df_insdel = _frame(cols, INSDEL_COLUMNS, order)
    else:
        df_insdel = _empty(INSDEL_COLUMNS)

    return df_snv, df_insdel
