"""K-mer orientation density on the GPU with the reference's table layout.

Mirrors what ``scripts/density.py`` produces (scripts/density.py:341-342 smoothed frame, :193-194 raw
frame, exit code 125 for soft failures :510-527) and ``pavlib.density.rl_encoder``
(pavlib/density.py:330-361). All k-mer, state and density work happens in ``pavgpu_density_batch_*``
(pav_b200/csrc/density.cu); this module moves windows in and shapes DataFrames out.
"""
import ctypes

import numpy as np
import pandas as pd

from .. import _capi, device, fasta
from .constants import ERR_INV_FAIL

# K-mer orientation matrix [in fwd set][in rev set] (scripts/density.py:38-43)
KMER_ORIENTATION_STATE = np.asarray([[-1, 2], [0, 1]])

SMOOTHED_COLUMNS = ['INDEX', 'STATE_MER', 'STATE', 'KERN_FWD', 'KERN_FWDREV', 'KERN_REV', 'KMER']
RAW_COLUMNS = ['KMER', 'INDEX', 'STATE', 'STATE_MER']

last_stats = None
MAX_WINDOWS_PER_BATCH = 512   # windows scored per device batch by density_windows (larger requests are split)


def default_params(k=31, min_informative=2000, min_state_count=20, smooth=1.0, delta=0.005, max_ref_kmer_count=100):
    p = _capi.DensityParams()
    _capi.lib().pavgpu_density_default_params(ctypes.byref(p))
    p.k, p.min_informative, p.min_state_count = int(k), int(min_informative), int(min_state_count)
    p.smooth, p.delta, p.max_ref_kmer_count = float(smooth), float(delta), int(max_ref_kmer_count)
    return p


class DensityBatch:
    """Windows resident in HBM; ``run`` keeps results on the device, ``fetch`` copies them out."""

    def __init__(self, ctx, windows, params):
        self.ctx = ctx
        self.windows = np.ascontiguousarray(windows, dtype=_capi.DENSITY_WINDOW)
        h = _capi.c_vp()
        _capi.check(_capi.lib().pavgpu_density_batch_create(ctx.handle, len(self.windows), _capi.ptr(self.windows), ctypes.byref(params),
                                                            ctypes.byref(h)), 'pavgpu_density_batch_create')
        self.handle = h

    def run(self, ref_store, tig_store):
        st = _capi.DensityStats()
        _capi.check(_capi.lib().pavgpu_density_batch_run(self.handle, ref_store.handle, tig_store.handle, ctypes.byref(st)),
                    'pavgpu_density_batch_run')
        return st

    def fetch(self):
        L = _capi.lib()
        n_win = len(self.windows)
        res = np.zeros(n_win, dtype=_capi.DENSITY_RESULT)
        ptrs = [_capi.c_vp() for _ in range(7)]
        n_rows = _capi.c_i64()
        _capi.check(L.pavgpu_density_batch_fetch(self.handle, _capi.ptr(res), *[ctypes.byref(p) for p in ptrs], ctypes.byref(n_rows)),
                    'pavgpu_density_batch_fetch')
        n = n_rows.value
        dts = [np.uint64, np.int32, np.int8, np.int8, np.float64, np.float64, np.float64]
        cols = [_capi.take_host_array(p.value, n, dt) for p, dt in zip(ptrs, dts)]
        return res, dict(zip(['KMER', 'INDEX', 'STATE_MER', 'STATE', 'KERN_FWD', 'KERN_FWDREV', 'KERN_REV'], cols))

    def fetch_runs(self):
        """Run lengths of STATE per window (``pavgpu_density_batch_fetch_runs``): ``(res, runs, run_off)``."""
        n_win = len(self.windows)
        res = np.zeros(n_win, dtype=_capi.DENSITY_RESULT)
        run_off = np.zeros(n_win + 1, dtype=np.int64)
        p, n = _capi.c_vp(), _capi.c_i64()
        _capi.check(_capi.lib().pavgpu_density_batch_fetch_runs(self.handle, _capi.ptr(res), ctypes.byref(p), _capi.ptr(run_off), ctypes.byref(n)),
                    'pavgpu_density_batch_fetch_runs')
        return res, _capi.take_host_array(p.value, n.value, _capi.STATE_RUN), run_off

    def fetch_window(self, win, n_rows):
        """All columns of one window (``pavgpu_density_batch_fetch_window``)."""
        dts = [np.uint64, np.int32, np.int8, np.int8, np.float64, np.float64, np.float64]
        cols = [np.empty(n_rows, dtype=dt) for dt in dts]
        _capi.check(_capi.lib().pavgpu_density_batch_fetch_window(self.handle, int(win), *[_capi.ptr(c) for c in cols]), 'pavgpu_density_batch_fetch_window')
        return dict(zip(['KMER', 'INDEX', 'STATE_MER', 'STATE', 'KERN_FWD', 'KERN_FWDREV', 'KERN_REV'], cols))

    def close(self):
        if getattr(self, 'handle', None):
            _capi.lib().pavgpu_density_batch_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def _split(res, cols):
    out = []
    for r in res:
        a, b = int(r['row_off']), int(r['row_off'] + r['n_rows'])
        d = {'status': int(r['status']), 'smoothed': bool(r['smoothed']), 'n_eval': int(r['n_eval']), 'n_rows': int(r['n_rows'])}
        for k, v in cols.items():
            d[k] = v[a:b]
        out.append(d)
    return out


class _OpenBatch:
    """A scored batch kept on the device for as long as one of its ``LazyWindow`` results is alive."""

    def __init__(self, batch, stores):
        self.batch, self.stores = batch, stores

    def __del__(self):
        try:
            self.batch.close()
            for s in self.stores:
                s.close()
        except Exception:  # noqa: BLE001
            pass


class LazyWindow(dict):
    """Result of one window scored with ``density_windows(..., lazy=True)``: status, row count and the run lengths of STATE come
    back with the batch (16 bytes per run); the column arrays -- 38 bytes per row -- stay in HBM and are copied the first time one
    of them is asked for, which the inversion scan does only for the window that becomes a call."""

    _COLS = ('KMER', 'INDEX', 'STATE_MER', 'STATE', 'KERN_FWD', 'KERN_FWDREV', 'KERN_REV')

    def __init__(self, base, owner, win):
        super().__init__(base)
        self._owner, self._win = owner, win

    def __missing__(self, key):
        if key not in self._COLS:
            raise KeyError(key)
        self.update(self._owner.batch.fetch_window(self._win, self['n_rows']))
        self._owner = None
        return dict.__getitem__(self, key)


def density_windows(windows, k=31, ctx=None, lazy=False, **kw):
    """Score in-memory windows: ``windows`` = iterable of ``(ref uint8 array, tig uint8 array, rev, srs)``.

    Returns one dict per window: ``status`` (0 / 125), ``smoothed``, ``n_eval``, ``n_rows``, ``runs`` (``rl_encoder`` tuples of STATE
    as a structured array) and the column arrays. ``lazy=True``: the columns are fetched per window on first access (``LazyWindow``).
    """
    global last_stats
    ctx = ctx or device.get_context()
    windows = list(windows)
    if len(windows) > MAX_WINDOWS_PER_BATCH:   # bound the device arena and the pinned result buffers (~5 MB + ~2 MB per 50 kbp window)
        out, total = [], {}
        for a in range(0, len(windows), MAX_WINDOWS_PER_BATCH):
            out.extend(density_windows(windows[a:a + MAX_WINDOWS_PER_BATCH], k=k, ctx=ctx, lazy=lazy, **kw))
            for name, sec in last_stats['seconds'].items():
                total[name] = total.get(name, 0.0) + sec
        last_stats['seconds_all_batches'] = total      # (the other entries describe the last batch)
        return out
    refs = [np.ascontiguousarray(w[0], dtype=np.uint8) for w in windows]
    tigs = [np.ascontiguousarray(w[1], dtype=np.uint8) for w in windows]
    import time
    t0 = time.perf_counter()
    rs = device.SeqStore(ctx, [f'r{i}' for i in range(len(refs))], refs, keep_host=False)
    ts = device.SeqStore(ctx, [f't{i}' for i in range(len(tigs))], tigs, keep_host=False)
    t1 = time.perf_counter()
    keep = False
    try:
        win = np.zeros(len(windows), dtype=_capi.DENSITY_WINDOW)
        win['ref_seq_id'] = win['tig_seq_id'] = np.arange(len(windows))
        win['ref_end'] = [len(r) for r in refs]
        win['tig_end'] = [len(t) for t in tigs]
        win['rev'] = [int(bool(w[2])) for w in windows]
        win['srs'] = [int(w[3]) for w in windows]
        batch = DensityBatch(ctx, win, default_params(k=k, **kw))
        t2 = time.perf_counter()
        try:
            st = batch.run(rs, ts)
            t3 = time.perf_counter()
            res, runs, run_off = batch.fetch_runs()
            if lazy:
                owner = _OpenBatch(batch, (rs, ts))
                keep = True
                out = []
                for i, r in enumerate(res):
                    base = {'status': int(r['status']), 'smoothed': bool(r['smoothed']), 'n_eval': int(r['n_eval']), 'n_rows': int(r['n_rows']),
                            'runs': runs[run_off[i]:run_off[i + 1]], '_tig': tigs[i], '_k': int(k)}
                    out.append(LazyWindow(base, owner, i))
            else:
                res, cols = batch.fetch()
                out = _split(res, cols)
                for i, d in enumerate(out):
                    d['runs'] = runs[run_off[i]:run_off[i + 1]]
                    d['_tig'], d['_k'] = tigs[i], int(k)
            t4 = time.perf_counter()
        finally:
            if not keep:
                batch.close()
    finally:
        if not keep:
            rs.close()
            ts.close()
    last_stats = st.as_dict()
    last_stats['seconds'] = {'stores': t1 - t0, 'batch_create': t2 - t1, 'run': t3 - t2, 'fetch': t4 - t3, 'split': time.perf_counter() - t4}
    return out


def frame_from_result(d, extra=None):
    """Column arrays of one window -> the DataFrame scripts/density.py would have pickled. ``extra``: further columns (name ->
    array) appended in the same construction (a column added to a finished frame costs about as much as the frame)."""
    if d['smoothed']:
        index = d['INDEX'].astype(np.int64)
        cols = {
            'INDEX': index, 'STATE_MER': d['STATE_MER'].astype(np.int64), 'STATE': d['STATE'].astype(np.int64),
            'KERN_FWD': d['KERN_FWD'], 'KERN_FWDREV': d['KERN_FWDREV'], 'KERN_REV': d['KERN_REV'], 'KMER': d['KMER'].astype(np.int64),
        }
        ix = pd.Index(index, name='INDEX')
        for name, arr in (extra or {}).items():     # object columns stay object (a bare object array would be inferred as str dtype)
            cols[name] = pd.Series(arr, index=ix, dtype=object)
        return pd.DataFrame(cols, columns=SMOOTHED_COLUMNS + list(extra or ()), index=ix)
    assert not extra
    # fewer than --mininf informative k-mers: frame returned before smoothing (density.py:193-194). It is the k-mer stream's frame
    # with rows filtered out (:161-190), so its row labels are the positions of the kept k-mers in the stream of valid k-mers
    index = None
    if d.get('_tig') is not None:
        index = pd.Index(stream_ordinals(d['_tig'], d['_k'], d['INDEX']))
    return pd.DataFrame({
        'KMER': d['KMER'].astype(np.int64), 'INDEX': d['INDEX'].astype(np.int64), 'STATE': d['STATE'].astype(np.int64),
        'STATE_MER': d['STATE_MER'].astype(np.int64),
    }, columns=RAW_COLUMNS, index=index)


_IS_ACGT = np.zeros(256, dtype=bool)
_IS_ACGT[list(b'ACGTacgt')] = True


def stream_ordinals(tig, k, index):
    """Position of the k-mers starting at window offsets ``index`` in the stream of valid k-mers of ``tig`` (kanapy's ``stream``
    emits one k-mer per offset whose k bases are all ACGTacgt, kmer.py:206-221): the row labels of the reference's raw frame."""
    bad = np.concatenate(([0], np.cumsum(~_IS_ACGT[np.asarray(tig, dtype=np.uint8)])))
    n_pos = len(tig) - k + 1
    if n_pos <= 0:
        return np.zeros(0, dtype=np.int64)
    valid = (bad[k:k + n_pos] - bad[:n_pos]) == 0
    return (np.cumsum(valid) - 1)[np.asarray(index, dtype=np.int64)].astype(np.int64)


class DensityTable:
    """Result of one window whose DataFrame is built only when somebody needs it: the scan driver decides from the run lengths of
    STATE alone whether to expand the locus again, and most tables are dropped at that point."""

    def __init__(self, res):
        self.res = res
        self._df = None

    @property
    def shape(self):
        n = self.res['n_rows'] if 'n_rows' in self.res else len(self.res['INDEX'])
        return (n, len(SMOOTHED_COLUMNS) if self.res['smoothed'] else len(RAW_COLUMNS))

    def rl(self):
        """``rl_encoder(frame)``: from the run lengths the device computed, else straight from the column arrays."""
        runs = self.res.get('runs')
        if runs is not None:
            for r in runs.tolist():
                yield r
            return
        ix = self.res['INDEX']
        st = self.res['STATE'] if self.res['smoothed'] else np.full(len(ix), -1, dtype=np.int8)
        n = len(st)
        if n == 0:
            return
        brk = np.flatnonzero(st[1:] != st[:-1]) + 1
        starts = np.concatenate(([0], brk))
        ends = np.concatenate((brk, [n]))
        for a, b in zip(starts.tolist(), ends.tolist()):
            yield (st[a].item(), b - a, ix[a].item(), ix[b - 1].item())

    def frame(self, extra=None):
        if extra:
            return frame_from_result(self.res, extra)
        if self._df is None:
            self._df = frame_from_result(self.res)
        return self._df


def density_table(region_ref, region_tig, ref_fa_name, tig_fa_name, k=31, rev=False, state_run_smooth=20, **kw):
    """One window addressed like ``scripts/density.py --refregion/--tigregion/--ref/--tig -k -r --staterunsmooth``.

    Returns ``(returncode, DataFrame or None)``: 0 with the table, or ``ERR_INV_FAIL`` (125) with ``None``.
    The reference window is always read on the forward strand and the contig window in forward contig
    coordinates (scripts/density.py:502,543; pavlib/seq.py:316).
    """
    ref = fasta.open_fasta(ref_fa_name).fetch_array(region_ref.chrom, region_ref.pos, region_ref.end)
    tig = fasta.open_fasta(tig_fa_name).fetch_array(region_tig.chrom, region_tig.pos, region_tig.end)
    res = density_windows([(ref, tig, rev, state_run_smooth)], k=k, **kw)[0]
    if res['status'] != 0:
        return ERR_INV_FAIL, None
    return 0, frame_from_result(res)


def rl_encoder(df, state_col='STATE'):
    """Run-length encode a state column: yields ``(state, count, first INDEX, last INDEX)`` (pavlib/density.py:330-361)."""
    st = df[state_col].to_numpy()
    ix = df['INDEX'].to_numpy()
    n = len(st)
    if n == 0:
        return
    brk = np.flatnonzero(st[1:] != st[:-1]) + 1
    starts = np.concatenate(([0], brk))
    ends = np.concatenate((brk, [n]))
    for a, b in zip(starts.tolist(), ends.tolist()):
        yield (st[a].item(), b - a, ix[a].item(), ix[b - 1].item())
