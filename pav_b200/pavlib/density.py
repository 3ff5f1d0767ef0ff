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
        d = {'status': int(r['status']), 'smoothed': bool(r['smoothed']), 'n_eval': int(r['n_eval'])}
        for k, v in cols.items():
            d[k] = v[a:b]
        out.append(d)
    return out


def density_windows(windows, k=31, ctx=None, **kw):
    """Score in-memory windows: ``windows`` = iterable of ``(ref uint8 array, tig uint8 array, rev, srs)``.

    Returns one dict per window: ``status`` (0 / 125), ``smoothed``, ``n_eval`` and the column arrays.
    """
    global last_stats
    ctx = ctx or device.get_context()
    windows = list(windows)
    if len(windows) > MAX_WINDOWS_PER_BATCH:   # bound the device arena and the pinned result buffers (~5 MB + ~2 MB per 50 kbp window)
        out = []
        for a in range(0, len(windows), MAX_WINDOWS_PER_BATCH):
            out.extend(density_windows(windows[a:a + MAX_WINDOWS_PER_BATCH], k=k, ctx=ctx, **kw))
        return out
    refs = [np.ascontiguousarray(w[0], dtype=np.uint8) for w in windows]
    tigs = [np.ascontiguousarray(w[1], dtype=np.uint8) for w in windows]
    import time
    t0 = time.perf_counter()
    rs = device.SeqStore(ctx, [f'r{i}' for i in range(len(refs))], refs, keep_host=False)
    ts = device.SeqStore(ctx, [f't{i}' for i in range(len(tigs))], tigs, keep_host=False)
    t1 = time.perf_counter()
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
            res, cols = batch.fetch()
            t4 = time.perf_counter()
        finally:
            batch.close()
    finally:
        rs.close()
        ts.close()
    out = _split(res, cols)
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
    # fewer than --mininf informative k-mers: frame returned before smoothing (density.py:193-194)
    return pd.DataFrame({
        'KMER': d['KMER'].astype(np.int64), 'INDEX': d['INDEX'].astype(np.int64), 'STATE': d['STATE'].astype(np.int64),
        'STATE_MER': d['STATE_MER'].astype(np.int64),
    }, columns=RAW_COLUMNS)


class DensityTable:
    """Result of one window whose DataFrame is built only when somebody needs it: the scan driver decides from the run lengths of
    STATE alone whether to expand the locus again, and most tables are dropped at that point."""

    def __init__(self, res):
        self.res = res
        self._df = None

    @property
    def shape(self):
        return (len(self.res['INDEX']), len(SMOOTHED_COLUMNS) if self.res['smoothed'] else len(RAW_COLUMNS))

    def rl(self):
        """``rl_encoder(frame)`` straight from the column arrays."""
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
