"""Thin object layer over the C ABI: device contexts, sequence stores, the CIGAR-walk call.

No torch, no CPU fallback: every function here ends in a libpavgpu.so call on a CUDA device.
Device selection follows SURVEY 8(b): ``PAVGPU_DEVICES`` (comma list, default all visible) and
``PAVGPU_DEVICE_INDEX`` (which entry this process uses, default 0) so concurrent Snakemake jobs can
be spread over the GPUs of a box without touching the rule files.
"""
import ctypes
import os

import numpy as np

from . import _capi
from ._capi import c_i64, c_vp

_CONTEXTS = {}


class Context:
    def __init__(self, device=0):
        L = _capi.lib()
        n = L.pavgpu_device_count()
        if n <= 0:
            raise RuntimeError('pav_b200: no usable CUDA device (pavgpu_device_count() = %d: %s); '
                               'the hot path has no CPU fallback' % (n, L.pavgpu_last_error().decode()))
        if not 0 <= device < n:
            raise RuntimeError(f'pav_b200: device {device} out of range (0..{n - 1})')
        h = c_vp()
        _capi.check(L.pavgpu_ctx_create(device, ctypes.byref(h)), 'pavgpu_ctx_create')
        self.handle = h
        self.device = device

    def l2_flush(self, nbytes=256 << 20):
        _capi.check(_capi.lib().pavgpu_l2_flush(self.handle, nbytes), 'pavgpu_l2_flush')

    def close(self):
        if self.handle:
            _capi.lib().pavgpu_ctx_destroy(self.handle)
            self.handle = None


def default_device():
    devs = os.environ.get('PAVGPU_DEVICES')
    idx = int(os.environ.get('PAVGPU_DEVICE_INDEX', os.environ.get('LOCAL_RANK', '0')))
    if devs:
        lst = [int(x) for x in devs.split(',') if x.strip() != '']
        return lst[idx % len(lst)]
    n = _capi.lib().pavgpu_device_count()
    return idx % n if n > 0 else 0


def get_context(device=None):
    if device is None:
        device = default_device()
    ctx = _CONTEXTS.get(device)
    if ctx is None:
        ctx = _CONTEXTS[device] = Context(device)
    return ctx


class SeqStore:
    """Named sequences packed into HBM (2-bit plane + N-mask plane), plus the host bytes for slicing."""

    def __init__(self, ctx, names, arrays, keep_host=True):
        L = _capi.lib()
        self.ctx = ctx
        self.names = [str(n) for n in names]
        self.ids = {n: i for i, n in enumerate(self.names)}
        arrays = [np.ascontiguousarray(a, dtype=np.uint8) for a in arrays]
        self.lengths = np.array([len(a) for a in arrays], dtype=np.int64)
        n = len(arrays)
        ptrs = (c_vp * max(n, 1))(*[a.ctypes.data for a in arrays])
        h = c_vp()
        _capi.check(L.pavgpu_seqstore_create(ctx.handle, n, ptrs, self.lengths.ctypes.data_as(ctypes.POINTER(c_i64)),
                                             ctypes.byref(h)), 'pavgpu_seqstore_create')
        self.handle = h
        self.host = arrays if keep_host else None

    @classmethod
    def from_packed(cls, ctx, names, lengths, pack2, nmask):
        L = _capi.lib()
        self = cls.__new__(cls)
        self.ctx = ctx
        self.names = [str(n) for n in names]
        self.ids = {n: i for i, n in enumerate(self.names)}
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        h = c_vp()
        if pack2 is None:
            _capi.check(L.pavgpu_seqstore_create_empty(ctx.handle, len(self.names), self.lengths.ctypes.data_as(ctypes.POINTER(c_i64)),
                                                       ctypes.byref(h)), 'pavgpu_seqstore_create_empty')
        else:
            pack2, nmask = np.ascontiguousarray(pack2), np.ascontiguousarray(nmask)
            _capi.check(L.pavgpu_seqstore_create_packed(ctx.handle, len(self.names), self.lengths.ctypes.data_as(ctypes.POINTER(c_i64)),
                                                        _capi.ptr(pack2), pack2.nbytes, _capi.ptr(nmask), nmask.nbytes, ctypes.byref(h)),
                        'pavgpu_seqstore_create_packed')
        self.handle = h
        self.host = None
        return self

    def plane_sizes(self):
        L = _capi.lib()
        p2, pm = c_vp(), c_vp()
        b2, bm = ctypes.c_size_t(), ctypes.c_size_t()
        _capi.check(L.pavgpu_seqstore_planes(self.handle, ctypes.byref(p2), ctypes.byref(b2), ctypes.byref(pm), ctypes.byref(bm)))
        return p2.value, b2.value, pm.value, bm.value

    def export(self):
        _, b2, _, bm = self.plane_sizes()
        pack2 = np.empty(b2 // 8, dtype=np.uint64)
        nmask = np.empty(bm // 4, dtype=np.uint32)
        _capi.check(_capi.lib().pavgpu_seqstore_export(self.handle, _capi.ptr(pack2), _capi.ptr(nmask)))
        return pack2, nmask

    def checksum(self):
        """(2-bit plane, mask plane) checksums computed on the device (``pavgpu_seqstore_checksum``)."""
        out = np.zeros(2, dtype=np.uint64)
        _capi.check(_capi.lib().pavgpu_seqstore_checksum(self.handle, _capi.ptr(out)), 'pavgpu_seqstore_checksum')
        return int(out[0]), int(out[1])

    def offset(self, i):
        return int(_capi.lib().pavgpu_seqstore_offset(self.handle, i))

    def broadcast(self, unique_id, rank, n_ranks):
        """NCCL broadcast of both planes from rank 0. Returns device milliseconds. ``unique_id`` None: reuse the
        communicator an earlier call of this process made for (device, rank, n_ranks) -- see ``nccl_comm_cached``."""
        ms = ctypes.c_float()
        buf = None if unique_id is None else np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        _capi.check(_capi.lib().pavgpu_seqstore_broadcast(self.ctx.handle, self.handle, None if buf is None else _capi.ptr(buf), rank, n_ranks,
                                                          ctypes.byref(ms)), 'pavgpu_seqstore_broadcast')
        return ms.value

    def close(self):
        if getattr(self, 'handle', None):
            _capi.lib().pavgpu_seqstore_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def pinned_empty(ctx, nbytes):
    """uint8 array over a pinned host buffer from the context's pool (goes back to the pool when the array and its views die)."""
    p = c_vp()
    _capi.check(_capi.lib().pavgpu_host_alloc(ctx.handle, int(max(nbytes, 1)), ctypes.byref(p)), 'pavgpu_host_alloc')
    return _capi.take_host_array(p.value, int(max(nbytes, 1)), np.uint8)[:nbytes]


def nccl_comm_cached(ctx, rank, n_ranks):
    """True when this process already holds an NCCL communicator for (ctx's device, rank, n_ranks)."""
    return bool(_capi.lib().pavgpu_nccl_comm_cached(ctx.handle, rank, n_ranks))


def nccl_unique_id():
    buf = np.zeros(128, dtype=np.uint8)
    _capi.check(_capi.lib().pavgpu_nccl_unique_id(_capi.ptr(buf)), 'pavgpu_nccl_unique_id')
    return buf.tobytes()


class LiftIndex:
    """Device index for batched coordinate lifts (``pavgpu_lift_index_*``): per-op first reference / contig coordinates of every
    record. ``bad_rec``: first record with an op the lift does not handle, or -1."""

    def __init__(self, ctx, ops, op_off, pos, rev, qry_len):
        self.ctx = ctx
        ops = np.ascontiguousarray(ops, dtype=np.uint32)
        op_off = np.ascontiguousarray(op_off, dtype=np.int64)
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        rev = np.ascontiguousarray(rev, dtype=np.uint8)
        qry_len = np.ascontiguousarray(qry_len, dtype=np.int64)
        self.n_rec = len(pos)
        h, bad = c_vp(), _capi.c_i32(-1)
        _capi.check(_capi.lib().pavgpu_lift_index_create(ctx.handle, _capi.ptr(ops) if len(ops) else None, _capi.ptr(op_off), self.n_rec, _capi.ptr(pos),
                                                         _capi.ptr(rev), _capi.ptr(qry_len), ctypes.byref(h), ctypes.byref(bad)), 'pavgpu_lift_index_create')
        self.handle, self.bad_rec = h, int(bad.value)

    def lift(self, rec, coord, to_qry):
        """-> (lifted int64 array, status int32 array: 0 lifted, 1 no block)."""
        rec = np.ascontiguousarray(rec, dtype=np.int32)
        coord = np.ascontiguousarray(coord, dtype=np.int64)
        out, status = np.zeros(len(rec), dtype=np.int64), np.zeros(len(rec), dtype=np.int32)
        if len(rec):
            _capi.check(_capi.lib().pavgpu_lift_points(self.handle, len(rec), _capi.ptr(rec), _capi.ptr(coord), int(bool(to_qry)), _capi.ptr(out), _capi.ptr(status)),
                        'pavgpu_lift_points')
        return out, status

    def close(self):
        if getattr(self, 'handle', None):
            _capi.lib().pavgpu_lift_index_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def cigar_record_stats(ops, op_off, ctx=None):
    """Per-record CIGAR summary on the device (``pavgpu_cigar_record_stats``): structured array ``_capi.CIGAR_REC_STATS``, one
    entry per record of ``parse_cigars``' output."""
    ctx = ctx or get_context()
    n_rec = len(op_off) - 1
    out = np.zeros(n_rec, dtype=_capi.CIGAR_REC_STATS)
    ops = np.ascontiguousarray(ops, dtype=np.uint32)
    op_off = np.ascontiguousarray(op_off, dtype=np.int64)
    _capi.check(_capi.lib().pavgpu_cigar_record_stats(ctx.handle, _capi.ptr(ops) if len(ops) else None, _capi.ptr(op_off), n_rec, _capi.ptr(out)),
                'pavgpu_cigar_record_stats')
    return out


def parse_cigars(cigars):
    """Tokenise CIGAR strings on the host (C): -> (ops uint32, op_off int64[n+1], ParseErr)."""
    L = _capi.lib()
    n = len(cigars)
    enc = [c.encode('ascii') if isinstance(c, str) else bytes(c) for c in cigars]
    text_off = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum([len(e) for e in enc], out=text_off[1:])
    blob = b''.join(enc)
    ops_p = c_vp()
    op_off = np.zeros(n + 1, dtype=np.int64)
    err = _capi.ParseErr()
    _capi.check(L.pavgpu_cigar_parse(blob, text_off.ctypes.data_as(ctypes.POINTER(c_i64)), n, ctypes.byref(ops_p),
                                     op_off.ctypes.data_as(ctypes.POINTER(c_i64)), ctypes.byref(err)), 'pavgpu_cigar_parse')
    ops = _capi.take_host_array(ops_p.value, int(op_off[n]), np.uint32)
    return ops, op_off, err


class CigarBatch:
    """Alignment records resident in HBM (what bench.py's device-resident `value` leg times)."""

    def __init__(self, ctx, ref_id, qry_id, pos, rev, ops, op_off):
        L = _capi.lib()
        self.ctx = ctx
        self.n_rec = len(ref_id)
        a = [np.ascontiguousarray(ref_id, np.int32), np.ascontiguousarray(qry_id, np.int32), np.ascontiguousarray(pos, np.int32),
             np.ascontiguousarray(rev, np.uint8), np.ascontiguousarray(ops, np.uint32), np.ascontiguousarray(op_off, np.int64)]
        h = c_vp()
        _capi.check(L.pavgpu_cigar_batch_create(ctx.handle, self.n_rec, *[_capi.ptr(x) for x in a], ctypes.byref(h)),
                    'pavgpu_cigar_batch_create')
        self.handle = h

    def run(self, ref_store, qry_store):
        st = _capi.CigarStats()
        _capi.check(_capi.lib().pavgpu_cigar_batch_run(self.handle, ref_store.handle, qry_store.handle, ctypes.byref(st)),
                    'pavgpu_cigar_batch_run')
        return st

    def fetch(self):
        L = _capi.lib()
        ps, pi = c_vp(), c_vp()
        ns, ni = c_i64(), c_i64()
        err = _capi.CigarErr()
        _capi.check(L.pavgpu_cigar_batch_fetch(self.handle, ctypes.byref(ps), ctypes.byref(ns), ctypes.byref(pi), ctypes.byref(ni),
                                               ctypes.byref(err)), 'pavgpu_cigar_batch_fetch')
        snv = _capi.take_host_array(ps.value, ns.value, _capi.SNV_ROW)
        indel = _capi.take_host_array(pi.value, ni.value, _capi.INDEL_ROW)
        return snv, indel, err

    def close(self):
        if getattr(self, 'handle', None):
            _capi.lib().pavgpu_cigar_batch_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def cigar_call(ctx, ref_store, qry_store, ref_id, qry_id, pos, rev, ops, op_off):
    """One-shot host-buffer call (pavgpu_cigar_call): returns (snv rows, indel rows, CigarErr, CigarStats)."""
    L = _capi.lib()
    a = [np.ascontiguousarray(ref_id, np.int32), np.ascontiguousarray(qry_id, np.int32), np.ascontiguousarray(pos, np.int32),
         np.ascontiguousarray(rev, np.uint8), np.ascontiguousarray(ops, np.uint32), np.ascontiguousarray(op_off, np.int64)]
    ps, pi = c_vp(), c_vp()
    ns, ni = c_i64(), c_i64()
    err, st = _capi.CigarErr(), _capi.CigarStats()
    _capi.check(L.pavgpu_cigar_call(ctx.handle, ref_store.handle, qry_store.handle, len(a[0]), *[_capi.ptr(x) for x in a],
                                    ctypes.byref(ps), ctypes.byref(ns), ctypes.byref(pi), ctypes.byref(ni), ctypes.byref(err),
                                    ctypes.byref(st)), 'pavgpu_cigar_call')
    snv = _capi.take_host_array(ps.value, ns.value, _capi.SNV_ROW)
    indel = _capi.take_host_array(pi.value, ni.value, _capi.INDEL_ROW)
    return snv, indel, err, st


def homology(seq, sv, positions, ctx=None):
    """Device evaluation of left/right homology for upper-case strings (parity helper)."""
    ctx = ctx or get_context()
    pos = np.ascontiguousarray(positions, dtype=np.int64)
    left = np.zeros(len(pos), np.int32)
    right = np.zeros(len(pos), np.int32)
    s, v = seq.encode(), sv.encode()
    _capi.check(_capi.lib().pavgpu_homology(ctx.handle, len(pos), s, len(s), v, len(v), _capi.ptr(pos), _capi.ptr(left), _capi.ptr(right)),
                'pavgpu_homology')
    return left, right
