"""ctypes binding of libpavgpu.so (include/pavgpu.h). This is the only way the Python host layer
reaches the GPU; there is no CPU fallback -- if the library is missing or no CUDA device is usable
every hot-path call raises ``RuntimeError``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpavgpu.so')

c_i32, c_i64, c_u8, c_vp, c_f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint8, ctypes.c_void_p, ctypes.c_float
P = ctypes.POINTER

SNV_ROW = np.dtype([('pos_ref', '<i4'), ('qry_pos', '<i4'), ('rec', '<i4'), ('op_idx', '<i4')])
INDEL_ROW = np.dtype([('rec', '<i4'), ('op_idx', '<i4'), ('svtype', '<i4'), ('svlen', '<i4'), ('pos', '<i4'), ('end', '<i4'),
                      ('qry_pos', '<i4'), ('qry_end', '<i4'), ('left_shift', '<i4'), ('hom_ref_l', '<i4'),
                      ('hom_ref_r', '<i4'), ('hom_tig_l', '<i4'), ('hom_tig_r', '<i4'), ('seq_start', '<i4'),
                      ('pad', '<i4', (2,))])
assert SNV_ROW.itemsize == 16 and INDEL_ROW.itemsize == 64

DENSITY_WINDOW = np.dtype([('ref_seq_id', '<i4'), ('tig_seq_id', '<i4'), ('ref_pos', '<i4'), ('ref_end', '<i4'),
                           ('tig_pos', '<i4'), ('tig_end', '<i4'), ('rev', '<i4'), ('srs', '<i4')])
CIGAR_REC_STATS = np.dtype([('ref_bp', '<i8'), ('qry_bp', '<i8'), ('lead', '<i8'), ('trail', '<i8'), ('first_body', '<i4'), ('last_body', '<i4'),
                            ('clip_h_first', '<i4'), ('lead_s', '<i4'), ('flags', '<i4'), ('n_ops', '<i4')])
STATE_RUN = np.dtype([('state', '<i4'), ('count', '<i4'), ('first_index', '<i4'), ('last_index', '<i4')])
DENSITY_RESULT = np.dtype([('status', '<i4'), ('smoothed', '<i4'), ('row_off', '<i8'), ('n_rows', '<i8'), ('n_eval', '<i8')])


class ParseErr(ctypes.Structure):
    _fields_ = [('code', c_i32), ('rec', c_i32), ('op_index', c_i64), ('text_pos', c_i64), ('ch', c_i32)]


class CigarErr(ctypes.Structure):
    _fields_ = [('code', c_i32), ('rec', c_i32), ('op_index', c_i64), ('opcode', c_i32), ('pos_ref', c_i32),
                ('pos_qry', c_i32)]


class CigarStats(ctypes.Structure):
    _fields_ = [('ms_h2d', c_f32), ('ms_kernels', c_f32), ('ms_d2h', c_f32), ('ms_scan', c_f32), ('ms_emit', c_f32),
                ('ms_homology', c_f32), ('n_ops', c_i64), ('n_snv', c_i64), ('n_indel', c_i64), ('n_chunks', c_i64),
                ('kernel_launches', c_i32), ('homology_tiled', c_i32), ('ms_count', c_f32), ('walk_passes', c_i32),
                ('graph', c_i32), ('pad0', c_i32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class DensityParams(ctypes.Structure):
    _fields_ = [('k', c_i32), ('min_informative', c_i32), ('min_state_count', c_i32), ('max_ref_kmer_count', c_i32),
                ('smooth', ctypes.c_double), ('delta', ctypes.c_double)]


class DensityStats(ctypes.Structure):
    _fields_ = [('ms_h2d', c_f32), ('ms_kernels', c_f32), ('ms_d2h', c_f32), ('ms_kmer', c_f32), ('ms_kde', c_f32),
                ('ms_fill', c_f32), ('bases', c_i64), ('rows', c_i64), ('kde_pairs', c_i64), ('kernel_launches', c_i32),
                ('kmer_tables_on_chip', c_i32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None

EXPORTS = [
    'pavgpu_last_error', 'pavgpu_device_count', 'pavgpu_ctx_create', 'pavgpu_ctx_destroy', 'pavgpu_ctx_device',
    'pavgpu_free_host', 'pavgpu_host_alloc', 'pavgpu_l2_flush', 'pavgpu_seqstore_create', 'pavgpu_seqstore_create_packed', 'pavgpu_seqstore_create_empty',
    'pavgpu_seqstore_free', 'pavgpu_seqstore_n_seq', 'pavgpu_seqstore_total_bases', 'pavgpu_seqstore_planes',
    'pavgpu_seqstore_export', 'pavgpu_seqstore_checksum', 'pavgpu_seqstore_offset', 'pavgpu_cigar_parse', 'pavgpu_cigar_batch_create',
    'pavgpu_cigar_batch_free', 'pavgpu_cigar_batch_run', 'pavgpu_cigar_batch_fetch', 'pavgpu_cigar_call',
    'pavgpu_homology', 'pavgpu_density_default_params', 'pavgpu_density_batch_create', 'pavgpu_density_batch_free',
    'pavgpu_density_batch_run', 'pavgpu_density_batch_fetch', 'pavgpu_density_batch_fetch_runs', 'pavgpu_density_batch_fetch_window',
    'pavgpu_nccl_unique_id', 'pavgpu_seqstore_broadcast', 'pavgpu_nccl_comm_cached', 'pavgpu_nccl_comm_release_all',
    'pavgpu_cigar_record_stats', 'pavgpu_lift_index_create', 'pavgpu_lift_index_free', 'pavgpu_lift_points',
]


def lib():
    """Load libpavgpu.so (no CUDA call is made by loading)."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    tuned = os.environ.get('PAVGPU_LIB')      # tuning builds only (python -m pav_b200.build --variant NAME ...): a path or a variant name
    if tuned:
        path = tuned if os.sep in tuned else os.path.join(os.path.dirname(LIB_PATH), f'libpavgpu.{tuned}.so')
    if not os.path.exists(path):
        raise RuntimeError(f'libpavgpu.so not built ({path}); run `python -m pav_b200.build`. '
                           'There is no CPU fallback for the hot path.')
    L = ctypes.CDLL(path)
    L.pavgpu_last_error.restype = ctypes.c_char_p
    L.pavgpu_device_count.restype = ctypes.c_int
    L.pavgpu_ctx_create.argtypes = [ctypes.c_int, P(c_vp)]
    L.pavgpu_ctx_destroy.argtypes = [c_vp]
    L.pavgpu_ctx_destroy.restype = None
    L.pavgpu_ctx_device.argtypes = [c_vp]
    L.pavgpu_free_host.argtypes = [c_vp]
    L.pavgpu_host_alloc.argtypes = [c_vp, ctypes.c_size_t, P(c_vp)]
    L.pavgpu_free_host.restype = None
    L.pavgpu_l2_flush.argtypes = [c_vp, ctypes.c_size_t]
    L.pavgpu_seqstore_create.argtypes = [c_vp, c_i32, P(c_vp), P(c_i64), P(c_vp)]
    L.pavgpu_seqstore_create_packed.argtypes = [c_vp, c_i32, P(c_i64), c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t, P(c_vp)]
    L.pavgpu_seqstore_create_empty.argtypes = [c_vp, c_i32, P(c_i64), P(c_vp)]
    L.pavgpu_seqstore_free.argtypes = [c_vp]
    L.pavgpu_seqstore_free.restype = None
    L.pavgpu_seqstore_n_seq.argtypes = [c_vp]
    L.pavgpu_seqstore_n_seq.restype = c_i32
    L.pavgpu_seqstore_total_bases.argtypes = [c_vp]
    L.pavgpu_seqstore_total_bases.restype = c_i64
    L.pavgpu_seqstore_offset.argtypes = [c_vp, c_i32]
    L.pavgpu_seqstore_offset.restype = c_i64
    L.pavgpu_seqstore_planes.argtypes = [c_vp, P(c_vp), P(ctypes.c_size_t), P(c_vp), P(ctypes.c_size_t)]
    L.pavgpu_seqstore_export.argtypes = [c_vp, c_vp, c_vp]
    L.pavgpu_seqstore_checksum.argtypes = [c_vp, c_vp]
    L.pavgpu_cigar_parse.argtypes = [ctypes.c_char_p, P(c_i64), c_i32, P(c_vp), P(c_i64), P(ParseErr)]
    L.pavgpu_cigar_record_stats.argtypes = [c_vp, c_vp, c_vp, c_i32, c_vp]
    L.pavgpu_lift_index_create.argtypes = [c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, P(c_vp), P(c_i32)]
    L.pavgpu_lift_index_free.argtypes = [c_vp]
    L.pavgpu_lift_index_free.restype = None
    L.pavgpu_lift_points.argtypes = [c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]
    L.pavgpu_cigar_batch_create.argtypes = [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, P(c_vp)]
    L.pavgpu_cigar_batch_free.argtypes = [c_vp]
    L.pavgpu_cigar_batch_free.restype = None
    L.pavgpu_cigar_batch_run.argtypes = [c_vp, c_vp, c_vp, P(CigarStats)]
    L.pavgpu_cigar_batch_fetch.argtypes = [c_vp, P(c_vp), P(c_i64), P(c_vp), P(c_i64), P(CigarErr)]
    L.pavgpu_cigar_call.argtypes = [c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    P(c_vp), P(c_i64), P(c_vp), P(c_i64), P(CigarErr), P(CigarStats)]
    L.pavgpu_homology.argtypes = [c_vp, c_i32, ctypes.c_char_p, c_i64, ctypes.c_char_p, c_i64, c_vp, c_vp, c_vp]
    if hasattr(L, 'pavgpu_density_batch_create'):
        L.pavgpu_density_default_params.argtypes = [P(DensityParams)]
        L.pavgpu_density_default_params.restype = None
        L.pavgpu_density_batch_create.argtypes = [c_vp, c_i32, c_vp, P(DensityParams), P(c_vp)]
        L.pavgpu_density_batch_free.argtypes = [c_vp]
        L.pavgpu_density_batch_free.restype = None
        L.pavgpu_density_batch_run.argtypes = [c_vp, c_vp, c_vp, P(DensityStats)]
        L.pavgpu_density_batch_fetch.argtypes = [c_vp, c_vp, P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_i64)]
        L.pavgpu_density_batch_fetch_runs.argtypes = [c_vp, c_vp, P(c_vp), c_vp, P(c_i64)]
        L.pavgpu_density_batch_fetch_window.argtypes = [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
    if hasattr(L, 'pavgpu_nccl_unique_id'):
        L.pavgpu_nccl_unique_id.argtypes = [c_vp]
        L.pavgpu_seqstore_broadcast.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i32, P(c_f32)]
        L.pavgpu_nccl_comm_cached.argtypes = [c_vp, c_i32, c_i32]
        L.pavgpu_nccl_comm_release_all.argtypes = []
        L.pavgpu_nccl_comm_release_all.restype = None
    _lib = L
    return L


def check(rc, what='pavgpu'):
    if rc != 0:
        msg = lib().pavgpu_last_error().decode(errors='replace')
        raise RuntimeError(f'{what} failed (code {rc}): {msg}')


class _HostBuffer:
    """Owner of a library-allocated host buffer (pinned memory from the context's pool, see pavgpu_free_host): numpy arrays
    created over it keep it alive through ``.base``; the buffer goes back to the library when the last view dies."""

    def __init__(self, address, nbytes):
        self._address = address
        self.__array_interface__ = {'data': (address, False), 'shape': (nbytes,), 'typestr': '|u1', 'version': 3}

    def __del__(self):
        addr, self._address = self._address, None
        if addr:
            try:
                lib().pavgpu_free_host(addr)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass


def take_host_array(ptr, n, dtype):
    """Wrap a library-allocated host buffer as a numpy array without copying; it is released with the array."""
    dtype = np.dtype(dtype)
    if not ptr or n == 0:
        if ptr:
            lib().pavgpu_free_host(ptr)
        return np.empty(0, dtype=dtype)
    return np.asarray(_HostBuffer(ptr, n * dtype.itemsize)).view(dtype)


def ptr(arr):
    return arr.ctypes.data_as(c_vp) if arr is not None else None
