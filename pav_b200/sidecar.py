"""Packed-reference sidecar: the reference genome as the walk wants it, next to the FASTA.

PAV starts one Python process per Snakemake job, and every ``call_cigar`` / ``call_inv_batch`` job opens the same
``data/ref/ref.fa.gz`` (``Snakefile:30``), inflates the chromosomes it needs, and -- here -- uploads them as ASCII and packs
them on the GPU. The sidecar is built once per reference (``python -m pav_b200.sidecar ref.fa.gz``) and holds

    header   magic, JSON: sequence names / lengths, section offsets, size + mtime of the FASTA it was built from
    ascii    the bases as they are in the FASTA (original case and IUPAC letters: REF / SEQ strings are sliced from here)
    pack2    the 2-bit plane      } exactly the planes ``pavgpu_seqstore_create`` builds on the device
    nmask    the non-ACGT plane   } (exported with ``pavgpu_seqstore_export``), all sequences in file order

It is memory-mapped: jobs on one box share its pages, nothing is inflated or parsed, the host arrays are views, and the upload
is 0.375 B/base (``pavgpu_seqstore_create_packed``) instead of 1 B/base + a pack kernel. ``cigarcall.make_insdel_snv_calls`` uses
``<ref>.pavsc`` when it exists and is fresh (same size and mtime as the FASTA), or the file named by ``PAVGPU_SIDECAR``; with
``PAVGPU_REF_CACHE=1`` the uploaded store also stays resident in the process for later calls.

Building needs the GPU (the planes come from the device pack kernel: there is no second, CPU implementation of the format).
"""
import json
import os
import struct
import sys

import numpy as np

from . import device, fasta

MAGIC = b'PAVSC1\n\0'
SUFFIX = '.pavsc'
_ALIGN = 4096


def _pad(n, a=_ALIGN):
    return (n + a - 1) // a * a


def write(path, names, arrays, pack2, nmask, source=None):
    """Write a sidecar from host arrays (``pack2`` uint64 / ``nmask`` uint32: the planes of a store holding ``arrays`` in order)."""
    lengths = [int(len(a)) for a in arrays]
    ascii_off, off = [], 0
    for ln in lengths:
        ascii_off.append(off)
        off += _pad(ln, 64)
    meta = {'names': [str(n) for n in names], 'lengths': lengths, 'ascii_off': ascii_off, 'ascii_bytes': off,
            'pack2_bytes': int(pack2.nbytes), 'nmask_bytes': int(nmask.nbytes), 'layout': PLANE_LAYOUT}
    if source:
        st = os.stat(source)
        meta['source'] = {'path': os.path.abspath(source), 'size': st.st_size, 'mtime_ns': st.st_mtime_ns}
    body = json.dumps(meta).encode()
    data0 = _pad(len(MAGIC) + 8 + len(body))
    meta_off = {'ascii': data0, 'pack2': data0 + _pad(off), 'nmask': data0 + _pad(off) + _pad(int(pack2.nbytes))}
    tmp = f'{path}.{os.getpid()}.tmp'
    with open(tmp, 'wb') as fh:
        fh.write(MAGIC + struct.pack('<Q', len(body)) + body)
        for a, o in zip(arrays, ascii_off):
            fh.seek(meta_off['ascii'] + o)
            fh.write(np.ascontiguousarray(a, dtype=np.uint8).tobytes())
        fh.seek(meta_off['pack2'])
        fh.write(np.ascontiguousarray(pack2, dtype=np.uint64).tobytes())
        fh.seek(meta_off['nmask'])
        fh.write(np.ascontiguousarray(nmask, dtype=np.uint32).tobytes())
        fh.truncate(meta_off['nmask'] + _pad(int(nmask.nbytes)))
    os.replace(tmp, path)
    return path


def build(fasta_path, out_path=None, ctx=None):
    """Pack every record of ``fasta_path`` on the GPU and write ``<fasta_path>.pavsc`` (or ``out_path``)."""
    fa = fasta.open_fasta(fasta_path)
    names = fa.names()
    arrays = [fa.fetch_array(n) for n in names]
    ctx = ctx or device.get_context()
    store = device.SeqStore(ctx, names, arrays, keep_host=False)
    try:
        pack2, nmask = store.export()
    finally:
        store.close()
    return write(out_path or fasta_path + SUFFIX, names, arrays, pack2, nmask, source=fasta_path)


# Layout of the packed planes a sidecar holds: sequence starts aligned to 128 bases, one 128-base tail guard, 32 bases per 64-bit word
# (first base most significant), one mask bit per base. A file written under another layout is refused (the C call checks the byte
# sizes against the layout the loaded library derives from the sequence lengths as well).
PLANE_LAYOUT = {'seq_align': 128, 'tail_guard': 128, 'version': 1}


class Sidecar:
    def __init__(self, path):
        self.path = path
        self._map = np.memmap(path, dtype=np.uint8, mode='r')
        if len(self._map) < 16 or bytes(self._map[:8]) != MAGIC:
            raise RuntimeError(f'{path}: not a pav_b200 sidecar')
        n = struct.unpack('<Q', bytes(self._map[8:16]))[0]
        self.meta = json.loads(bytes(self._map[16:16 + n]).decode())
        self.names = list(self.meta['names'])
        self.lengths = np.array(self.meta['lengths'], dtype=np.int64)
        self.ids = {nm: i for i, nm in enumerate(self.names)}
        d0 = _pad(16 + n)
        self._ascii = d0
        self._pack2 = d0 + _pad(self.meta['ascii_bytes'])
        self._nmask = self._pack2 + _pad(self.meta['pack2_bytes'])
        if len(self._map) < self._nmask + self.meta['nmask_bytes']:
            raise RuntimeError(f'{path}: truncated sidecar')
        if self.meta.get('layout', PLANE_LAYOUT) != PLANE_LAYOUT:
            raise RuntimeError(f'{path}: packed under plane layout {self.meta.get("layout")}, this build uses {PLANE_LAYOUT}; rebuild the sidecar')
        words = int(sum((int(n) + 127) // 128 * 128 for n in self.lengths) + 128) // 32
        if self.meta['pack2_bytes'] != 8 * words or self.meta['nmask_bytes'] != 4 * words:
            raise RuntimeError(f'{path}: plane sizes do not match the sequence lengths (corrupt or foreign sidecar)')

    def fresh_for(self, fasta_path):
        src = self.meta.get('source')
        if not src:
            return False
        st = os.stat(fasta_path)
        return src['size'] == st.st_size and src['mtime_ns'] == st.st_mtime_ns

    def fetch_array(self, name):
        """Bases of one sequence as a read-only view of the mapped file (original case / IUPAC letters)."""
        i = self.ids[str(name)]
        o = self._ascii + self.meta['ascii_off'][i]
        return self._map[o:o + int(self.lengths[i])]

    def arrays(self):
        return [self.fetch_array(n) for n in self.names]

    def planes(self):
        pack2 = self._map[self._pack2:self._pack2 + self.meta['pack2_bytes']].view(np.uint64)
        nmask = self._map[self._nmask:self._nmask + self.meta['nmask_bytes']].view(np.uint32)
        return pack2, nmask

    def store(self, ctx=None):
        """Upload the packed planes (no pack kernel, 0.375 B/base) -> ``device.SeqStore`` with all sequences in file order."""
        ctx = ctx or device.get_context()
        pack2, nmask = self.planes()
        return device.SeqStore.from_packed(ctx, self.names, self.lengths, pack2, nmask)


_OPEN = {}
_RESIDENT = {}   # PAVGPU_REF_CACHE=1: sidecar key -> SeqStore kept on the device for the life of the process


def find(fasta_path):
    """Sidecar to use for ``fasta_path``: ``PAVGPU_SIDECAR`` if set (trusted), else a fresh ``<fasta_path>.pavsc``; else ``None``."""
    forced = os.environ.get('PAVGPU_SIDECAR')
    path = forced or (fasta_path + SUFFIX)
    if not os.path.exists(path):
        if forced:
            raise RuntimeError(f'PAVGPU_SIDECAR={forced}: no such file')
        return None
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    sc = _OPEN.get(key)
    if sc is None:
        if len(_OPEN) > 4:
            _OPEN.clear()
        sc = _OPEN[key] = Sidecar(path)
        sc.key = key
    if not forced and not sc.fresh_for(fasta_path):
        return None
    return sc


def reference_store(sc, ctx=None):
    """``(store, owned)``: a store with the sidecar's planes; ``owned`` is False when it is the process-resident one."""
    if os.environ.get('PAVGPU_REF_CACHE') == '1':
        st = _RESIDENT.get(sc.key)
        if st is None or not st.handle:
            if len(_RESIDENT) > 1:
                for old in _RESIDENT.values():
                    old.close()
                _RESIDENT.clear()
            st = _RESIDENT[sc.key] = sc.store(ctx)
        return st, False
    return sc.store(ctx), True


if __name__ == '__main__':
    for p in sys.argv[1:]:
        print(build(p))
