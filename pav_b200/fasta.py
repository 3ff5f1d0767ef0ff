"""Indexed FASTA access without pysam/htslib (absent from the image and not needed on the hot path).

Replaces ``pysam.FastaFile(fn).fetch(name[, start, end])`` as used by the reference
(pavlib/cigarcall.py:58-66, pavlib/seq.py:339-351). Sequences come back as ``numpy.uint8`` arrays
of the original bytes (case and IUPAC codes preserved) -- they are uploaded to the GPU as they are
and sliced on the host for the REF / ALT / SEQ columns.

Plain FASTA is memory-mapped and addressed through the samtools ``.fai`` (built in memory when the
index file is missing). bgzip-compressed FASTA (what PAV's ``data/ref/ref.fa.gz`` is; reference:
rules/data.snakefile:89-118) is read block-wise: BGZF blocks are located through the ``.gzi`` index or, when it is
missing, by walking the block headers (no inflation), and only the blocks covering the requested record are inflated
(in parallel threads; zlib releases the GIL). Plain gzip files are inflated once per process.
"""
import gzip
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np


class _Bgzf:
    """Random access into a BGZF file: ``read(u0, u1)`` returns uncompressed bytes [u0, u1)."""

    def __init__(self, path):
        self.path = path
        self.raw = np.memmap(path, dtype=np.uint8, mode='r')
        self.c_off, self.u_off = self._index()

    @staticmethod
    def is_bgzf(path):
        with open(path, 'rb') as fh:
            h = fh.read(18)
        return len(h) >= 18 and h[:4] == b'\x1f\x8b\x08\x04' and h[12:14] == b'BC'

    def _index(self):
        gzi = self.path + '.gzi'
        if os.path.exists(gzi):
            with open(gzi, 'rb') as fh:
                n = struct.unpack('<Q', fh.read(8))[0]
                arr = np.frombuffer(fh.read(16 * n), dtype='<u8').reshape(n, 2)
            c = np.concatenate(([0], arr[:, 0])).astype(np.int64)
            u = np.concatenate(([0], arr[:, 1])).astype(np.int64)
            return c, u
        # no .gzi: walk the block headers (BSIZE in the 'BC' extra field, ISIZE in the trailer)
        c, u, pos, upos, n = [], [], 0, 0, len(self.raw)
        raw = self.raw
        while pos + 18 <= n:
            bsize = int(raw[pos + 16]) | (int(raw[pos + 17]) << 8)
            isize = int.from_bytes(bytes(raw[pos + bsize - 3:pos + bsize + 1]), 'little')
            c.append(pos)
            u.append(upos)
            pos += bsize + 1
            upos += isize
        return np.array(c, dtype=np.int64), np.array(u, dtype=np.int64)

    def _inflate(self, lo, hi):
        """Inflate blocks lo..hi-1 (concatenated gzip members)."""
        c0 = int(self.c_off[lo])
        c1 = int(self.c_off[hi]) if hi < len(self.c_off) else len(self.raw)
        d = zlib.decompressobj(31)
        data = bytes(self.raw[c0:c1])
        out = []
        while data:
            out.append(d.decompress(data))
            data = d.unused_data
            if not data:
                break
            d = zlib.decompressobj(31)
        return b''.join(out)

    def read(self, u0, u1):
        if u1 <= u0:
            return b''
        lo = int(np.searchsorted(self.u_off, u0, side='right')) - 1
        hi = int(np.searchsorted(self.u_off, u1 - 1, side='right'))
        n_blk = hi - lo
        if n_blk > 64:   # ~64 KiB per block: split into ~4 MiB jobs
            step = 64
            jobs = [(a, min(a + step, hi)) for a in range(lo, hi, step)]
            with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
                parts = list(pool.map(lambda ab: self._inflate(*ab), jobs))
            data = b''.join(parts)
        else:
            data = self._inflate(lo, hi)
        start = u0 - int(self.u_off[lo])
        return data[start:start + (u1 - u0)]

_CACHE = {}


class Fasta:
    def __init__(self, path):
        self.path = path
        self._plain = not str(path).endswith('.gz')
        self._bgzf = None
        if self._plain:
            self._buf = np.memmap(path, dtype=np.uint8, mode='r') if os.path.getsize(path) else np.zeros(0, np.uint8)
        elif os.path.exists(path + '.fai') and _Bgzf.is_bgzf(path):
            self._bgzf = _Bgzf(path)      # block-wise access; nothing is inflated until a record is fetched
            self._buf = None
        else:
            with gzip.open(path, 'rb') as fh:
                self._buf = np.frombuffer(fh.read(), dtype=np.uint8)
        self.index = self._read_fai() if os.path.exists(path + '.fai') and (self._plain or self._bgzf) else self._scan_index()
        self._seq_cache = {}

    def _read_fai(self):
        idx = {}
        with open(self.path + '.fai') as fh:
            for line in fh:
                tok = line.rstrip('\n').split('\t')
                if len(tok) >= 5:
                    idx[tok[0]] = (int(tok[1]), int(tok[2]), int(tok[3]), int(tok[4]))
        return idx

    def _scan_index(self):
        buf = self._buf
        idx = {}
        gt = np.flatnonzero(buf == ord('>'))
        nl = np.flatnonzero(buf == 10)
        # keep only '>' at line starts
        starts = [g for g in gt.tolist() if g == 0 or buf[g - 1] == 10]
        for i, g in enumerate(starts):
            k = np.searchsorted(nl, g)
            hdr_end = int(nl[k]) if k < len(nl) else len(buf)
            name = bytes(buf[g + 1:hdr_end]).split()[0].decode() if hdr_end > g + 1 else ''
            body_start = hdr_end + 1
            body_end = starts[i + 1] if i + 1 < len(starts) else len(buf)
            body_nl = nl[np.searchsorted(nl, body_start):np.searchsorted(nl, body_end)]
            if len(body_nl):
                linebases = int(body_nl[0]) - body_start
                if linebases > 0 and body_start + linebases - 1 < len(buf) and buf[body_start + linebases - 1] == 13:
                    linebases -= 1
                    linewidth = linebases + 2
                else:
                    linewidth = linebases + 1
            else:
                linebases = body_end - body_start
                linewidth = linebases
            n_nl = len(body_nl)
            total = body_end - body_start
            cr = (linewidth - linebases - 1) if linewidth > linebases else 0
            length = total - n_nl * (1 + max(cr, 0))
            idx[name] = (int(length), int(body_start), int(max(linebases, 1)), int(max(linewidth, 1)))
        return idx

    def names(self):
        return list(self.index.keys())

    def length(self, name):
        return self.index[str(name)][0]

    def fetch_array(self, name, start=None, end=None):
        """Bases ``[start, end)`` of record ``name`` as a uint8 array (whole record by default)."""
        name = str(name)
        if name not in self.index:
            raise KeyError(f'sequence {name!r} not found in {self.path}')
        length, offset, linebases, linewidth = self.index[name]
        start = 0 if start is None else max(int(start), 0)
        end = length if end is None else min(int(end), length)
        if end <= start:
            return np.zeros(0, dtype=np.uint8)
        if start == 0 and end == length and name in self._seq_cache:
            return self._seq_cache[name]
        first_line, last_line = start // linebases, (end - 1) // linebases
        b0 = offset + first_line * linewidth
        if self._bgzf is not None:
            b1 = offset + last_line * linewidth + linebases
            raw = np.frombuffer(self._bgzf.read(b0, b1), dtype=np.uint8)
        else:
            b1 = min(offset + last_line * linewidth + linebases, len(self._buf))
            raw = np.asarray(self._buf[b0:b1])
        n_lines = last_line - first_line + 1
        if n_lines == 1:
            seq = raw
        else:
            pad = n_lines * linewidth - len(raw)
            block = np.concatenate((raw, np.zeros(pad, dtype=np.uint8))) if pad else raw
            seq = block.reshape(n_lines, linewidth)[:, :linebases].reshape(-1)
        lo = start - first_line * linebases
        seq = np.ascontiguousarray(seq[lo:lo + (end - start)])
        if start == 0 and end == length:
            if len(self._seq_cache) > 64:
                self._seq_cache.clear()
            self._seq_cache[name] = seq
        return seq

    def fetch_into(self, name, out):
        """Whole record ``name`` written into ``out`` (uint8, at least the record's length; e.g. a pinned staging buffer):
        one strided copy of the full lines + the last partial line. Returns ``out[:length]``."""
        name = str(name)
        if name not in self.index:
            raise KeyError(f'sequence {name!r} not found in {self.path}')
        length, offset, linebases, linewidth = self.index[name]
        dst = out[:length]
        if length == 0:
            return dst
        if self._bgzf is not None or self._buf is None or len(self._buf) < offset:
            dst[:] = self.fetch_array(name)
            return dst
        full = length // linebases
        if full:
            end_full = offset + full * linewidth
            if end_full <= len(self._buf):
                src = self._buf[offset:end_full].reshape(full, linewidth)[:, :linebases]
            else:   # the file ends right after the last full line, without its newline
                full -= 1
                src = self._buf[offset:offset + full * linewidth].reshape(full, linewidth)[:, :linebases]
            np.copyto(dst[:full * linebases].reshape(full, linebases), src)
        rem = length - full * linebases
        if rem:
            b0 = offset + full * linewidth
            dst[full * linebases:] = self._buf[b0:b0 + rem]
        return dst

    def line_chunks(self, name, chunk_bases):
        """Split record ``name`` for ``fetch_lines_into``: ``[(line0, line1), ...]`` over its FULL lines, about ``chunk_bases`` each
        (the last chunk also carries the partial last line); ``None`` when the record cannot be copied by lines (compressed file)."""
        name = str(name)
        length, offset, linebases, linewidth = self.index[name]
        if self._bgzf is not None or self._buf is None or len(self._buf) < offset or length == 0 or linebases <= 0:
            return None
        full = length // linebases
        if full and offset + full * linewidth > len(self._buf):
            full -= 1       # (the file ends right after the last full line, without its newline: fetch_into handles that record whole)
            return None
        per = max(int(chunk_bases) // linebases, 1)
        cuts = list(range(0, full, per)) + [full]
        if len(cuts) == 1:
            cuts = [0, 0]
        return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]

    def fetch_lines_into(self, name, out, line0, line1):
        """Full lines ``[line0, line1)`` of record ``name`` into their place in ``out`` (laid out like ``fetch_into``'s result); the
        chunk that ends at the record's last full line also copies the partial line behind it. Chunks of one record can be copied
        by different threads."""
        name = str(name)
        length, offset, linebases, linewidth = self.index[name]
        full = length // linebases
        n = line1 - line0
        if n > 0:
            src = self._buf[offset + line0 * linewidth:offset + line1 * linewidth].reshape(n, linewidth)[:, :linebases]
            np.copyto(out[line0 * linebases:line1 * linebases].reshape(n, linebases), src)
        if line1 == full:
            rem = length - full * linebases
            if rem:
                b0 = offset + full * linewidth
                out[full * linebases:length] = self._buf[b0:b0 + rem]

    def fetch(self, name, start=None, end=None):
        """``pysam.FastaFile.fetch`` look-alike returning ``str``."""
        return self.fetch_array(name, start, end).tobytes().decode('ascii')


def open_fasta(path):
    """Process-wide cache keyed by (path, mtime, size)."""
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    fa = _CACHE.get(key)
    if fa is None:
        if len(_CACHE) > 8:
            _CACHE.clear()
        fa = _CACHE[key] = Fasta(path)
    return fa


# Complement over bytes: Biopython's ambiguous-DNA table, case preserved (Bio.Seq.reverse_complement is
# what the reference applies to REV contigs, pavlib/cigarcall.py:69-70).
COMPLEMENT = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b'ACGTMRWSYKVHDBNacgtmrwsykvhdbn', b'TGCAKYWSRMBDHVNtgcakywsrmbdhvn'):
    COMPLEMENT[_a] = _b
UPPER = np.arange(256, dtype=np.uint8)
UPPER[ord('a'):ord('z') + 1] -= 32


def reverse_complement(arr):
    return COMPLEMENT[arr[::-1]]
