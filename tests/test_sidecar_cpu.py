"""Packed-reference sidecar file format (pav_b200/sidecar.py) without a GPU: write -> map -> read, freshness, errors.
The planes written here come from a small numpy restatement of the plane layout (test code; the product packs on the device)."""
import os

import numpy as np
import pytest

from pav_b200 import sidecar, synth

SEQ_ALIGN = 128


def numpy_planes(arrays):
    """pack2 / nmask exactly as DESIGN.md section 2 describes them (first base most significant, 128-base aligned rows, guard)."""
    offs, off = [], 0
    for a in arrays:
        offs.append(off)
        off += (len(a) + SEQ_ALIGN - 1) // SEQ_ALIGN * SEQ_ALIGN
    off += SEQ_ALIGN
    code = np.full(off, 4, dtype=np.uint8)
    lut = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b'ACGT'):
        lut[ch] = i
        lut[ch | 0x20] = i
    for a, o in zip(arrays, offs):
        code[o:o + len(a)] = lut[a]
    c = code.reshape(-1, 32)
    shifts = (62 - 2 * np.arange(32)).astype(np.uint64)
    pack2 = ((c & 3).astype(np.uint64) << shifts).sum(axis=1, dtype=np.uint64)
    nmask = ((c >> 2).astype(np.uint32) << np.arange(32, dtype=np.uint32)).sum(axis=1, dtype=np.uint32)
    return pack2, nmask


def test_roundtrip_and_freshness(tmp_path):
    rng = np.random.default_rng(3)
    seqs = {'chr1': synth.random_seq(rng, 1000), 'chr2': synth.random_seq(rng, 257), 'chrEmpty': np.zeros(0, np.uint8)}
    seqs['chr1'][100:140] |= 0x20
    seqs['chr2'][7] = ord('N')
    fa = str(tmp_path / 'ref.fa')
    synth.write_fasta(fa, seqs)
    pack2, nmask = numpy_planes(list(seqs.values()))
    path = sidecar.write(fa + sidecar.SUFFIX, list(seqs), list(seqs.values()), pack2, nmask, source=fa)
    sc = sidecar.Sidecar(path)
    assert sc.names == list(seqs) and sc.lengths.tolist() == [1000, 257, 0]
    for k, v in seqs.items():
        assert (sc.fetch_array(k) == v).all()
    p2, nm = sc.planes()
    assert (p2 == pack2).all() and (nm == nmask).all()
    assert sc.fresh_for(fa)
    found = sidecar.find(fa)
    assert found is not None and found.path == path
    # a changed FASTA makes the sidecar stale: it is ignored, not trusted
    os.utime(fa, ns=(1, 1))
    assert not sc.fresh_for(fa) and sidecar.find(fa) is None


def test_rejects_foreign_and_truncated_files(tmp_path):
    p = tmp_path / 'x.pavsc'
    p.write_bytes(b'not a sidecar at all, just bytes')
    with pytest.raises(RuntimeError, match='not a pav_b200 sidecar'):
        sidecar.Sidecar(str(p))
    seqs = [np.frombuffer(b'ACGT' * 100, dtype=np.uint8)]
    pack2, nmask = numpy_planes(seqs)
    good = sidecar.write(str(tmp_path / 'g.pavsc'), ['s'], seqs, pack2, nmask)
    data = open(good, 'rb').read()
    bad = tmp_path / 'b.pavsc'
    bad.write_bytes(data[:len(data) // 2])
    with pytest.raises(RuntimeError, match='truncated'):
        sidecar.Sidecar(str(bad))


def test_forced_sidecar_must_exist(tmp_path, monkeypatch):
    monkeypatch.setenv('PAVGPU_SIDECAR', str(tmp_path / 'missing.pavsc'))
    with pytest.raises(RuntimeError, match='no such file'):
        sidecar.find(str(tmp_path / 'ref.fa'))


def test_rejects_planes_of_another_layout(tmp_path):
    """A sidecar whose plane sizes do not follow from its sequence lengths under this build's layout (written by another layout, or
    with a doctored header) is refused when it is opened -- before any byte of it could be handed to the device."""
    import json
    import struct
    from pav_b200 import sidecar
    names, arrays = ['a', 'b'], [np.frombuffer(b'ACGT' * 100, np.uint8), np.frombuffer(b'GGCC' * 10, np.uint8)]
    words = sum((len(a) + 127) // 128 * 128 for a in arrays) // 32 + 4
    good = sidecar.write(str(tmp_path / 'ok.pavsc'), names, arrays, np.zeros(words, np.uint64), np.zeros(words, np.uint32))
    sidecar.Sidecar(good)
    short = sidecar.write(str(tmp_path / 'short.pavsc'), names, arrays, np.zeros(words - 4, np.uint64), np.zeros(words - 4, np.uint32))   # no tail guard
    with pytest.raises(RuntimeError, match='plane sizes'):
        sidecar.Sidecar(short)
    raw = bytearray(open(good, 'rb').read())
    n = struct.unpack('<Q', raw[8:16])[0]
    meta = json.loads(raw[16:16 + n].decode())
    meta['layout'] = {'seq_align': 64, 'tail_guard': 128, 'version': 1}
    blob = json.dumps(meta).encode()
    assert len(blob) <= n + 64
    blob = blob.ljust(n)[:n] if len(blob) <= n else blob
    if len(blob) == n:
        raw[16:16 + n] = blob
        other = tmp_path / 'other.pavsc'
        other.write_bytes(bytes(raw))
        with pytest.raises(RuntimeError, match='plane layout'):
            sidecar.Sidecar(str(other))
