"""The device functions of pav_b200/csrc/seqbits.cuh run on the host (tests/host_emul/seqbits_host.cpp compiles the same
source text with g++ and spells out the CUDA intrinsics) against the oracle, the golden vectors and plain string code:
sequence packing, 32-base windows in both orientations, the word-parallel homology scans, the per-indel scoring that all
three homology kernels share (with and without staged tiles), and k-mer extraction. No GPU involved; what this cannot see
(launch geometry, shuffles, shared memory, atomics) is covered by the `-m gpu` parity tests."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle
from pav_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SEQ_ALIGN = 128
c_i64, c_i32, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p


@pytest.fixture(scope='module')
def emu():
    src = os.path.join(HERE, 'host_emul', 'seqbits_host.cpp')
    out_dir = os.path.join(HERE, 'host_emul', '_build')
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, 'seqbits_host.so')
    hdr = os.path.join(REPO, 'pav_b200', 'csrc', 'seqbits.cuh')
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-Wno-unknown-pragmas', '-I', os.path.dirname(hdr), '-o', so + '.tmp', src])
        os.replace(so + '.tmp', so)
    L = ctypes.CDLL(so)
    L.emu_pack.argtypes = [c_vp, c_i64, c_vp, c_vp]
    L.emu_base.argtypes = [c_vp, c_vp, c_i64, c_i64, ctypes.c_int, c_i64]
    L.emu_window.argtypes = [c_vp, c_vp, c_i64, c_i64, ctypes.c_int, c_i32, c_vp, c_vp]
    L.emu_homology.argtypes = [c_vp, c_vp, c_i64, c_i64, ctypes.c_int, c_i64, c_i64, c_i64, ctypes.c_int, c_i64, ctypes.c_int, ctypes.c_int]
    L.emu_score_indel.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, ctypes.c_int] + [c_i32] * 5 + [c_i64, c_i32, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]
    L.emu_kmers.argtypes = [c_vp, c_vp, c_i64, c_i32, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]
    L.emu_nsum.argtypes = [c_vp, c_i64, c_vp]
    L.emu_nbr_first_word.argtypes = [c_i64, c_i64]
    L.emu_nbr_first_word.restype = c_i64
    L.emu_tile_range.argtypes = [c_i64, c_i64, c_i64, c_vp, c_vp]
    return L


class Planes:
    """Sequences laid out like pavgpu_seqstore (SEQ_ALIGN-aligned starts, 'N' padding, one guard block) and packed by the
    device function."""

    def __init__(self, emu, seqs):
        self.seqs = [np.frombuffer(s.encode(), np.uint8) if isinstance(s, str) else np.asarray(s, np.uint8) for s in seqs]
        self.off, off = [], 0
        for s in self.seqs:
            self.off.append(off)
            off += (len(s) + SEQ_ALIGN - 1) // SEQ_ALIGN * SEQ_ALIGN
        self.total = off + SEQ_ALIGN
        self.ascii = np.full(self.total, ord('N'), np.uint8)
        for s, o in zip(self.seqs, self.off):
            self.ascii[o:o + len(s)] = s
        self.words = self.total // 32
        self.pack2 = np.zeros(self.words, np.uint64)
        self.nmask = np.zeros(self.words, np.uint32)
        emu.emu_pack(self.ascii.ctypes.data, self.words, self.pack2.ctypes.data, self.nmask.ctypes.data)
        self.nsum = np.zeros((self.words + 1) // 256 + 2, np.uint32)     # N summary as pavgpu_seqstore sizes it
        emu.emu_nsum(self.nmask.ctypes.data, self.words, self.nsum.ctypes.data)

    def p(self):
        return self.pack2.ctypes.data, self.nmask.ctypes.data


_CODE = np.full(256, 4, np.uint8)
for _i, _c in enumerate('ACGT'):
    _CODE[ord(_c)] = _CODE[ord(_c.lower())] = _i
_COMP = bytes.maketrans(b'ACGTacgt', b'TGCAtgca')


def test_pack_word_layout(emu):
    """pack_word32: base g in bits [62 - 2 (g % 32), +2) of plane word g / 32, mask bit g % 32 for anything but ACGTacgt."""
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 256, 32 * 257, dtype=np.uint8)           # every byte value, not only letters
    raw[:64] = np.frombuffer(b'ACGTacgtNnRYKMSWBDHVU-*.xX@[`{' + b'A' * 34, np.uint8)
    pl = Planes(emu, [raw])
    code = _CODE[pl.ascii]
    g = np.arange(pl.total)
    want_p = np.zeros(pl.words, np.uint64)
    np.bitwise_or.at(want_p, g >> 5, (code & 3).astype(np.uint64) << (62 - 2 * (g & 31)).astype(np.uint64))
    want_m = np.zeros(pl.words, np.uint32)
    np.bitwise_or.at(want_m, g >> 5, (code >> 2).astype(np.uint32) << (g & 31).astype(np.uint32))
    assert (pl.pack2 == want_p).all() and (pl.nmask == want_m).all()


def test_oriented_bases_and_windows(emu):
    """oseq_base / oseq_window in both orientations equal slicing the (reverse-complemented) string, including windows that
    hang over either end of the sequence and sequences that start at a non-zero plane offset."""
    rng = np.random.default_rng(6)
    a = synth.random_seq(rng, 301)
    b = synth.random_seq(rng, 77)
    b[[3, 40, 76]] = ord('N')
    b[10:14] = np.frombuffer(b'acgt', np.uint8)
    pl = Planes(emu, [a, b])
    pp, pm = pl.p()
    for si, s in enumerate(pl.seqs):
        fwd = s.tobytes()
        for rev in (0, 1):
            text = fwd[::-1].translate(_COMP) if rev else fwd
            code = _CODE[np.frombuffer(text, np.uint8)]
            for t in list(range(-40, 45)) + list(range(len(s) - 45, len(s) + 40)):
                want = int(code[t]) if 0 <= t < len(s) else 4
                assert emu.emu_base(pp, pm, pl.off[si], len(s), rev, t) == want
                bases, mask = ctypes.c_uint64(), ctypes.c_uint32()
                emu.emu_window(pp, pm, pl.off[si], len(s), rev, t, ctypes.byref(bases), ctypes.byref(mask))
                for i in range(32):
                    u = t + i
                    c = int(code[u]) if 0 <= u < len(s) else 4
                    assert (mask.value >> i) & 1 == (1 if c == 4 else 0), (si, rev, t, i)
                    if c != 4:
                        assert (bases.value >> (62 - 2 * i)) & 3 == c, (si, rev, t, i)


def test_homology_scans_golden_and_random(emu):
    """dev_homology_raw against the reference's own left_homology / right_homology outputs (tests/golden/homology.json) and
    against the oracle on random flanks with short, repetitive and N-containing SV sequences."""
    cases = json.load(open(os.path.join(HERE, 'golden', 'homology.json')))
    n_checked = 0
    for c in cases:
        if not c['seq'] or not c['sv']:
            continue
        pl = Planes(emu, [c['seq'].upper(), c['sv'].upper()])
        pp, pm = pl.p()
        T, V = len(c['seq']), len(c['sv'])
        p = c['pos']
        if c['left'] is not None and p < T:
            assert emu.emu_homology(pp, pm, pl.off[0], T, 0, p, pl.off[1], V, 0, 0, V, 1) == c['left'], c
            n_checked += 1
        if c['right'] is not None and p >= 0:
            assert emu.emu_homology(pp, pm, pl.off[0], T, 0, p, pl.off[1], V, 0, 0, V, 0) == c['right'], c
            n_checked += 1
    assert n_checked > 20
    rng = np.random.default_rng(7)
    for _ in range(300):
        unit = synth.random_seq(rng, int(rng.integers(1, 9))).tobytes().decode()
        sv = (unit * 40)[:int(rng.integers(1, 70))]
        flank_l = synth.random_seq(rng, int(rng.integers(0, 90))).tobytes().decode() + unit * int(rng.integers(0, 30))
        flank_r = unit * int(rng.integers(0, 30)) + synth.random_seq(rng, int(rng.integers(0, 90))).tobytes().decode()
        seq = list(flank_l + sv + flank_r)
        if rng.random() < 0.3 and seq:
            seq[int(rng.integers(0, len(seq)))] = 'N'
        seq = ''.join(seq)
        pl = Planes(emu, [seq, sv])
        pp, pm = pl.p()
        for p in {len(flank_l) - 1, len(flank_l) + len(sv), 0, len(seq) - 1, int(rng.integers(0, len(seq)))}:
            if 0 <= p < len(seq):
                assert emu.emu_homology(pp, pm, pl.off[0], len(seq), 0, p, pl.off[1], len(sv), 0, 0, len(sv), 1) == pyoracle.left_homology(p, seq, sv)
                assert emu.emu_homology(pp, pm, pl.off[0], len(seq), 0, p, pl.off[1], len(sv), 0, 0, len(sv), 0) == pyoracle.right_homology(p, seq, sv)


def _stubs(df):
    """What the walk hands to the homology kernels for every I / D op: (rec, svtype, n, pos_ref, pos_qry, '=' run before)."""
    out = []
    for rec, row in enumerate(df.itertuples(index=False)):
        pr, pq, prev = int(row.POS), 0, (0, '')
        num = ''
        for ch in row.CIGAR:
            if ch.isdigit():
                num += ch
                continue
            n, num = int(num), ''
            if ch in 'ID':
                out.append((rec, 0 if ch == 'I' else 1, n, pr, pq, prev[0] if prev[1] == '=' else 0))
            if ch in '=XD':
                pr += n
            if ch in '=XISH':
                pq += n
            prev = (n, ch)
    return out


@pytest.mark.parametrize('seed,kw', [
    (21, dict(n_chrom=2, chrom_len=120_000, n_contig=12, contig_len=15_000, edit_rate=0.012, rev_frac=0.5)),
    (22, dict(n_chrom=1, chrom_len=90_000, n_contig=3, contig_len=30_000, edit_rate=0.02, rev_frac=0.5, clip=(7, 5), soft_mask_frac=0.5,
              n_block_frac=0.05)),
])
def test_score_indel_matches_oracle(emu, tmp_path, seed, kw):
    """score_indel and score_indel2 (the bodies the homology kernels share: per-scan loops / convergent first trips against the
    circular SV pattern) on every indel of a seeded workload: plain, with the 8-word neighbourhood the nbr and bulk kernels stage,
    and with random tiles -- all equal the oracle's rows."""
    ref, tigs, df = synth.make_cigar_workload(seed, **kw)
    ref_fa, tig_fa, _ = synth.write_cigar_workload(str(tmp_path), ref, tigs, df)
    _, o_indel, _ = pyoracle.walk_rows(df, ref_fa, tig_fa)
    names_r, names_t = list(ref), list(tigs)
    R = Planes(emu, [ref[n] for n in names_r])
    Q = Planes(emu, [tigs[n] for n in names_t])
    stubs = _stubs(df)
    assert len(stubs) == len(o_indel) > 200
    rng = np.random.default_rng(seed)
    fields = ('pos', 'end', 'qry_pos', 'qry_end', 'left_shift', 'hom_ref_l', 'hom_ref_r', 'hom_tig_l', 'hom_tig_r')
    rp, rm = R.p()
    qp, qm = Q.p()
    out = (c_i32 * 10)()
    rows = df.reset_index(drop=True)
    for k, (rec, svtype, n, pr, pq, eqb) in enumerate(stubs):
        ri, qi = names_r.index(rows.at[rec, '#CHROM']), names_t.index(rows.at[rec, 'QRY_ID'])
        rl, ql, rev = len(R.seqs[ri]), len(Q.seqs[qi]), int(bool(rows.at[rec, 'REV']))
        want = [int(o_indel[f][k]) for f in fields]
        assert (int(o_indel['rec'][k]), int(o_indel['svtype'][k]), int(o_indel['svlen'][k])) == (rec, svtype, n)
        cr = R.off[ri] + pr
        cq = Q.off[qi] + (ql - 1 - pq if rev else pq)
        tiles = [(0, 0, 0, 0), (emu.emu_nbr_first_word(cr, R.words), 8, emu.emu_nbr_first_word(cq, Q.words), 8)]
        w0r = int(rng.integers(0, R.words - 1)); w0q = int(rng.integers(0, Q.words - 1))
        tiles.append((w0r, int(rng.integers(1, min(64, R.words - w0r) + 1)), w0q, int(rng.integers(1, min(64, Q.words - w0q) + 1))))
        tiles.append((max((cr >> 5) - 3, 0), min(7, R.words - max((cr >> 5) - 3, 0)), 0, 0))      # reference staged, contig not
        for w0r, nr, w0q, nq in tiles:
            for version in (1, 2):
                for sums in ((None, None), (R.nsum.ctypes.data, Q.nsum.ctypes.data)) if nr == 0 and nq == 0 else ((None, None),):   # with / without N summaries
                    emu.emu_score_indel(rp, rm, R.off[ri], rl, qp, qm, Q.off[qi], ql, rev, svtype, n, pr, pq, eqb, w0r, nr, w0q, nq, version,
                                        sums[0], sums[1], out)
                    assert list(out)[:9] == want, (version, sums[0] is not None, k, rec, svtype, n, pr, pq, eqb, (w0r, nr, w0q, nq))


def test_kmers_from_planes(emu):
    """kmer_at / kmer_revcomp over a window with N runs and lower case equal kanapy's stream (oracle port) for k = 31, 21, 5."""
    rng = np.random.default_rng(8)
    s = synth.random_seq(rng, 3000)
    s[100:103] = ord('N')
    s[1500] = ord('n')
    s[2000:2100] |= 0x20
    lead = synth.random_seq(rng, 333)
    pl = Planes(emu, [lead, s])
    pp, pm = pl.p()
    for k in (31, 21, 5):
        n_pos = len(s) - k + 1
        km = np.zeros(n_pos, np.uint64); rc = np.zeros(n_pos, np.uint64); ok = np.zeros(n_pos, np.uint8)
        emu.emu_kmers(pp, pm, pl.off[1], n_pos, k, km.ctypes.data, rc.ctypes.data, ok.ctypes.data, None)
        km2 = np.zeros(n_pos, np.uint64); rc2 = np.zeros(n_pos, np.uint64); ok2 = np.zeros(n_pos, np.uint8)
        emu.emu_kmers(pp, pm, pl.off[1], n_pos, k, km2.ctypes.data, rc2.ctypes.data, ok2.ctypes.data, pl.nsum.ctypes.data)   # mask loads skipped where the summary is clear
        assert (km2 == km).all() and (rc2 == rc).all() and (ok2 == ok).all()
        want_km, want_ix = pyoracle.kmer_stream(s.tobytes(), k)
        assert (np.flatnonzero(ok) == want_ix).all()
        assert (km[ok == 1] == want_km).all()
        for i in np.flatnonzero(ok)[::97]:
            assert int(rc[i]) == pyoracle.kmer_rc(int(km[i]), k)


def test_staging_ranges(emu):
    """tile_range / nbr_first_word: what the opt-in kernels copy on chip is aligned for their copy sizes, inside the plane, and
    covers the breakpoints with the promised margin; oversize spans and undersized planes are refused."""
    rng = np.random.default_rng(9)
    W, M = emu.emu_tile_words(), emu.emu_tile_margin()
    assert W % 4 == 0
    for _ in range(20000):
        words = int(rng.choice([4, 8, 12, 64, 4096, 1 << 22])) if rng.random() < 0.5 else 4 * int(rng.integers(1, 1 << 20))
        total = words * 32
        c = int(rng.integers(0, total))
        w = emu.emu_nbr_first_word(c, words)
        if words < 8:
            assert w == -1
        else:
            assert w % 4 == 0 and 0 <= w <= words - 8
            lo, hi = w * 32, w * 32 + 256
            assert lo <= c < hi
            assert (c - lo >= 64 or w == 0) and (hi - c > 64 or w == words - 8)    # 64 bases either side unless clamped
        span = int(rng.integers(0, 40000)) if rng.random() < 0.8 else int(rng.integers(0, total))
        lo_g = int(rng.integers(0, total))
        hi_g = min(lo_g + span, total - 1)
        w0, nw = c_i64(), c_i32()
        emu.emu_tile_range(lo_g, hi_g, words, ctypes.byref(w0), ctypes.byref(nw))
        a, b = w0.value, w0.value + nw.value
        want_a = max((lo_g - M) >> 5, 0) & ~3
        want_b = min((((hi_g + M) >> 5) + 2 + 3) & ~3, words)
        assert a == want_a
        if want_b - want_a <= W:
            assert nw.value == want_b - want_a and a % 4 == 0 and nw.value % 4 == 0 and b <= words
            assert a * 32 <= max(lo_g - M, 0) and (b * 32 >= min(hi_g + M + 33, total))    # every window within the margin has both its words staged
        else:
            assert nw.value == 0


# ---- coordinate lifts: pav_b200/csrc/liftcore.cuh on the host -------------------------------------------------------------------
@pytest.fixture(scope='module')
def lift_emu():
    src = os.path.join(HERE, 'host_emul', 'lift_host.cpp')
    out_dir = os.path.join(HERE, 'host_emul', '_build')
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, 'lift_host.so')
    hdr = os.path.join(REPO, 'pav_b200', 'csrc', 'liftcore.cuh')
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-Wno-unknown-pragmas', '-I', os.path.dirname(hdr), '-I', os.path.join(REPO, 'include'),
                               '-o', so + '.tmp', src])
        os.replace(so + '.tmp', so)
    L = ctypes.CDLL(so)
    L.emu_lift_prefix.argtypes = [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp]
    L.emu_lift_prefix.restype = c_i32
    L.emu_lift_points.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]
    return L


class _HostLiftIndex:
    """Stand-in for device.LiftIndex (same ``lift`` interface) running liftcore.cuh on the host."""

    def __init__(self, emu, ops, op_off, pos, rev, qry_len):
        p = lambda a: a.ctypes.data_as(c_vp)  # noqa: E731
        self.emu, self.p = emu, p
        self.ops, self.op_off = np.ascontiguousarray(ops, np.uint32), np.ascontiguousarray(op_off, np.int64)
        self.rev, self.qry_len = np.ascontiguousarray(rev, np.uint8), np.ascontiguousarray(qry_len, np.int64)
        pos = np.ascontiguousarray(pos, np.int64)
        self.ref_start, self.qry_start = np.zeros(len(self.ops), np.int64), np.zeros(len(self.ops), np.int64)
        self.bad_rec = emu.emu_lift_prefix(p(self.ops), p(self.op_off), len(pos), p(pos), p(self.ref_start), p(self.qry_start))

    def lift(self, rec, coord, to_qry):
        rec, coord = np.ascontiguousarray(rec, np.int32), np.ascontiguousarray(coord, np.int64)
        out, status = np.zeros(len(rec), np.int64), np.zeros(len(rec), np.int32)
        p = self.p
        self.emu.emu_lift_points(p(self.ops), p(self.op_off), p(self.ref_start), p(self.qry_start), p(self.rev), p(self.qry_len), len(rec), p(rec), p(coord),
                                 int(bool(to_qry)), p(out), p(status))
        return out, status


def _host_indexed(lift_emu, df, fai):
    from pav_b200 import device
    from pav_b200.pavlib import lift
    al = lift.AlignLift(df, fai)
    ops, op_off, perr = device.parse_cigars(al._cigar.tolist())
    assert perr.code == 0
    qlen = np.array([int(fai[q]) for q in al._qid.tolist()], np.int64)
    idx = _HostLiftIndex(lift_emu, ops, op_off, al._pos, np.asarray(al._rev, bool).astype(np.uint8), qlen)
    assert idx.bad_rec == -1
    al._dev_index = idx
    return al


def test_liftcore_matches_reference_golden(lift_emu):
    """The device lift arithmetic (liftcore.cuh, host build) behind AlignLift.lift_points / lift_regions_to_qry against the answers
    of the reference's own pavlib.align.AlignLift (tests/golden/lift: 3,000 point lifts incl. gap=True and the positions it raises on,
    400 region lifts) -- the CPU twin of tests/test_lift_gpu.py."""
    import json

    import pandas as pd
    from pav_b200.pavlib import seq
    d = os.path.join(REPO, 'tests', 'golden', 'lift')
    df = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t')
    fai = pd.read_csv(os.path.join(d, 'tig.fai.tsv'), sep='\t', header=None, index_col=0)[1]
    al = _host_indexed(lift_emu, df, fai)
    queries = json.load(open(os.path.join(d, 'queries.json')))

    def plain(x):
        return None if x is None else [x[0], int(x[1]), bool(x[2]), int(x[3]), int(x[4]), [int(i) for i in x[5]]]
    n = 0
    for kind, to_qry, gap in (('to_qry', True, False), ('to_sub', False, False), ('to_sub', False, True)):
        qs = [q for q in queries if q['f'] == kind and (kind == 'to_qry' or q['gap'] == gap)]
        good = [q for q in qs if 'error' not in q]
        got = al.lift_points([q['id'] for q in good], [q['pos'] for q in good], to_qry, gap=gap)
        assert [plain(g) for g in got] == [q['result'] for q in good]
        for q in qs:
            if 'error' in q:
                with pytest.raises(RuntimeError):
                    al.lift_points([q['id']], [q['pos']], to_qry, gap=gap)
        n += len(qs)
    regions = [q for q in queries if q['f'] == 'region']
    got = al.lift_regions_to_qry([seq.Region(q['chrom'], q['pos'], q['end']) for q in regions])
    assert [None if r is None else [r.chrom, r.pos, r.end, bool(r.is_rev)] for r in got] == [q['qry'] for q in regions]
    assert n + len(regions) == 3400


def test_liftcore_equals_host_lifts_on_random_tables(lift_emu):
    import pandas as pd
    from pav_b200.pavlib import lift
    for seed, clip in ((5, (0, 0)), (6, (40, 25)), (7, (3, 0))):
        ref, tigs, df = synth.make_cigar_workload(seed, 2, 150_000, 9, 40_000, edit_rate=0.02, rev_frac=0.5, clip=clip)
        df = df.reset_index(drop=True)
        fai = pd.Series({k: len(v) for k, v in tigs.items()})
        host = lift.AlignLift(df, fai)
        host._dev_index = False
        dev = _host_indexed(lift_emu, df, fai)
        rng = np.random.default_rng(seed)
        for to_qry in (True, False):
            ids, coords = [], []
            for _ in range(1500):
                row = df.iloc[int(rng.integers(0, df.shape[0]))]
                if to_qry:
                    ids.append(row['#CHROM']); coords.append(int(rng.integers(row['POS'] - 20, row['END'] + 20)))
                else:
                    ids.append(row['QRY_ID']); coords.append(int(rng.integers(max(row['QRY_POS'] - 60, 0), row['QRY_END'] + 60)))
            exp, bad = [], set()
            for k, (i, c) in enumerate(zip(ids, coords)):
                try:
                    exp.append(host.lift_to_qry(i, c) if to_qry else host.lift_to_sub(i, c))
                except RuntimeError:
                    exp.append('error'); bad.add(k)
            keep = [k for k in range(len(ids)) if k not in bad]
            got = dev.lift_points([ids[k] for k in keep], [coords[k] for k in keep], to_qry)
            assert got == [exp[k] for k in keep]
            for k in sorted(bad)[:10]:
                with pytest.raises(RuntimeError):
                    dev.lift_points([ids[k]], [coords[k]], to_qry)
