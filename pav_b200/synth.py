"""Seeded synthetic inputs for the two hot paths (SURVEY.md §8(d) shapes).

Everything is numpy-vectorised so that the BASELINE C2 shape (1,000 contigs x 200 kbp against a
200 Mbp reference, ~2 M variant rows) is generated in seconds. All randomness comes from
``numpy.random.default_rng(seed)`` (PCG64) so tests, the golden-fixture maker and ``bench.py`` see
identical data on every box.

Path A inputs mirror what PAV's ``rule call_cigar`` hands to
``pavlib.cigarcall.make_insdel_snv_calls`` (reference: rules/call.snakefile:805-810): an alignment
table with ``#CHROM POS END INDEX QRY_ID QRY_POS QRY_END QRY_LEN REV CIGAR`` (=/X CIGARs with hard
clips, reference: pavlib/align/align.py:735-770) plus a reference FASTA and a contig FASTA.

Path B inputs are (reference window, contig window) pairs with a planted inversion, optional
inverted-repeat flanks, SNV divergence and negative controls.
"""
import os

import numpy as np
import pandas as pd

ACGT = np.frombuffer(b'ACGT', dtype=np.uint8)

# Complement table over bytes (IUPAC, case preserved) -- same table Biopython uses.
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b'ACGTMRWSYKVHDBNacgtmrwsykvhdbn', b'TGCAKYWSRMBDHVNtgcakywsrmbdhvn'):
    _COMP[_a] = _b


def revcomp(arr):
    """Reverse complement of an ASCII uint8 array (IUPAC aware, case preserved)."""
    return _COMP[arr[::-1]]


def random_seq(rng, n):
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def _different_base(rng, base):
    """For each ASCII base (ACGT, any case) pick a different upper-case ACGT base."""
    code = np.searchsorted(ACGT, np.frombuffer(bytes(base).upper(), dtype=np.uint8))
    code = np.where(code > 3, 0, code)
    shift = rng.integers(1, 4, size=len(base))
    return ACGT[(code + shift) % 4]


# ----------------------------------------------------------------------------------------------
# FASTA writing
# ----------------------------------------------------------------------------------------------

def write_fasta(path, seqs, line_width=80):
    """Write ``{name: uint8 array}`` as a plain FASTA with a samtools-style ``.fai`` index."""
    fai_lines = []
    offset = 0
    with open(path, 'wb') as fh:
        for name, arr in seqs.items():
            arr = np.ascontiguousarray(arr, dtype=np.uint8)
            hdr = ('>' + name + '\n').encode()
            fh.write(hdr)
            offset += len(hdr)
            n = len(arr)
            fai_lines.append(f'{name}\t{n}\t{offset}\t{line_width}\t{line_width + 1}\n')
            n_full = n // line_width
            body = np.empty((n_full, line_width + 1), dtype=np.uint8)
            body[:, :line_width] = arr[:n_full * line_width].reshape(n_full, line_width)
            body[:, line_width] = 10
            fh.write(body.tobytes())
            offset += body.size
            rem = n - n_full * line_width
            if rem:
                fh.write(arr[n_full * line_width:].tobytes() + b'\n')
                offset += rem + 1
    with open(path + '.fai', 'w') as fh:
        fh.writelines(fai_lines)
    return path


# ----------------------------------------------------------------------------------------------
# Path A: alignments with edits
# ----------------------------------------------------------------------------------------------

def _indel_lengths(rng, n, sv_frac):
    ln = np.minimum(rng.geometric(0.2, size=n), 60).astype(np.int64)
    big = rng.random(n) < sv_frac
    ln[big] = rng.integers(50, 5001, size=int(big.sum()))
    return ln


def plant_alignment(rng, ref, start, span, n_edit, grid=50, snv_frac=0.8, ins_frac=0.1,
                    x_run_frac=0.05, sv_frac=0.02, tr_frac=0.2, clip_l=0, clip_r=0, lower_ins=False):
    """Build one aligned query over ``ref[start:start+span]`` (``ref`` is modified in place where
    tandem repeats are planted).

    Returns ``(query uint8 array in reference orientation incl. clips, cigar str, ref_span)``.
    """
    n_cells = span // grid - 2
    n_edit = min(n_edit, max(n_cells, 0))
    cells = np.sort(rng.choice(n_cells, size=n_edit, replace=False)) + 1 if n_edit else np.zeros(0, np.int64)
    pos = cells.astype(np.int64) * grid + rng.integers(0, grid // 2, size=n_edit)

    u = rng.random(n_edit)
    kind = np.where(u < snv_frac, 0, np.where(u < snv_frac + ins_frac, 1, 2))  # 0 X, 1 I, 2 D
    ln = np.ones(n_edit, dtype=np.int64)
    xrun = (kind == 0) & (rng.random(n_edit) < x_run_frac)
    ln[xrun] = rng.integers(2, 4, size=int(xrun.sum()))
    is_indel = kind > 0
    ln[is_indel] = _indel_lengths(rng, int(is_indel.sum()), sv_frac)

    # Tandem repeats: plant unit x copies in the reference right at the site, the indel is k units
    # placed after j whole units (not left-aligned => exercises left-shift + wrap-around homology).
    tr = is_indel & (rng.random(n_edit) < tr_frac)
    tr_unit = rng.integers(1, 7, size=n_edit)
    tr_copies = rng.integers(5, 51, size=n_edit)
    tr_k = np.minimum(rng.integers(1, 4, size=n_edit), tr_copies - 1)
    tr_j = (rng.integers(0, 1 << 30, size=n_edit) % (tr_copies - tr_k + 1))
    tr_len = np.where(tr, tr_unit * tr_copies, 0)
    ln = np.where(tr, tr_unit * tr_k, ln)
    edit_pos = np.where(tr, pos + tr_unit * tr_j, pos)

    # Footprint of every edit on the reference; drop edits overlapping an earlier footprint.
    ref_use = np.where(kind == 1, 0, ln)
    foot_end = np.maximum(edit_pos + ref_use, pos + tr_len) + 2
    keep = np.ones(n_edit, dtype=bool)
    if n_edit:
        prev_end = np.concatenate(([0], np.maximum.accumulate(foot_end)[:-1]))
        keep = (pos > prev_end) & (foot_end < span - grid)
        # dropping an edit shrinks footprints, which is conservative (never creates overlaps)
    idx = np.flatnonzero(keep)

    seg = ref[start:start + span]
    for i in idx[tr[idx]]:  # plant repeat arrays (few; python loop is fine)
        unit = random_seq(rng, int(tr_unit[i]))
        seg[pos[i]:pos[i] + tr_len[i]] = np.tile(unit, int(tr_copies[i]))

    kind, ln, edit_pos, tr, tr_unit = kind[idx], ln[idx], edit_pos[idx], tr[idx], tr_unit[idx]
    n = len(idx)

    q = seg.copy()
    # substitutions
    xi = np.flatnonzero(kind == 0)
    if len(xi):
        sub_pos = np.repeat(edit_pos[xi], ln[xi]) + _ragged_arange(ln[xi])
        q[sub_pos] = _different_base(rng, seg[sub_pos])
    # deletions
    keep_mask = np.ones(span, dtype=bool)
    di = np.flatnonzero(kind == 2)
    if len(di):
        del_pos = np.repeat(edit_pos[di], ln[di]) + _ragged_arange(ln[di])
        keep_mask[del_pos] = False
    # insertions (inserted before reference position edit_pos)
    ii = np.flatnonzero(kind == 1)
    if len(ii):
        tot = int(ln[ii].sum())
        ins_vals = random_seq(rng, tot)
        off = np.concatenate(([0], np.cumsum(ln[ii])))
        for a, i in enumerate(ii):
            if tr[i]:  # inserted sequence = copies of the unit that follows the site
                unit = seg[edit_pos[i]:edit_pos[i] + tr_unit[i]]
                ins_vals[off[a]:off[a + 1]] = np.tile(unit, int(ln[i] // tr_unit[i]))
        if lower_ins:
            ins_vals = ins_vals | 0x20
        ins_at = np.repeat(edit_pos[ii], ln[ii])
        q = np.insert(q, ins_at, ins_vals)
        keep_mask = np.insert(keep_mask, ins_at, True)
    q = q[keep_mask]

    # CIGAR
    adv = np.where(kind == 1, 0, ln)
    after = edit_pos + adv
    eq = edit_pos - np.concatenate(([0], after[:-1])) if n else np.zeros(0, np.int64)
    tail = span - (after[-1] if n else 0)
    opc = np.array(['X', 'I', 'D'])[kind] if n else np.zeros(0, dtype='<U1')
    parts = []
    if clip_l:
        parts.append(f'{clip_l}H')
    eq_l, ln_l, opc_l = eq.tolist(), ln.tolist(), opc.tolist()
    for k in range(n):
        if eq_l[k] > 0:
            parts.append(f'{eq_l[k]}=')
        parts.append(f'{ln_l[k]}{opc_l[k]}')
    if tail > 0:
        parts.append(f'{tail}=')
    if clip_r:
        parts.append(f'{clip_r}H')
    cigar = ''.join(parts)

    if clip_l or clip_r:
        q = np.concatenate((random_seq(rng, clip_l), q, random_seq(rng, clip_r)))
    return q, cigar, span


def _ragged_arange(lengths):
    """Concatenated ``arange(l)`` for each l in lengths."""
    lengths = np.asarray(lengths, dtype=np.int64)
    tot = int(lengths.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.int64)
    starts = np.cumsum(lengths) - lengths
    return np.arange(tot, dtype=np.int64) - np.repeat(starts, lengths)


def make_cigar_workload(seed, n_chrom, chrom_len, n_contig, contig_len, edit_rate=0.01,
                        rev_frac=0.5, clip=(0, 0), soft_mask_frac=0.0, n_block_frac=0.0,
                        tr_frac=0.2, hap='h1', chrom_prefix='chr', contig_prefix='tig'):
    """Reference + contigs + alignment table. Contigs tile the chromosomes end to end.

    Returns ``(ref: dict name->uint8, tigs: dict name->uint8 (as stored in the contig FASTA, i.e.
    reverse-complemented when REV), df_align)``.
    """
    rng = np.random.default_rng(seed)
    ref = {f'{chrom_prefix}{c + 1}': random_seq(rng, chrom_len) for c in range(n_chrom)}
    names = list(ref.keys())
    per_chrom = chrom_len // contig_len
    tigs = {}
    rows = []
    n_edit = int(contig_len * edit_rate)
    for t in range(n_contig):
        chrom = names[(t // per_chrom) % n_chrom]
        start = (t % per_chrom) * contig_len
        rev = bool(rng.random() < rev_frac)
        q, cigar, span = plant_alignment(rng, ref[chrom], start, contig_len, n_edit,
                                         clip_l=clip[0], clip_r=clip[1], tr_frac=tr_frac)
        name = f'{contig_prefix}{t:05d}'
        qlen = len(q)
        tigs[name] = revcomp(q) if rev else q
        qpos, qend = clip[0], qlen - clip[1]
        rows.append((chrom, start, start + span, t, name,
                     qlen - qend if rev else qpos, qlen - qpos if rev else qend, qlen,
                     'NA', 'NA', 60, rev, '0x0010' if rev else '0x0000', hap, cigar))
    if n_block_frac > 0 or soft_mask_frac > 0:
        for chrom in names:
            _mask_runs(rng, ref[chrom], soft_mask_frac, n_block_frac)
    df = pd.DataFrame(rows, columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END',
                                     'QRY_LEN', 'RG', 'AO', 'MAPQ', 'REV', 'FLAGS', 'HAP', 'CIGAR'])
    df.sort_values(['#CHROM', 'POS', 'END', 'QRY_ID'], ascending=[True, True, False, True], inplace=True)
    return ref, tigs, df


def _mask_runs(rng, arr, soft_frac, n_frac, run=2000):
    """Lower-case (soft-mask) and N-out random runs of a chromosome, in place (hg38-shaped)."""
    n = len(arr)
    if soft_frac > 0:
        k = int(n * soft_frac / run)
        for s in rng.integers(0, max(n - run, 1), size=k):
            arr[s:s + run] |= 0x20
    if n_frac > 0:
        k = max(int(n * n_frac / run), 1)
        for s in rng.integers(0, max(n - run, 1), size=k):
            arr[s:s + run] = ord('N')


def config_c1(seed=1001):
    """BASELINE configs[0]: one 50 kbp contig vs 50 kbp reference, ~500 edits, ``2H...1H`` clips."""
    return make_cigar_workload(seed, 1, 50_000, 1, 50_000, edit_rate=0.01, rev_frac=0.0, clip=(2, 1))


def config_c2(seed=1002, n_contig=1000, contig_len=200_000, n_chrom=4):
    """BASELINE configs[1]: 1,000 x 200 kbp contigs vs 200 Mbp reference (4 x 50 Mbp)."""
    chrom_len = n_contig * contig_len // n_chrom
    return make_cigar_workload(seed, n_chrom, chrom_len, n_contig, contig_len, edit_rate=0.01, rev_frac=0.5)


def write_cigar_workload(out_dir, ref, tigs, df_align, prefix='wl'):
    os.makedirs(out_dir, exist_ok=True)
    ref_fa = write_fasta(os.path.join(out_dir, f'{prefix}_ref.fa'), ref)
    tig_fa = write_fasta(os.path.join(out_dir, f'{prefix}_tig.fa'), tigs)
    bed = os.path.join(out_dir, f'{prefix}_align.bed')
    df_align.to_csv(bed, sep='\t', index=False)
    return ref_fa, tig_fa, bed


# ----------------------------------------------------------------------------------------------
# Path B: inversion windows
# ----------------------------------------------------------------------------------------------

def make_inv_window(rng, win_len=50_000, inv_len=None, flank_rep=0, divergence=0.0, negative=False,
                    n_run=0):
    """One (reference window, contig window) pair with a central inversion.

    ``flank_rep`` > 0 plants an inverted repeat of that length on both sides of the inversion (gives
    FWDREV k-mer states). ``divergence`` applies SNVs to the contig after inverting. ``n_run`` puts
    a run of ``N`` into the contig (k-mer stream reset).
    """
    ref = random_seq(rng, win_len)
    if inv_len is None:
        inv_len = int(rng.integers(2000, 20001))
    margin = min(2000, win_len // 5)
    inv_len = max(min(inv_len, win_len - 2 * margin - 2 * flank_rep), 64)
    a = (win_len - inv_len) // 2
    b = a + inv_len
    if flank_rep:
        # right flank repeat = reverse complement of the left flank repeat (inverted repeat pair)
        ref[b:b + flank_rep] = revcomp(ref[a - flank_rep:a])
    tig = ref.copy()
    if not negative:
        tig[a:b] = revcomp(ref[a:b])
    if divergence > 0:
        m = rng.random(win_len) < divergence
        tig[m] = _different_base(rng, tig[m])
    if n_run:
        s = int(rng.integers(0, win_len - n_run))
        tig[s:s + n_run] = ord('N')
    return ref, tig, (a, b)


def make_inv_workload(seed=1005, n_win=16, win_len=50_000, flank_frac=0.3, divergence=0.005,
                      neg_frac=0.1):
    """BASELINE configs[4] shape: ``n_win`` windows, one reference record + one contig record each.

    Returns ``(ref dict, tig dict, list of (ref_name, tig_name, inv interval, negative))``.
    """
    rng = np.random.default_rng(seed)
    ref, tig, meta = {}, {}, []
    for w in range(n_win):
        neg = bool(rng.random() < neg_frac)
        flank = int(rng.integers(1000, 3001)) if rng.random() < flank_frac else 0
        r, t, iv = make_inv_window(rng, win_len, None, flank, divergence, neg)
        ref[f'rw{w:05d}'] = r
        tig[f'tw{w:05d}'] = t
        meta.append((f'rw{w:05d}', f'tw{w:05d}', iv, neg))
    return ref, tig, meta
