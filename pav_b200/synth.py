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


TR_SPACING = 2500   # one synthetic tandem-repeat array per TR_SPACING bp of reference
TR_OFFSET = 100
TR_MAX_LEN = 6 * 50


def make_reference(seed, n_chrom, chrom_len, chrom_prefix='chr', tandem_repeats=True):
    """Reference chromosomes (iid uniform ACGT) with tandem-repeat arrays (unit 1-6 bp x 5-50 copies)
    planted every ``TR_SPACING`` bp. Depends on ``seed`` only, so every rank of a multi-GPU run
    regenerates the identical reference. ``chrom_len`` is one length for all chromosomes or a sequence of
    ``n_chrom`` lengths. Returns ``(ref dict, tr dict: chrom -> (pos, unit_len, copies))``.
    """
    rng = np.random.default_rng([seed, 0xA11])
    ref, trs = {}, {}
    lens = [int(chrom_len)] * n_chrom if np.isscalar(chrom_len) else [int(x) for x in chrom_len]
    for c in range(n_chrom):
        name = f'{chrom_prefix}{c + 1}'
        chrom_len = lens[c]
        arr = random_seq(rng, chrom_len)
        n_tr = max((chrom_len - TR_OFFSET - TR_MAX_LEN) // TR_SPACING, 0) if tandem_repeats else 0
        pos = np.arange(n_tr, dtype=np.int64) * TR_SPACING + TR_OFFSET
        unit = rng.integers(1, 7, size=n_tr)
        copies = rng.integers(5, 51, size=n_tr)
        units = random_seq(rng, 6 * n_tr).reshape(n_tr, 6) if n_tr else np.zeros((0, 6), np.uint8)
        # fill arrays: base t of array i is units[i, t % unit[i]]
        tot = unit * copies
        idx = _ragged_arange(tot)
        owner = np.repeat(np.arange(n_tr), tot)
        arr[np.repeat(pos, tot) + idx] = units[owner, idx % np.repeat(unit, tot)]
        ref[name] = arr
        trs[name] = (pos, unit, copies)
    return ref, trs


def plant_alignment(rng, ref, tr, start, span, n_edit, grid=50, snv_frac=0.8, ins_frac=0.1,
                    x_run_frac=0.05, sv_frac=0.02, tr_frac=0.2, clip_l=0, clip_r=0, lower_ins=False):
    """Build one aligned query over ``ref[start:start+span]`` (``ref`` is read-only here).

    ``tr`` = ``(pos, unit_len, copies)`` of the tandem-repeat arrays of this chromosome; a fraction
    ``tr_frac`` of the indels is placed inside such arrays, k whole units after j whole units (not
    left-aligned, which exercises left-shift and wrap-around homology).

    Returns ``(query uint8 array in reference orientation incl. clips, cigar str, ref_span)``.
    """
    seg = ref[start:start + span]
    n_cells = span // grid - 2
    n_edit = min(n_edit, max(n_cells, 0))
    u = rng.random(n_edit)
    kind = np.where(u < snv_frac, 0, np.where(u < snv_frac + ins_frac, 1, 2))  # 0 X, 1 I, 2 D
    is_indel = kind > 0
    want_tr = is_indel & (rng.random(n_edit) < tr_frac)

    # tandem-repeat arrays fully inside the span
    tpos, tunit, tcopies = tr if tr is not None else (np.zeros(0, np.int64),) * 3
    inside = np.flatnonzero((tpos >= start + grid) & (tpos + tunit * tcopies <= start + span - grid)) if len(tpos) else np.zeros(0, np.int64)
    n_tr = min(int(want_tr.sum()), len(inside))
    tr_idx = np.flatnonzero(want_tr)[:n_tr]
    arrays = rng.choice(inside, size=n_tr, replace=False) if n_tr else np.zeros(0, np.int64)
    is_tr = np.zeros(n_edit, dtype=bool)
    is_tr[tr_idx] = True
    want_tr &= is_tr

    cells = (np.sort(rng.choice(n_cells, size=n_edit, replace=False)) + 1) if n_edit else np.zeros(0, np.int64)
    pos = cells.astype(np.int64) * grid + rng.integers(0, grid // 2, size=n_edit)
    ln = np.ones(n_edit, dtype=np.int64)
    xrun = (kind == 0) & (rng.random(n_edit) < x_run_frac)
    ln[xrun] = rng.integers(2, 4, size=int(xrun.sum()))
    ln[is_indel] = _indel_lengths(rng, int(is_indel.sum()), sv_frac)

    tr_unit = np.ones(n_edit, dtype=np.int64)
    tr_len = np.zeros(n_edit, dtype=np.int64)
    edit_pos = pos.copy()
    if n_tr:
        a_pos, a_unit, a_cop = tpos[arrays] - start, tunit[arrays], tcopies[arrays]
        k = np.minimum(rng.integers(1, 4, size=n_tr), a_cop - 1)
        j = rng.integers(0, 1 << 30, size=n_tr) % (a_cop - k + 1)
        pos[tr_idx] = a_pos
        edit_pos[tr_idx] = a_pos + a_unit * j
        ln[tr_idx] = a_unit * k
        tr_unit[tr_idx] = a_unit
        tr_len[tr_idx] = a_unit * a_cop
    order = np.argsort(pos, kind='stable')
    pos, edit_pos, kind, ln, is_tr, tr_unit, tr_len = (x[order] for x in (pos, edit_pos, kind, ln, is_tr, tr_unit, tr_len))

    # Footprint of every edit on the reference; drop edits overlapping an earlier footprint.
    ref_use = np.where(kind == 1, 0, ln)
    foot_end = np.maximum(edit_pos + ref_use, pos + tr_len) + 2
    if n_edit:
        prev_end = np.concatenate(([0], np.maximum.accumulate(foot_end)[:-1]))
        keep = (pos > prev_end) & (foot_end < span - grid)
        # dropping an edit shrinks footprints, which is conservative (never creates overlaps)
    else:
        keep = np.ones(0, dtype=bool)
    idx = np.flatnonzero(keep)
    kind, ln, edit_pos, tr, tr_unit = kind[idx], ln[idx], edit_pos[idx], is_tr[idx], tr_unit[idx]
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
        for a in np.flatnonzero(tr[ii]).tolist():  # inserted sequence = copies of the unit that follows the site
            i = ii[a]
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


def make_contigs(ref, trs, hap_seed, n_contig, contig_len, edit_rate=0.01, rev_frac=0.5, clip=(0, 0), tr_frac=0.2,
                 hap='h1', contig_prefix='tig', index_base=0):
    """Contigs tiling the chromosomes end to end + their alignment table (one haplotype)."""
    rng = np.random.default_rng([hap_seed, 0xC16])
    names = list(ref.keys())
    chrom_len = len(ref[names[0]])
    per_chrom = max(chrom_len // contig_len, 1)
    tigs = {}
    rows = []
    n_edit = int(contig_len * edit_rate)
    for t in range(n_contig):
        chrom = names[(t // per_chrom) % len(names)]
        start = (t % per_chrom) * contig_len
        rev = bool(rng.random() < rev_frac)
        q, cigar, span = plant_alignment(rng, ref[chrom], trs.get(chrom) if trs else None, start, contig_len, n_edit,
                                         clip_l=clip[0], clip_r=clip[1], tr_frac=tr_frac)
        name = f'{contig_prefix}{t:05d}'
        qlen = len(q)
        tigs[name] = revcomp(q) if rev else q
        qpos, qend = clip[0], qlen - clip[1]
        rows.append((chrom, start, start + span, index_base + t, name,
                     qlen - qend if rev else qpos, qlen - qpos if rev else qend, qlen,
                     'NA', 'NA', 60, rev, '0x0010' if rev else '0x0000', hap, cigar))
    df = pd.DataFrame(rows, columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END',
                                     'QRY_LEN', 'RG', 'AO', 'MAPQ', 'REV', 'FLAGS', 'HAP', 'CIGAR'])
    df.sort_values(['#CHROM', 'POS', 'END', 'QRY_ID'], ascending=[True, True, False, True], inplace=True)
    return tigs, df


# hg38 primary assembly, chr1..chr22, chrX, chrY (sum 3.09 Gbp) -- the shape of BASELINE configs[2]/[3]
HG38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
                135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
                46709983, 50818468, 156040895, 57227415]

# edit regimes of SURVEY 8(d) for C3/C4: (edits per base, SNV fraction, INS fraction)
C3_REGIMES = {'human': (1.2e-3, 1.0 / 1.2, 0.1 / 1.2), 'stress': (1e-2, 0.8, 0.1)}


def make_contigs_tiled(ref, trs, hap_seed, contig_len, edit_rate=0.01, snv_frac=0.8, ins_frac=0.1, rev_frac=0.5, tr_frac=0.2,
                       hap='h1', contig_prefix='tig', min_len=20_000):
    """Contigs of ``contig_len`` tiling every chromosome end to end (the last one of a chromosome is shorter; tails under
    ``min_len`` are left uncovered) + their alignment table. Chromosomes may have different lengths (hg38 shape)."""
    rng = np.random.default_rng([hap_seed, 0xC17])
    tigs, rows, t = {}, [], 0
    for chrom, arr in ref.items():
        for start in range(0, len(arr), contig_len):
            span = min(contig_len, len(arr) - start)
            if span < min_len:
                break
            rev = bool(rng.random() < rev_frac)
            q, cigar, span = plant_alignment(rng, arr, trs.get(chrom) if trs else None, start, span, int(span * edit_rate),
                                             snv_frac=snv_frac, ins_frac=ins_frac, tr_frac=tr_frac)
            name = f'{contig_prefix}{t:05d}'
            qlen = len(q)
            tigs[name] = revcomp(q) if rev else q
            rows.append((chrom, start, start + span, t, name, 0, qlen, qlen, 'NA', 'NA', 60, rev, '0x0010' if rev else '0x0000', hap, cigar))
            t += 1
    df = pd.DataFrame(rows, columns=['#CHROM', 'POS', 'END', 'INDEX', 'QRY_ID', 'QRY_POS', 'QRY_END',
                                     'QRY_LEN', 'RG', 'AO', 'MAPQ', 'REV', 'FLAGS', 'HAP', 'CIGAR'])
    df.sort_values(['#CHROM', 'POS', 'END', 'QRY_ID'], ascending=[True, True, False, True], inplace=True)
    return tigs, df.reset_index(drop=True)


def config_c3_reference(seed=1003, scale=1.0, soft_mask_frac=0.5, n_block_frac=0.05):
    """BASELINE configs[2] reference: 24 chromosomes with hg38 primary lengths (x ``scale``), 50 % soft-masked runs, 5 % of the
    bases in N blocks. Returns ``(ref, trs, ref_pristine_is_masked)`` -- the contigs are derived from the reference *before*
    masking (use ``config_c3_haplotype``), as real assemblies carry sequence where the reference has N."""
    lens = [max(int(x * scale), 40_000) for x in HG38_LENGTHS]
    return make_reference(seed, len(lens), lens)


def config_c3_mask(ref, seed=1003, soft_mask_frac=0.5, n_block_frac=0.05):
    rng = np.random.default_rng([seed, 0x3A5])
    for chrom in ref:
        _mask_runs(rng, ref[chrom], soft_mask_frac, n_block_frac)


def config_c3_haplotype(ref, trs, hap, seed=1003, scale=1.0, regime='human', contig_len=10_000_000):
    """One haplotype of BASELINE configs[2]: 10 Mbp contigs (x ``scale``) tiling the (still unmasked) reference."""
    rate, snv_frac, ins_frac = C3_REGIMES[regime]
    hap_seed = seed * 10 + (1 if hap == 'h1' else 2)
    return make_contigs_tiled(ref, trs, hap_seed, max(int(contig_len * scale), 40_000), edit_rate=rate, snv_frac=snv_frac,
                              ins_frac=ins_frac, hap=hap, contig_prefix=f'{hap}tig')


def make_cigar_workload(seed, n_chrom, chrom_len, n_contig, contig_len, edit_rate=0.01,
                        rev_frac=0.5, clip=(0, 0), soft_mask_frac=0.0, n_block_frac=0.0,
                        tr_frac=0.2, hap='h1', chrom_prefix='chr', contig_prefix='tig', hap_seed=None):
    """Reference + contigs + alignment table. The reference depends on ``seed`` only; the contigs on
    ``hap_seed`` (default ``seed``), so ranks of a multi-GPU run share one reference.

    Returns ``(ref: dict name->uint8, tigs: dict name->uint8 (as stored in the contig FASTA, i.e.
    reverse-complemented when REV), df_align)``.
    """
    ref, trs = make_reference(seed, n_chrom, chrom_len, chrom_prefix)
    tigs, df = make_contigs(ref, trs, seed if hap_seed is None else hap_seed, n_contig, contig_len, edit_rate, rev_frac,
                            clip, tr_frac, hap, contig_prefix)
    if n_block_frac > 0 or soft_mask_frac > 0:
        rng = np.random.default_rng([seed, 0x3A5])
        for chrom in ref:
            _mask_runs(rng, ref[chrom], soft_mask_frac, n_block_frac)
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


def config_c2(seed=1002, n_contig=1000, contig_len=200_000, n_chrom=4, hap_seed=None):
    """BASELINE configs[1]: 1,000 x 200 kbp contigs vs 200 Mbp reference (4 x 50 Mbp)."""
    chrom_len = n_contig * contig_len // n_chrom
    return make_cigar_workload(seed, n_chrom, chrom_len, n_contig, contig_len, edit_rate=0.01, rev_frac=0.5, hap_seed=hap_seed)


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


# ----------------------------------------------------------------------------------------------
# SAM text (input of pavlib.align.get_align_bed)
# ----------------------------------------------------------------------------------------------

def write_sam(path, df_align, ref, soft_clip_every=2, extra_lines=()):
    """Write the alignment table as SAM text (SEQ/QUAL omitted). Every ``soft_clip_every``-th record gets its hard
    clips rewritten as soft clips so that readers exercise soft->hard conversion. ``extra_lines`` are appended verbatim
    (unmapped / filtered records)."""
    with open(path, 'w') as fh:
        fh.write('@HD\tVN:1.6\tSO:unsorted\n')
        for name, arr in ref.items():
            fh.write(f'@SQ\tSN:{name}\tLN:{len(arr)}\n')
        for i, (_, row) in enumerate(df_align.iterrows()):
            cigar = row['CIGAR']
            if soft_clip_every and i % soft_clip_every == 1:
                cigar = cigar.replace('H', 'S')
            flag = 16 if row['REV'] else 0
            tags = ['RG:Z:grp1'] if i % 3 == 0 else []
            if i % 4 == 0:
                tags.append('AO:i:%d' % i)
            fh.write('\t'.join([row['QRY_ID'], str(flag), row['#CHROM'], str(int(row['POS']) + 1), str(int(row['MAPQ'])), cigar, '*', '0', '0',
                                '*', '*'] + tags) + '\n')
        for line in extra_lines:
            fh.write(line.rstrip('\n') + '\n')
    return path


def write_bgzf(path, data, block=0xff00, write_gzi=True):
    """Write ``data`` (bytes) as a BGZF file (the container bgzip produces) and, optionally, its ``.gzi`` index."""
    import struct
    import zlib
    entries = []
    c_off = u_off = 0
    with open(path, 'wb') as fh:
        for i in range(0, max(len(data), 1), block):
            chunk = data[i:i + block]
            if i > 0:
                entries.append((c_off, u_off))
            comp = zlib.compressobj(6, zlib.DEFLATED, -15)
            body = comp.compress(chunk) + comp.flush()
            bsize = len(body) + 25
            blk = (b'\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00' + struct.pack('<H', bsize) + body +
                   struct.pack('<II', zlib.crc32(chunk) & 0xffffffff, len(chunk)))
            fh.write(blk)
            c_off += len(blk)
            u_off += len(chunk)
        fh.write(bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000'))  # EOF marker block
    if write_gzi:
        with open(path + '.gzi', 'wb') as fh:
            fh.write(struct.pack('<Q', len(entries)))
            for c, u in entries:
                fh.write(struct.pack('<QQ', c, u))
    return path
