"""Inversion scan over flagged loci with the reference's interface (pavlib/inv.py).

``scan_for_inv`` keeps the reference's control flow (expand the region x1.5, balance .25/.5/.75,
stop rules, breakpoint derivation, size-proportion test, flank annotation: pavlib/inv.py:149-454) but
each expansion scores its window through ``pavgpu_density_batch_*`` in-process instead of spawning
``python3 scripts/density.py`` and unpickling its stdout (pavlib/inv.py:249-288).
"""
import numpy as np
import pandas as pd

from . import density as pavdensity
from . import seq as pavseq
from .constants import ERR_INV_FAIL
from .. import fasta

INITIAL_EXPAND = 4000
EXPAND_FACTOR = 1.5
MAX_REGION_SIZE = 1200000
MIN_INFORMATIVE_KMERS = 2000
MIN_KMER_STATE_COUNT = 20
DENSITY_SMOOTH_FACTOR = 1
MIN_INV_KMER_RUN = 100
MIN_QRY_REF_PROP = 0.6
DEFAULT_MIN_EXP_COUNT = 1
DEFAULT_STATE_RUN_SMOOTH = 20
CALL_SOURCE = 'FLAG-DEN'

# KMER_LOC_STATE[in-upstream, in-dnstream] (pavlib/inv.py:46-51)
KMER_LOC_STATE = np.asarray([['NA', 'OTHER'], ['SAME', 'NA']])


class InvCall:
    """An inversion call and the density table that supports it (reference: pavlib/inv.py:54-118)."""

    def __init__(self, region_ref_outer, region_ref_inner, region_tig_outer, region_tig_inner,
                 region_ref_discovery, region_tig_discovery, region_flag, df):
        self.region_ref_outer = region_ref_outer
        self.region_ref_inner = region_ref_inner
        self.region_tig_outer = region_tig_outer
        self.region_tig_inner = region_tig_inner
        self.region_ref_discovery = region_ref_discovery
        self.region_tig_discovery = region_tig_discovery
        self.region_flag = region_flag
        self._df = df           # the density table, or a zero-argument callable that builds it when somebody asks (batch driver)
        self.svlen = len(region_ref_outer)
        self.id = '{}-{}-INV-{}'.format(region_ref_outer.chrom, region_ref_outer.pos + 1, self.svlen)

    @property
    def df(self):
        """Density table with the FLANK / MATCH columns (pavlib/inv.py:440-454). The batch driver hands over the column arrays
        and builds the frame on first access: a caller that only wants the regions never pays for it (1.2 ms of pandas per call)."""
        if callable(self._df):
            self._df = self._df()
        return self._df

    @df.setter
    def df(self, value):
        self._df = value

    def __repr__(self):
        return self.id


class _Span:
    __slots__ = ('begin', 'end', 'data')

    def __init__(self, begin, end, data):
        self.begin, self.end, self.data = begin, end, data


class SrsTree:
    """Minimal stand-in for the intervaltree the reference uses for --staterunsmooth by region size:
    supports ``tree[a:b] = value`` and ``tree[point]`` -> set of objects with ``.data``."""

    def __init__(self):
        self._spans = []

    def __setitem__(self, key, value):
        self._spans.append(_Span(key.start, key.stop, value))

    def __getitem__(self, point):
        return {s for s in self._spans if s.begin <= point < s.end}

    def __len__(self):
        return len(self._spans)


def get_srs_tree(srs_tuple_list):
    """(region size limit, state-run-smooth) tuples -> lookup structure (reference: pavlib/inv.py:564-620)."""
    tree = SrsTree()
    if srs_tuple_list is None or len(srs_tuple_list) == 0:
        tree[0:np.inf] = DEFAULT_STATE_RUN_SMOOTH
        return tree
    for el in srs_tuple_list:
        if len(el) != 2:
            raise RuntimeError('Element in "state run smooth" tuple list that is not length 2: ' + str(el))
    srs_tuple_list = sorted(srs_tuple_list)
    last_lim, last_smooth = int(srs_tuple_list[0][0]), int(srs_tuple_list[0][1])
    if last_lim < 0:
        raise RuntimeError('State run inversion size limits must be 0 or greater: {}'.format(last_lim))
    if last_smooth < 4:
        raise RuntimeError('Not tested with "state run smooth" factor less than 4: {}'.format(last_smooth))
    if last_lim > 0:
        tree[0:last_lim] = np.min([last_lim, 20])
    for lim, smooth in srs_tuple_list[1:]:
        lim, smooth = int(lim), int(smooth)
        if smooth < 20:
            raise RuntimeError('Not tested with "state run smooth" factor less than 20: {}'.format(smooth))
        if lim == last_lim:
            raise RuntimeError('Duplicate limit in state run limits: {}'.format(lim))
        tree[last_lim:lim] = last_smooth
        last_lim, last_smooth = lim, smooth
    tree[last_lim:np.inf] = last_smooth
    return tree


def _write_log(message, log):
    if log is None:
        return
    log.write(message)
    log.write('\n')
    log.flush()


class _InvScan:
    """State machine of one flagged locus: the body of the reference's ``scan_for_inv`` loop (pavlib/inv.py:149-454) cut
    at the point where it needs a density table, so one locus (``scan_for_inv``) or many loci per GPU batch
    (``scan_for_inv_batch``) share the same control flow, log lines and stop rules."""

    def __init__(self, region_flag, ref_fa_name, tig_fa_name, align_lift, k_util, n_tree, max_region_size, log, srs_tree, min_exp_count,
                 df_fai=None):
        self.region_flag = region_flag
        # Regions handed back inside the InvCall are built from the CALLER's Region class: a rule that passes the reference's
        # pavlib.seq.Region goes on to hand them to the reference's own functions, which test `region.__class__ == Region`
        # (pavlib/seq.py:341-346, region_seq_fasta in rule call_inv_batch)
        self.region_cls = type(region_flag) if hasattr(region_flag, 'to_base1_string') else pavseq.Region
        self.ref_fa_name, self.tig_fa_name = ref_fa_name, tig_fa_name
        self.align_lift, self.k_util, self.log = align_lift, k_util, log
        self.k_size = int(k_util.k_size)
        self.min_exp_count = DEFAULT_MIN_EXP_COUNT if min_exp_count is None else min_exp_count
        self.max_region_size = MAX_REGION_SIZE if max_region_size is None else max_region_size
        _write_log('Scanning for inversions in flagged region: {} (flagged region record id = {})'.format(
            region_flag, region_flag.region_id()), log)
        self.df_fai = pavseq.get_df_fai(ref_fa_name + '.fai') if df_fai is None else df_fai
        self.region_ref = region_flag.copy()
        self.region_ref.expand(INITIAL_EXPAND, min_pos=0, max_end=self.df_fai, shift=True)
        self.expansion_count = 0
        self.n_tree_chrom = n_tree[self.region_ref.chrom] if (n_tree is not None and self.region_ref.chrom in n_tree.keys()) else None
        if srs_tree is None:
            srs_tree = get_srs_tree(None)
        elif not hasattr(srs_tree, '__getitem__'):
            raise NotImplementedError('Custom state-run-smooth parameters are not currently implemented')
        self.srs_tree = srs_tree
        self.region_tig = None
        self._prelift = None
        self.done = False
        self.result = None

    def _finish(self, result):
        self.done, self.result = True, result
        return None

    def next_window(self):
        """Next ``(region_ref, region_tig, rev, srs)`` to score, or ``None`` when the scan has ended (see ``result``)."""
        log, region_ref = self.log, self.region_ref
        if 0 < self.max_region_size < len(region_ref):
            _write_log('Region size exceeds max: {} ({} > {})'.format(region_ref, len(region_ref), self.max_region_size), log)
            return self._finish(None)
        if self.n_tree_chrom is not None and len(self.n_tree_chrom[region_ref.pos:region_ref.end]) > 0:
            _write_log('Region overlaps N bases: {}'.format(region_ref), log)  # logged only, as in the reference
        if self._prelift is not None:      # the batch driver lifted the regions of all open loci in one device call
            self.region_tig, self._prelift = self._prelift[0], None
        else:
            self.region_tig = self.align_lift.lift_region_to_qry(region_ref)
        if self.region_tig is None:
            _write_log('Could not lift reference region onto contigs: {}'.format(region_ref), log)
            return self._finish(None)
        self.expansion_count += 1
        _write_log('Scanning region: {}'.format(region_ref), log)
        srs = list(self.srs_tree[len(self.region_tig)])[0].data
        return region_ref, self.region_tig, bool(self.region_tig.is_rev), int(srs)

    def feed(self, returncode, df):
        """Consume the density table of the window handed out last; afterwards either ``done`` or ready for ``next_window``."""
        log, region_ref = self.log, self.region_ref
        if returncode != 0:
            _write_log('Received return code {} from the density scan for region {}:\n'.format(returncode, str(region_ref)), log)
            if returncode != ERR_INV_FAIL:
                raise RuntimeError('Density scan failed with code {} for region {}'.format(returncode, region_ref))
            return self._finish(None)
        if df.shape[0] == 0:
            _write_log('No informative reference k-mers in forward or reverse orientation in region', log)
            return self._finish(None)
        lazy = isinstance(df, pavdensity.DensityTable)      # batch driver: the frame is only built when a call is characterised
        state_rl = [record for record in (df.rl() if lazy else pavdensity.rl_encoder(df))]
        condensed_states = [record[0] for record in state_rl]
        if len(state_rl) == 1 and state_rl[0][0] in {0, -1} and self.expansion_count >= self.min_exp_count:
            _write_log('Found no inverted k-mer states after {} expansion(s)'.format(self.expansion_count), log)
            return self._finish(None)
        if len(condensed_states) > 2 and condensed_states[0] == 0 and condensed_states[-1] == 0:
            return self._finish(self._characterise(df, state_rl))
        last_len = len(region_ref)
        expand_bp = np.int32(len(region_ref) * EXPAND_FACTOR)
        if len(condensed_states) > 2:
            balance = 0.25 if condensed_states[0] == 0 else (0.75 if condensed_states[-1] == 0 else 0.5)
        else:
            balance = 0.5
        region_ref.expand(expand_bp, min_pos=0, max_end=self.df_fai, shift=True, balance=balance)
        if len(region_ref) == last_len:
            _write_log('Reached reference limits, cannot expand', log)
            return self._finish(None)
        return None

    def _characterise(self, df, state_rl):
        """Breakpoints, size-proportion test, flank annotation (pavlib/inv.py:346-454)."""
        log, region_ref, region_tig, k_size, align_lift = self.log, self.region_ref, self.region_tig, self.k_size, self.align_lift
        if not np.any([record[0] == 2 for record in state_rl]):
            _write_log('No inverted states found', log)
            return None
        max_inv_run = np.max([record[1] for record in state_rl if record[0] == 2])
        if max_inv_run < MIN_INV_KMER_RUN:
            _write_log('Longest run of strictly inverted k-mers ({}) does not meet the minimum threshold ({})'.format(
                max_inv_run, MIN_INV_KMER_RUN), log)
            return None
        if state_rl[0][0] != 0 or state_rl[-1][0] != 0:
            raise RuntimeError('Found INV region not flanked by reference sequence (program bug): {}'.format(region_ref))
        state_rl_inv = [record for record in state_rl if record[0] == 2]
        region_tig_outer = self.region_cls(region_tig.chrom, state_rl[1][2] + region_tig.pos,
                                         state_rl[-2][3] + region_tig.pos + k_size, is_rev=region_tig.is_rev)
        region_tig_inner = self.region_cls(region_tig.chrom, state_rl_inv[0][2] + region_tig.pos,
                                         state_rl_inv[-1][3] + region_tig.pos + k_size, is_rev=region_tig.is_rev)
        region_ref_outer = align_lift.lift_region_to_sub(region_tig_outer)
        if region_ref_outer is None:
            _write_log('Failed lifting outer INV region to reference: {}'.format(region_tig_outer), log)
            return None
        region_ref_inner = align_lift.lift_region_to_sub(region_tig_inner, gap=True)
        if region_ref_inner is None:
            region_ref_inner = region_ref_outer
        print('INV Found: outer={}, inner={} (ref outer={}, inner={})'.format(
            region_tig_outer, region_tig_inner, region_ref_outer, region_ref_inner))
        if len(region_ref_outer) < len(region_tig_outer) * MIN_QRY_REF_PROP:
            _write_log('Reference region too short: Reference region length ({:,d}) is not within {:.2f}% of the contig region length ({:,d})'.format(
                len(region_ref_outer), MIN_QRY_REF_PROP * 100, len(region_tig_outer)), log)
            return None
        if len(region_tig_outer) < len(region_ref_outer) * MIN_QRY_REF_PROP:
            _write_log('Contig region too short: Contig region length ({:,d}) is not within {:.2f}% of the reference region length ({:,d})'.format(
                len(region_tig_outer), MIN_QRY_REF_PROP * 100, len(region_ref_outer)), log)
            return None
        # NOTE: the reference passes region_ref where annotate_inv_dup_mers expects the contig discovery region
        # (pavlib/inv.py:440-442); reproduced as is.
        if isinstance(df, pavdensity.DensityTable):   # batch driver: the frame is built once, with the two flank columns already in it,
            table, res = df, df.res                   # and only when the call's table is asked for (InvCall.df)
            res['INDEX']                              # a lazy window copies its columns off the device now: the batch does not outlive the round
            ref_fa_name, k_util = self.ref_fa_name, self.k_util

            def df():
                flank, match = _dup_mer_columns(res['INDEX'], res['KMER'].astype(np.int64), region_ref_outer, region_ref_inner, region_tig_outer,
                                                region_tig_inner, region_ref, ref_fa_name, k_util)
                return table.frame(extra={'FLANK': flank, 'MATCH': match})
        else:
            df = annotate_inv_dup_mers(df, region_ref_outer, region_ref_inner, region_tig_outer, region_tig_inner, region_ref,
                                       self.ref_fa_name, self.k_util)
        inv_call = InvCall(region_ref_outer, region_ref_inner, region_tig_outer, region_tig_inner, region_ref, region_tig,
                           self.region_flag, df)
        _write_log('Found inversion: {}'.format(inv_call), log)
        return inv_call


def scan_for_inv(region_flag, ref_fa_name, tig_fa_name, align_lift, k_util, n_tree=None, max_region_size=None, threads=1,
                 log=None, srs_tree=None, min_exp_count=DEFAULT_MIN_EXP_COUNT):
    """
    Scan a flagged region for an inversion, expanding as necessary.

    Same parameters and return value as the reference (``InvCall`` or ``None``, pavlib/inv.py:149-454). ``k_util`` only
    needs a ``k_size`` attribute; ``threads`` is accepted and ignored (the window is scored on the GPU); ``align_lift``
    needs ``lift_region_to_qry`` / ``lift_region_to_sub`` (the reference's AlignLift or ``pav_b200.pavlib.lift.AlignLift``).
    Each expansion scores its window through ``pavgpu_density_batch_*`` in-process instead of spawning
    ``python3 scripts/density.py`` and unpickling its stdout (pavlib/inv.py:249-288).
    """
    scan = _InvScan(region_flag, ref_fa_name, tig_fa_name, align_lift, k_util, n_tree, max_region_size, log, srs_tree, min_exp_count)
    while not scan.done:
        win = scan.next_window()
        if win is None:
            break
        region_ref, region_tig, rev, srs = win
        returncode, df = pavdensity.density_table(
            region_ref, region_tig, ref_fa_name, tig_fa_name, k=scan.k_size, rev=rev, state_run_smooth=srs,
            min_informative=MIN_INFORMATIVE_KMERS, min_state_count=MIN_KMER_STATE_COUNT, smooth=DENSITY_SMOOTH_FACTOR)
        scan.feed(returncode, df)
    return scan.result


def scan_for_inv_batch(region_flags, ref_fa_name, tig_fa_name, align_lift, k_util, n_tree=None, max_region_size=None, log=None,
                       srs_tree=None, min_exp_count=DEFAULT_MIN_EXP_COUNT, catch=False):
    """
    ``scan_for_inv`` for many flagged loci at once (extension; the reference scans one locus per call,
    rules/call_inv.snakefile:191-196). All loci that still need a density table are scored together in one GPU batch per
    expansion round, so a Snakemake batch of thousands of 50 kbp windows costs a handful of launches instead of one
    ``scripts/density.py`` process per window and expansion.

    The log lines of every locus are buffered and written to ``log`` locus by locus, so the log reads as if the loci had been
    scanned one after the other. ``catch=True``: a ``RuntimeError`` raised while scanning a locus becomes that locus' result
    instead of ending the batch (``rule call_inv_batch`` logs it and goes on, rules/call_inv.snakefile:198-200).

    :return: list with one ``InvCall`` or ``None`` per flagged region, identical to calling ``scan_for_inv`` on each.
    """
    import io

    from .. import fasta as _fasta
    df_fai = pavseq.get_df_fai(ref_fa_name + '.fai')
    logs = [io.StringIO() if log is not None else None for _ in region_flags]

    class _Log:   # StringIO without flush noise
        def __init__(self, buf):
            self.buf = buf

        def write(self, text):
            self.buf.write(text)

        def flush(self):
            pass
    scans = []
    for r, lg in zip(region_flags, logs):
        try:
            scans.append(_InvScan(r, ref_fa_name, tig_fa_name, align_lift, k_util, n_tree, max_region_size, None if lg is None else _Log(lg),
                                  srs_tree, min_exp_count, df_fai=df_fai))
        except RuntimeError as ex:
            if not catch:
                raise
            scans.append(_Failed(ex))
    ref_fa, tig_fa = _fasta.open_fasta(ref_fa_name), _fasta.open_fasta(tig_fa_name)
    k_size = int(k_util.k_size)
    batched_lift = getattr(align_lift, 'lift_regions_to_qry', None)
    while True:
        pending, windows = [], []
        open_scans = [sc for sc in scans if not sc.done]
        if batched_lift is not None and len(open_scans) > 1:
            try:     # both ends of every open locus' region in one device call (pavgpu_lift_points)
                for sc, r in zip(open_scans, batched_lift([sc.region_ref for sc in open_scans])):
                    sc._prelift = [r]
            except RuntimeError:   # some locus cannot be lifted: every locus meets its own error in its own turn below
                for sc in open_scans:
                    sc._prelift = None
        for sc in scans:
            if sc.done:
                continue
            try:
                win = sc.next_window()
            except RuntimeError as ex:
                if not catch:
                    raise
                sc.done, sc.result, win = True, ex, None
            if win is None:
                continue
            region_ref, region_tig, rev, srs = win
            windows.append((ref_fa.fetch_array(region_ref.chrom, region_ref.pos, region_ref.end),
                            tig_fa.fetch_array(region_tig.chrom, region_tig.pos, region_tig.end), rev, srs))
            pending.append(sc)
        if not pending:
            break
        # lazy: run lengths of STATE for every window; full columns only for the windows that become calls
        results = pavdensity.density_windows(windows, k=k_size, lazy=True, min_informative=MIN_INFORMATIVE_KMERS,
                                             min_state_count=MIN_KMER_STATE_COUNT, smooth=DENSITY_SMOOTH_FACTOR)
        for sc, res in zip(pending, results):
            try:
                if res['status'] != 0:
                    sc.feed(ERR_INV_FAIL, None)
                else:
                    sc.feed(0, pavdensity.DensityTable(res))
            except RuntimeError as ex:
                if not catch:
                    raise
                sc.done, sc.result = True, ex
    if log is not None:
        for lg, sc in zip(logs, scans):
            log.write(lg.getvalue())
            if isinstance(sc.result, RuntimeError):
                log.write('RuntimeError in scan_for_inv(): {}\n'.format(sc.result))
        log.flush()
    return [sc.result for sc in scans]


class _Failed:
    """A locus whose scan could not even start (``catch=True``)."""

    def __init__(self, ex):
        self.done, self.result = True, ex


_CODE = np.full(256, 255, dtype=np.uint8)
for _c, _v in zip(b'ACGTacgt', (0, 1, 2, 3, 0, 1, 2, 3)):
    _CODE[_c] = _v


def _region_canonical_kmers(region, fa_name, k):
    """Set of canonical k-mers (min of k-mer and reverse complement) of a small reference region (host numpy;
    flank regions are a few kbp -- reference: pavlib/inv.py:507-513, kanapy kmer.py:135-148)."""
    if len(region) < k:
        return set()
    arr = fasta.open_fasta(fa_name).fetch_array(region.chrom, region.pos, region.end)
    code = _CODE[arr]
    n = len(code) - k + 1
    if n <= 0:
        return set()
    bad = np.concatenate(([0], np.cumsum(code == 255)))
    ok = (bad[k:] - bad[:-k]) == 0
    c64 = np.where(code == 255, 0, code).astype(np.uint64)
    fwd = np.zeros(n, dtype=np.uint64)
    rc = np.zeros(n, dtype=np.uint64)
    for t in range(k):
        fwd = (fwd << np.uint64(2)) | c64[t:t + n]
        rc = rc | ((np.uint64(3) - c64[t:t + n]) << np.uint64(2 * t))
    can = np.minimum(fwd, rc)[ok]
    return set(int(x) for x in can.tolist())


def annotate_inv_dup_mers(df, region_ref_outer, region_ref_inner, region_tig_outer, region_tig_inner, region_tig_discovery,
                          ref_fa, k_util):
    """Add FLANK (UP / DN / '') and MATCH (SAME / OTHER / NaN) columns for k-mers inside flanking inverted
    duplications (reference: pavlib/inv.py:457-561; MATCH tests the raw k-mer against canonical sets, as there)."""
    flank, match = _dup_mer_columns(df['INDEX'].to_numpy(), df['KMER'].to_numpy(), region_ref_outer, region_ref_inner, region_tig_outer,
                                    region_tig_inner, region_tig_discovery, ref_fa, k_util)
    df['FLANK'] = pd.Series(flank, index=df.index, dtype=object)   # explicit dtype: no string-dtype inference pass over 50 k cells
    df['MATCH'] = pd.Series(match, index=df.index, dtype=object)
    return df


def _dup_mer_columns(index_col, kmer_col, region_ref_outer, region_ref_inner, region_tig_outer, region_tig_inner, region_tig_discovery,
                     ref_fa, k_util):
    """The FLANK and MATCH columns of ``annotate_inv_dup_mers`` as object arrays, from the INDEX and KMER columns."""
    k = int(k_util.k_size)
    dup_ref_up = pavseq.Region(region_ref_outer.chrom, region_ref_outer.pos, region_ref_inner.pos)
    dup_ref_dn = pavseq.Region(region_ref_outer.chrom, region_ref_inner.end, region_ref_outer.end)
    dup_tig_up = pavseq.Region(region_tig_outer.chrom, region_tig_outer.pos, region_tig_inner.pos)
    dup_tig_dn = pavseq.Region(region_tig_outer.chrom, region_tig_inner.end, region_tig_outer.end)
    ref_set_up = _region_canonical_kmers(dup_ref_up, ref_fa, k)
    ref_set_dn = _region_canonical_kmers(dup_ref_dn, ref_fa, k)

    qry_index = index_col.astype(np.int64) + region_tig_discovery.pos
    up = (qry_index >= dup_tig_up.pos) & (qry_index < dup_tig_up.end - k)
    dn = (qry_index >= dup_tig_dn.pos) & (qry_index < dup_tig_dn.end - k)
    flank = np.array(['', 'UP', 'DN'], dtype=object)[np.where(dn, 2, np.where(up, 1, 0))]     # DN is assigned last, as in the reference
    match = np.full(len(index_col), '', dtype=object)
    for idx, first, second in ((np.flatnonzero(up & ~dn), ref_set_up, ref_set_dn), (np.flatnonzero(dn), ref_set_dn, ref_set_up)):
        for i, km in zip(idx.tolist(), kmer_col[idx].tolist()):      # only the k-mers inside the flanking duplications
            v = KMER_LOC_STATE[int(km in first), int(km in second)]
            match[i] = np.nan if v == 'NA' else v
    return flank, match
