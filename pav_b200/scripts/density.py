"""Command-line twin of the reference's ``scripts/density.py`` (:423-571): same arguments, same outputs
(base64 pickle of the DataFrame on stdout, or .tsv / .tsv.gz / .xlsx files), same exit codes (0, or 125 for
the two soft failures :510-527), with the table computed by ``pavgpu_density_batch_*`` on the GPU.

``pavlib.inv.scan_for_inv`` no longer needs a process per expansion (it calls the library directly); this
entry point keeps the inner process boundary for callers that still spawn it:

    python3 pav_b200/scripts/density.py --tigregion R --refregion R --ref FA --tig FA -k 31 -t 1 -r false --staterunsmooth 20
"""
import argparse
import codecs
import os
import pickle
import sys

import numpy as np

sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.realpath(__file__)))))

from pav_b200 import fasta  # noqa: E402
from pav_b200.pavlib import constants, density, seq  # noqa: E402

MAX_REF_KMER_COUNT = 100   # scripts/density.py:45

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate('ACGT'):
    _CODE[ord(_c)] = _CODE[ord(_c.lower())] = _i


def get_bool(bool_str):
    """'true' / 't' / '1' and 'false' / 'f' / '0', any case; anything else raises (scripts/density.py:394-416)."""
    bs_lower = bool_str.lower()
    if bs_lower in {'true', 't', '1'}:
        return True
    if bs_lower in {'false', 'f', '0'}:
        return False
    raise RuntimeError(f'Unrecognized boolean string value: {bool_str}')


def write_table(df, out_file_name, test=False):
    """.tsv, .tsv.gz or .xlsx by extension; ``test`` only checks the name (scripts/density.py:367-391)."""
    out_lower = out_file_name.lower()
    if out_lower.endswith('.tsv'):
        if not test:
            df.to_csv(out_file_name, sep='\t', index=False)
    elif out_lower.endswith('.tsv.gz'):
        if not test:
            df.to_csv(out_file_name, sep='\t', index=False, compression='gzip')
    elif out_lower.endswith('.xlsx'):
        if not test:
            df.to_excel(out_file_name, index=False)
    else:
        raise RuntimeError(f'No recognized extension on output file: {out_file_name}: Expected tsv, tsv.gz, or xlsx')


def ref_kmer_failure(ref_arr, k):
    """Why the reference window was refused, for the message only (the decision itself came from the device): ``None`` if the
    window has no k-mer at all, else ``(max count, its k-mer as text)`` with ties resolved to the k-mer seen first, the order
    of the reference's Counter (scripts/density.py:505-519)."""
    code = _CODE[np.asarray(ref_arr, dtype=np.uint8)]
    n = len(code) - k + 1
    if n <= 0:
        return None
    bad = np.concatenate(([0], np.cumsum(code == 255)))
    ok = (bad[k:] - bad[:-k]) == 0
    if not ok.any():
        return None
    val = np.zeros(n, dtype=np.uint64)
    c64 = (code & 3).astype(np.uint64)
    for j in range(k):
        val = (val << np.uint64(2)) | c64[j:j + n]
    start = np.flatnonzero(ok)
    uniq, first, count = np.unique(val[start], return_index=True, return_counts=True)
    top = count.max()
    at = start[first[count == top].min()]
    return int(top), ''.join('ACGT'[c] for c in code[at:at + k])


def main(argv=None):
    parser = argparse.ArgumentParser('Inversion density calculation')
    parser.add_argument('--tigregion', help='Contig region to extract.')
    parser.add_argument('--refregion', help='Reference region to extract.')
    parser.add_argument('--ref', help='Reference FASTA file')
    parser.add_argument('--tig', help='Contig FASTA file')
    parser.add_argument('-k', type=int, default=31, help='K-mer size')
    parser.add_argument('-t', '--threads', default=1, type=int, help='Accepted for compatibility; the GPU does the work.')
    parser.add_argument('-r', '--revcompl', help='Reverse-complement reference k-mers: true/t/1 or false/f/0.')
    parser.add_argument('--mininf', type=int, default=2000, help='Minimum informative k-mers for a smoothed table.')
    parser.add_argument('--densmooth', type=int, default=1, help='Factor on the Scott bandwidth.')
    parser.add_argument('--minstatecount', type=int, default=20, help='States with fewer k-mers are dropped.')
    parser.add_argument('--staterunsmooth', type=int, default=20, help='Density is sampled once per this many k-mers, then filled.')
    parser.add_argument('--staterundelta', type=float, default=0.005, help='Density change that forces exact values between samples.')
    parser.add_argument('outfile', nargs='*', help='.tsv, .tsv.gz or .xlsx; none = base64 pickle on stdout.')
    args = parser.parse_args(argv)

    do_stdout = len(args.outfile) == 0
    is_rev = get_bool(args.revcompl)
    if not do_stdout:
        for out_file_name in args.outfile:
            write_table(None, out_file_name, test=True)

    region_ref = seq.region_from_string(args.refregion)
    region_tig = seq.region_from_string(args.tigregion)
    # reference window always forward (pavlib/seq.py:316); contig window reverse-complemented only if its coordinates
    # arrived reversed (pavlib/seq.py:353-355)
    ref = fasta.open_fasta(args.ref).fetch_array(region_ref.chrom, region_ref.pos, region_ref.end)
    tig = fasta.open_fasta(args.tig).fetch_array(region_tig.chrom, region_tig.pos, region_tig.end)
    if region_tig.is_rev:
        tig = fasta.reverse_complement(tig)
    res = density.density_windows([(ref, tig, is_rev, args.staterunsmooth)], k=args.k, min_informative=args.mininf,
                                  min_state_count=args.minstatecount, smooth=float(args.densmooth), delta=args.staterundelta,
                                  max_ref_kmer_count=MAX_REF_KMER_COUNT)[0]
    if res['status'] != 0:
        why = ref_kmer_failure(ref, args.k)
        if why is None:
            print(f'No reference k-mers for region {region_ref}', file=sys.stdout)   # stdout, like the reference (:511)
        else:
            print('K-mer count exceeds max: {} > {} ({}): {}'.format(why[0], MAX_REF_KMER_COUNT, why[1], region_ref), file=sys.stderr)
        return constants.ERR_INV_FAIL

    df = density.frame_from_result(res)
    if do_stdout:
        sys.stdout.write(codecs.encode(pickle.dumps(df), 'base64').decode())
    else:
        for out_file_name in args.outfile:
            write_table(df, out_file_name)
    return 0


if __name__ == '__main__':
    sys.exit(main())
