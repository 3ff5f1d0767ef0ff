"""Host-side mirror of the PAV ``pavlib`` interface for the hot path (same names, arguments,
return types and error behaviour as the reference), implemented on libpavgpu.so.

    pavlib.cigarcall.make_insdel_snv_calls   (reference: pavlib/cigarcall.py:24)
    pavlib.call.left_homology/right_homology (reference: pavlib/call.py:542,595)
    pavlib.align.cigar_str_to_tuples         (reference: pavlib/align/align.py:286)
    pavlib.seq.Region / region_from_string   (reference: pavlib/seq.py:20-302)
    pavlib.density.rl_encoder                (reference: pavlib/density.py:330)
    pavlib.inv.scan_for_inv                  (reference: pavlib/inv.py:149)

and, as functions, the bodies of the Snakemake rules either side of the two paths (``pavlib.flag``; the reference has them
inline in rules/call.snakefile and rules/call_inv.snakefile): call_cigar (+ direct gz-TSV writer), call_cigar_merge,
call_inv_cluster, call_inv_flag_insdel_cluster, call_inv_merge_flagged_loci, call_inv_batch.

INTEGRATION.md shows how a PAV checkout binds these names.
"""
from . import align, call, cigarcall, constants, density, flag, inv, lift, seq, variant  # noqa: F401
