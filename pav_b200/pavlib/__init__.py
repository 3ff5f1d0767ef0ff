"""Host-side mirror of the PAV ``pavlib`` interface for the hot path (same names, arguments,
return types and error behaviour as the reference), implemented on libpavgpu.so.

    pavlib.cigarcall.make_insdel_snv_calls   (reference: pavlib/cigarcall.py:24)
    pavlib.call.left_homology/right_homology (reference: pavlib/call.py:542,595)
    pavlib.align.cigar_str_to_tuples         (reference: pavlib/align/align.py:286)
    pavlib.seq.Region / region_from_string   (reference: pavlib/seq.py:20-302)
    pavlib.density.rl_encoder                (reference: pavlib/density.py:330)
    pavlib.inv.scan_for_inv                  (reference: pavlib/inv.py:149)

INTEGRATION.md shows how a PAV checkout binds these names.
"""
from . import align, call, cigarcall, constants, density, inv, lift, seq, variant  # noqa: F401
