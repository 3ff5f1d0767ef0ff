"""pav_b200: B200-native (sm_100a) implementation of PAV's variant-calling hot path.

Path A: CIGAR walk -> SNV / INS / DEL rows (reference: pavlib/cigarcall.py:24-362).
Path B: k-mer orientation density scan for inversions (reference: scripts/density.py,
pavlib/inv.py:149-454).

The compute path is hand-written CUDA behind the C-ABI in ``include/pavgpu.h``
(``pav_b200/csrc`` -> ``pav_b200/libpavgpu.so``), loaded with ctypes; there is no CPU fallback.
"""
__version__ = '0.1.0'
