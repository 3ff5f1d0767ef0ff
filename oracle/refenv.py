"""Test / baseline infrastructure: put the UNMODIFIED reference (PAV 2.4.6.0) on sys.path together with the stub third-party
modules in oracle/ref_stubs/. The reference is taken from /root/reference (build container) or, where that does not exist (GPU
box), from the byte-for-byte staged copy oracle/_ref/ (oracle/stage_ref.py; git-ignored, travels with gpurun). Used by
tests/golden/make_golden*.py to generate golden fixtures, by container-only cross-checks, and by bench.py's --impl reference and
cpu_baseline legs. Nothing under pav_b200/ imports this."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(_HERE, 'ref_stubs')
REF_ROOT = '/root/reference' if os.path.isdir('/root/reference/pavlib') else os.path.join(_HERE, '_ref')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'pavlib'))


def pythonpath_entries():
    return [STUBS, REF_ROOT, os.path.join(REF_ROOT, 'dep', 'svpop'),
            os.path.join(REF_ROOT, 'dep', 'svpop', 'dep'),
            os.path.join(REF_ROOT, 'dep', 'svpop', 'dep', 'ply')]


def activate():
    if not available():
        raise RuntimeError('reference tree not present: ' + REF_ROOT)
    for p in reversed(pythonpath_entries()):
        if p not in sys.path:
            sys.path.insert(0, p)
