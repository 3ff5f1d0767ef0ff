"""Test infrastructure: put the UNMODIFIED reference (/root/reference, PAV 2.4.6.0) on sys.path
together with the stub third-party modules in oracle/ref_stubs/. Only usable in the build
container (the reference tree does not exist on the GPU box). Used by tests/golden/make_golden.py
to generate golden fixtures and by container-only cross-checks."""
import os
import sys

REF_ROOT = '/root/reference'
_HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(_HERE, 'ref_stubs')


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
