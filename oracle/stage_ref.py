#!/usr/bin/env python3
"""TEST / BASELINE INFRASTRUCTURE -- stage the UNMODIFIED reference for the machine without /root/reference.

The reference (EichlerLab/pav 2.4.6.0) is pure Python, so "building" it is copying the modules its hot path imports, byte for
byte, from /root/reference into oracle/_ref/ (git-ignored: no reference source enters the history; not gpurun-ignored: the
directory travels to the GPU box with the tree). Only the package trees the hot path imports are staged:

    pavlib/                      Path A (pavlib/cigarcall.py:24-362, pavlib/call.py:542-647, pavlib/align/align.py:286-322) and the
                                 Path B driver (pavlib/inv.py, pavlib/seq.py, pavlib/density.py)
    scripts/density.py           Path B body (scripts/density.py:423-571), run as its own process like pavlib/inv.py:249-266 does
    dep/svpop/svpoplib/          svpoplib.variant.version_id, svpoplib.ref.get_df_fai (imported by pavlib/__init__)
    dep/svpop/dep/kanapy/        kanapy.util.kmer (k-mer arithmetic / stream)
    dep/svpop/dep/ply/ply/       PLY (imported by svpoplib.svmergeconfig at import time)
    rules/call.snakefile,        the rule bodies `call_cigar` and `call_inv_batch`, executed UNMODIFIED against the overlay by
    rules/call_inv.snakefile     tests/test_dropin_gpu.py (the drop-in check at the rule level)

oracle/_ref/MANIFEST.json records the sha256 of every staged file next to the sha256 of its source, so "unmodified" can be checked.
Used by: bench.py --impl reference and bench.py's cpu_baseline leg (oracle/refenv.py resolves /root/reference first, then
oracle/_ref). Never imported by anything under pav_b200/.

    python oracle/stage_ref.py            (build container only)
"""
import hashlib
import json
import os
import shutil
import sys

SRC = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
TREES = ['pavlib', 'dep/svpop/svpoplib', 'dep/svpop/dep/kanapy', 'dep/svpop/dep/ply/ply']
FILES = ['scripts/density.py', 'LICENSE', 'dep/svpop/LICENSE', 'rules/call.snakefile', 'rules/call_inv.snakefile']


def _sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as fh:
        h.update(fh.read())
    return h.hexdigest()


def stage(force=False):
    """Copy the trees; returns the manifest. No-op (returns None) when the reference is not present (GPU box)."""
    if not os.path.isdir(os.path.join(SRC, 'pavlib')):
        return None
    man_path = os.path.join(DST, 'MANIFEST.json')
    if not force and os.path.exists(man_path):
        try:
            man = json.load(open(man_path))
            wanted = [f for f in FILES if os.path.exists(os.path.join(SRC, f))]
            if all(os.path.exists(os.path.join(DST, f)) for f in man['files']) and all(f in man['files'] for f in wanted):
                return man
        except Exception:  # noqa: BLE001
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    files = {}
    for tree in TREES:
        for root, _, names in os.walk(os.path.join(SRC, tree)):
            for n in names:
                if n.endswith('.py'):
                    rel = os.path.relpath(os.path.join(root, n), SRC)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
                    files[rel] = {'sha256': _sha(os.path.join(DST, rel)), 'source_sha256': _sha(os.path.join(SRC, rel))}
    for rel in FILES:
        if os.path.exists(os.path.join(SRC, rel)):
            os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
            shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
            files[rel] = {'sha256': _sha(os.path.join(DST, rel)), 'source_sha256': _sha(os.path.join(SRC, rel))}
    man = {'source': SRC, 'version': 'PAV 2.4.6.0', 'files': files}
    with open(man_path, 'w') as fh:
        json.dump(man, fh, indent=1, sort_keys=True)
    return man


if __name__ == '__main__':
    m = stage(force='--force' in sys.argv)
    if m is None:
        print('reference tree not present: nothing staged')
    else:
        bad = [f for f, v in m['files'].items() if v['sha256'] != v['source_sha256']]
        print(f"staged {len(m['files'])} files under {DST}; modified: {bad or 'none'}")
