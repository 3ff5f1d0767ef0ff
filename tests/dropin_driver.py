#!/usr/bin/env python3
"""Child process of tests/test_dropin_gpu.py: the reference's OWN package and the UNMODIFIED `run:` blocks of its rules, with the three
hot-path functions bound to this repository exactly as INTEGRATION.md section 1 says. Test infrastructure (uses oracle/).

    python tests/dropin_driver.py call_cigar <golden dir> <out dir> <batch>
    python tests/dropin_driver.py call_inv_batch <golden dir> <out dir> <batch>
"""
import collections
import gc
import os
import re
import sys
import textwrap
import types

import numpy as np
import pandas as pd

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import refenv  # noqa: E402

refenv.activate()
import intervaltree  # noqa: E402  (stub)
import kanapy.util.kmer  # noqa: E402,F401
import pavlib  # noqa: E402  (the reference's package)
import svpoplib  # noqa: E402

# ---- the binding of INTEGRATION.md section 1: three assignments, nothing else of the reference is touched
import pav_b200.pavlib.call  # noqa: E402
import pav_b200.pavlib.cigarcall  # noqa: E402
import pav_b200.pavlib.inv  # noqa: E402

pavlib.cigarcall.make_insdel_snv_calls = pav_b200.pavlib.cigarcall.make_insdel_snv_calls
pavlib.call.left_homology = pav_b200.pavlib.call.left_homology
pavlib.call.right_homology = pav_b200.pavlib.call.right_homology
pavlib.inv.scan_for_inv = pav_b200.pavlib.inv.scan_for_inv


def rule_body(snakefile, rule):
    """Source of the ``run:`` block of ``rule`` as a function ``run(input, output, params, wildcards)`` (Snakemake wraps it the same way)."""
    lines = open(os.path.join(refenv.REF_ROOT, 'rules', snakefile)).read().split('\n')
    i = next(k for k, ln in enumerate(lines) if re.match(r'^rule\s+%s\s*:' % re.escape(rule), ln))
    j = next(k for k in range(i, len(lines)) if lines[k].rstrip() == '    run:')
    body = []
    for ln in lines[j + 1:]:
        if ln.strip() and len(ln) - len(ln.lstrip()) < 8:
            break
        body.append(ln)
    return 'def run(input, output, params, wildcards):\n' + textwrap.indent(textwrap.dedent('\n'.join(body)), '    ') + '\n'


def run_rule(snakefile, rule, input, output, wildcards, extra):
    g = {'pd': pd, 'np': np, 'collections': collections, 'intervaltree': intervaltree, 'pavlib': pavlib, 'os': os,
         'get_config': lambda wc, key=None, default=None, default_none=False: {} if key is None else default, 'BATCH_COUNT_DEFAULT': 60}
    g.update(extra)
    exec(compile(rule_body(snakefile, rule), f'{snakefile}:{rule}', 'exec'), g)

    class Wildcards(dict):
        __getattr__ = dict.__getitem__
    ns = types.SimpleNamespace
    g['run'](ns(**input), ns(**output), ns(), Wildcards(wildcards))


def main():
    what, gold, out, batch = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    os.makedirs(out, exist_ok=True)
    if what == 'call_cigar':
        run_rule('call.snakefile', 'call_cigar',
                 {'bed': os.path.join(gold, 'wl_align.bed'), 'bed_trim': os.path.join(gold, 'wl_align_trim.bed'), 'tig_fa_name': os.path.join(gold, 'wl_tig.fa')},
                 {'bed_insdel': os.path.join(out, f'insdel_{batch}.bed.gz'), 'bed_snv': os.path.join(out, f'snv_{batch}.bed.gz')},
                 {'batch': batch, 'hap': 'h1'}, {'REF_FA': os.path.join(gold, 'wl_ref.fa')})
    elif what == 'call_inv_batch':
        os.chdir(out)    # the rule writes its density tables relative to the working directory
        run_rule('call_inv.snakefile', 'call_inv_batch',
                 {'bed_flag': os.path.join(gold, 'flagged.bed.gz'), 'bed_aln': os.path.join(gold, 'align.bed'), 'tig_fa': os.path.join(gold, 'tig.fa'),
                  'fai': os.path.join(gold, 'tig.fa.fai')},
                 {'bed': os.path.join(out, f'inv_call_{batch}.bed.gz')}, {'asm_name': 'asm', 'hap': 'h1', 'batch': batch},
                 {'REF_FA': os.path.join(gold, 'ref.fa'), 'kanapy': kanapy, 'svpoplib': svpoplib, 'gc': gc, 'threads': 1,
                  'log': types.SimpleNamespace(log=os.path.join(out, f'inv_call_{batch}.log'))})
    else:
        raise SystemExit('unknown rule ' + what)


if __name__ == '__main__':
    main()
