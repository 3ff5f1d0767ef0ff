"""The drop-in boundary at the RULE level: the reference's own `pavlib` package and the unmodified `run:` blocks of `rule call_cigar`
(rules/call.snakefile:800-846) and `rule call_inv_batch` (rules/call_inv.snakefile:127-311), with the three hot-path functions bound
to this repository as INTEGRATION.md section 1 prescribes, must write the tables (and the log) the reference wrote with its own
functions (tests/golden/flag/filter, tests/golden/flag/inv_batch). Runs in a child process (tests/dropin_driver.py) because it puts
the reference's packages on sys.path. Needs the reference tree: /root/reference or the staged copy oracle/_ref (oracle/stage_ref.py)."""
import gzip
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, 'tests', 'golden', 'flag')


def _reference_rules():
    sys.path.insert(0, REPO)
    from oracle import refenv
    return refenv.available() and os.path.exists(os.path.join(refenv.REF_ROOT, 'rules', 'call_inv.snakefile'))


def _run(what, gold, out, batch):
    p = subprocess.run([sys.executable, os.path.join(REPO, 'tests', 'dropin_driver.py'), what, gold, str(out), str(batch)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-3000:]


@pytest.mark.parametrize('batch', [0, 1])
def test_unmodified_call_cigar_rule_body(batch, tmp_path):
    if not _reference_rules():
        pytest.skip('reference rules not staged (oracle/stage_ref.py)')
    gold = os.path.join(GOLDEN, 'filter')
    _run('call_cigar', gold, tmp_path, batch)
    for name in (f'snv_{batch}.bed.gz', f'insdel_{batch}.bed.gz'):
        assert gzip.open(tmp_path / name, 'rt').read() == gzip.open(os.path.join(gold, name), 'rt').read(), name


@pytest.mark.parametrize('batch', [0, 1])
def test_unmodified_call_inv_batch_rule_body(batch, tmp_path):
    if not _reference_rules():
        pytest.skip('reference rules not staged (oracle/stage_ref.py)')
    gold = os.path.join(GOLDEN, 'inv_batch')
    _run('call_inv_batch', gold, tmp_path, batch)
    name = f'inv_call_{batch}.bed.gz'
    assert gzip.open(tmp_path / name, 'rt').read() == gzip.open(os.path.join(gold, name), 'rt').read()
    log = os.path.join(gold, f'inv_call_{batch}.log')
    if os.path.exists(log):
        assert open(tmp_path / f'inv_call_{batch}.log').read() == open(log).read()
