"""Packed-reference sidecar on the GPU: planes identical to a fresh pack, and make_insdel_snv_calls through the sidecar gives
the same DataFrames as through the FASTA (golden tables of the reference)."""
import json
import os
import shutil

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cigar')


def _tsv(df):
    return df.to_csv(sep='\t', index=False).encode()


@pytest.mark.parametrize('case', ['c1', 'multi', 'edge_homopolymer_rev', 'kat1'])
@pytest.mark.parametrize('resident', [False, True])
def test_calls_through_sidecar(case, resident, tmp_path, monkeypatch):
    from pav_b200 import fasta, sidecar
    from pav_b200.pavlib import cigarcall
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, 'meta.json')))
    ref_fa = str(tmp_path / 'ref.fa')
    shutil.copy(os.path.join(d, 'ref.fa'), ref_fa)
    if os.path.exists(os.path.join(d, 'ref.fa.fai')):
        shutil.copy(os.path.join(d, 'ref.fa.fai'), ref_fa + '.fai')
    path = sidecar.build(ref_fa)
    assert path == ref_fa + sidecar.SUFFIX and sidecar.find(ref_fa) is not None
    if resident:
        monkeypatch.setenv('PAVGPU_REF_CACHE', '1')
    df_align = pd.read_csv(os.path.join(d, 'align.bed'), sep='\t', dtype={'#CHROM': str, 'QRY_ID': str}, keep_default_na=False)
    for _ in range(2):   # second call: resident store (or a fresh upload) again
        df_snv, df_insdel = cigarcall.make_insdel_snv_calls(df_align, ref_fa, os.path.join(d, 'tig.fa'), meta['hap'], version_id=meta['version_id'])
        assert cigarcall.last_phase_seconds['sidecar'] is True
        assert _tsv(df_snv) == open(os.path.join(d, 'snv.tsv'), 'rb').read()
        assert _tsv(df_insdel) == open(os.path.join(d, 'insdel.tsv'), 'rb').read()
        assert [int(i) for i in df_snv.index] == meta['snv_index'] and [int(i) for i in df_insdel.index] == meta['insdel_index']
    # planes in the file == planes of a fresh pack of the same sequences
    from pav_b200 import device
    fa = fasta.open_fasta(ref_fa)
    st = device.SeqStore(device.get_context(), fa.names(), [fa.fetch_array(n) for n in fa.names()], keep_host=False)
    p2, nm = st.export()
    st.close()
    sc = sidecar.Sidecar(path)
    q2, qm = sc.planes()
    assert (np.asarray(q2) == p2).all() and (np.asarray(qm) == nm).all()
    for st in sidecar._RESIDENT.values():
        st.close()
    sidecar._RESIDENT.clear()
