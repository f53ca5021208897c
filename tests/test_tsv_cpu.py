"""Row 8f-1 (multi-vector scoring + TSV wire format), CPU side: the oracle restatement against the outputs of the
reference's own functions (tests/golden/tsv_scores.json, written by tests/golden/make_golden_tsv.py) and the TSV
write / back-fill host logic (generate_tsv_to_reproduce2.py:228-247)."""
import json

import pytest

from oracle import tsv_oracle as T


@pytest.fixture(scope='module')
def golden(golden_dir):
    return json.loads((golden_dir / 'tsv_scores.json').read_text())


@pytest.mark.parametrize('case', ['small_row', 'medium_row', 'small_flat', 'medium_flat'])
def test_oracle_matches_reference_outputs(golden, case):
    c = golden[case]
    n_q, n_g, n_ids, seed, flat = c['args']
    rows = T.calc_scores(T.synth_db(n_q, 512, seed, n_ids, 'q', flat=flat), T.synth_db(n_g, 512, seed + 100, n_ids, 'g', flat=flat))
    assert len(rows) == len(c['rows']) > 0
    for a, b in zip(rows, c['rows']):
        assert a[0] == b[0] and a[4] == b[4]
        assert a[1] == pytest.approx(b[1], abs=1e-7) and a[2] == pytest.approx(b[2], abs=1e-7) and a[3] == pytest.approx(b[3], abs=1e-7)


def test_sign_agreement_identity():
    """What the (1, D) layout computes, stated directly: the mean over pairs and coordinates of [sign(a) == sign(b)]."""
    import torch
    db_q, db_g = T.synth_db(6, 64, 9, 5, 'q'), T.synth_db(30, 64, 10, 5, 'g')
    name, enroll = next((k, v) for k, v in db_q.items() if len(v['head_vectors']) > 1)
    verify = next(v for v in db_g.values() if len(v['head_vectors']) > 1)
    a = torch.cat(enroll['head_vectors']).sign()
    b = torch.cat(verify['head_vectors']).sign()
    direct = ((a.unsqueeze(1) * b.unsqueeze(0) + 1) / 2).mean().item()
    assert T.mean_strategy(enroll['head_vectors'], verify['head_vectors']) == pytest.approx(direct, abs=1e-6)


def test_tsv_write_and_backfill(tmp_path):
    import sys
    import pandas as pd
    from b200 import multivector as mv                      # host-side helpers only: no CUDA call here
    rows = [('q1', 0.9, 0.8, 0.7, 'g3,g1'), ('q3', 0.6, 0.5, 0.4, 'g2')]
    out = tmp_path / 'pred_scores_test2.tsv'
    mv.write_tsv(pd.DataFrame(rows, columns=mv.COLUMNS), out)
    assert out.read_text().splitlines()[0] == 'query\tmatched_1\tmatched_3\tmatched_10\tanswer'
    preds = tmp_path / 'preds.tsv'
    pd.DataFrame([('q1', 0.1, 0.1, 0.1, 'x'), ('q2', 0.2, 0.2, 0.2, 'y'), ('q3', 0.3, 0.3, 0.3, 'z')], columns=mv.COLUMNS).to_csv(preds, index=False, sep='\t')
    mv.backfill(out, preds)
    df = pd.read_csv(out, sep='\t')
    assert df['query'].tolist() == ['q1', 'q2', 'q3']        # preds.tsv order; q2 back-filled
    assert df['answer'].tolist() == ['g3,g1', 'y', 'g2'] and df['matched_1'].tolist() == [0.9, 0.2, 0.6]


def test_ensemble_oracle_equals_reference_outputs(golden_dir):
    """oracle.tsv_oracle.calc_scores_ensemble == the reference's own calc_scores of generate_tsv_to_reproduce1.py (:88-120),
    run by tests/golden/make_golden_tsv.py on the same seeded folder databases."""
    import json
    from oracle import tsv_oracle as T
    cases = json.loads((golden_dir / 'tsv_scores_ensemble.json').read_text())
    for case, c in cases.items():
        n_q, n_g, n_ids, seed = c['args']
        rows = T.calc_scores_ensemble(T.synth_db_ensemble(n_q, 512, seed, n_ids, 'q'), T.synth_db_ensemble(n_g, 512, seed + 100, n_ids, 'g'))
        assert len(rows) == len(c['rows'])
        for a, b in zip(rows, c['rows']):
            assert a[0] == b[0] and a[4] == b[4]
            assert abs(a[1] - b[1]) < 1e-7 and abs(a[2] - b[2]) < 1e-7 and abs(a[3] - b[3]) < 1e-7
