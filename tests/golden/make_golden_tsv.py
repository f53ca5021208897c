"""Generate tests/golden/tsv_scores.json by RUNNING THE REFERENCE's own scoring functions (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_tsv.py

generate_tsv_to_reproduce2.py imports cv2, the detector stack and the datasets at module level, none of which the
scoring needs; so only the function definitions similarity_f / mean_strategy_cal_scores / max_strategy_cal_scores /
calc_scores are taken from its source (ast, unmodified) and executed with torch / numpy / typing in scope.  Inputs come
from oracle.tsv_oracle.synth_db (seeded), so the fixture holds only the reference's outputs.
"""
import ast
import json
import sys
from pathlib import Path
from typing import Any, Dict, List

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.tsv_oracle import synth_db      # noqa: E402

REF = Path('/root/reference/generate_tsv_to_reproduce2.py')
WANT = ('similarity_f', 'mean_strategy_cal_scores', 'max_strategy_cal_scores', 'calc_scores')


def reference_functions():
    tree = ast.parse(REF.read_text())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    assert len(body) == len(WANT)
    ns = {'torch': torch, 'F': F, 'np': np, 'List': List, 'Dict': Dict, 'Any': Any, 'Path': Path, 'tqdm': lambda it, **kw: it}
    exec(compile(ast.Module(body=body, type_ignores=[]), str(REF), 'exec'), ns)
    return ns


REF1 = Path('/root/reference/generate_tsv_to_reproduce1.py')


def reference_functions_ensemble():
    """The head + body ensemble: the same four function names, taken from generate_tsv_to_reproduce1.py (:63-120)."""
    tree = ast.parse(REF1.read_text())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    assert len(body) == len(WANT)
    ns = {'torch': torch, 'F': F, 'np': np, 'List': List, 'Dict': Dict, 'Any': Any, 'Path': Path, 'tqdm': lambda it, **kw: it}
    exec(compile(ast.Module(body=body, type_ignores=[]), str(REF1), 'exec'), ns)
    return ns


def main_ensemble():
    from oracle.tsv_oracle import synth_db_ensemble
    ns = reference_functions_ensemble()
    cases = {}
    for case, (n_q, n_g, n_ids, seed) in {'small': (14, 70, 16, 5), 'medium': (40, 300, 50, 6)}.items():
        init = {Path(k): v for k, v in synth_db_ensemble(n_q, 512, seed, n_ids, 'q').items()}
        extra = {Path(k): v for k, v in synth_db_ensemble(n_g, 512, seed + 100, n_ids, 'g').items()}
        rows = ns['calc_scores'](init, extra)
        cases[case] = {'args': [n_q, n_g, n_ids, seed], 'rows': [[r[0], float(r[1]), float(r[2]), float(r[3]), r[4]] for r in rows]}
        print('ensemble', case, len(rows), 'rows')
    (Path(__file__).resolve().parent / 'tsv_scores_ensemble.json').write_text(json.dumps(cases, indent=0))


def main():
    ns = reference_functions()
    cases = {}
    # 'row': vectors shaped (1, D) as the reference's pipeline stores them (sign-agreement scores, see oracle.tsv_oracle);
    # 'flat': (D,) vectors (pair cosine)
    for case, (n_q, n_g, n_ids, seed, flat) in {'small_row': (12, 60, 15, 1, False), 'medium_row': (40, 260, 60, 2, False),
                                                'small_flat': (12, 60, 15, 3, True), 'medium_flat': (40, 260, 60, 4, True)}.items():
        init = {Path(k): v for k, v in synth_db(n_q, 512, seed, n_ids, 'q', flat=flat).items()}
        extra = {Path(k): v for k, v in synth_db(n_g, 512, seed + 100, n_ids, 'g', flat=flat).items()}
        rows = ns['calc_scores'](init, extra)
        cases[case] = {'args': [n_q, n_g, n_ids, seed, flat], 'rows': [[r[0], float(r[1]), float(r[2]), float(r[3]), r[4]] for r in rows]}
        v1 = next(v['head_vectors'] for v in init.values() if len(v['head_vectors']) > 1)
        v2 = next(v['head_vectors'] for v in extra.values() if len(v['head_vectors']) > 1)
        cases[case]['max_strategy_first_pair'] = ns['max_strategy_cal_scores'](v1, v2)
        print(case, len(rows), 'rows')
    (Path(__file__).resolve().parent / 'tsv_scores.json').write_text(json.dumps(cases, indent=0))


if __name__ == '__main__':
    main()
    main_ensemble()
