"""CPU restatement of the multi-vector query != gallery scoring behind the submission TSV.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows generate_tsv_to_reproduce2.py of the reference:
  * similarity_f                 :63-67   (cos + 1) / 2 over stacked pairs
  * mean_strategy_cal_scores     :70-77   all |v1| x |v2| pairs, mean, clamp(min=0)
  * max_strategy_cal_scores      :80-87   the same pairs, max
  * calc_scores                  :90-120  per enroll folder: verify folders of the same type with at least one head
                                          vector, sorted by score descending (stable sort: ties keep the verify order),
                                          first 100 names, matched_1/3/10 = best / mean of best 3 / mean of best 10
  * create_table / to_csv        :123-136, :228  columns query, matched_1, matched_3, matched_10, answer; tab separated

A db maps a folder name to {'head_vectors': [tensor(1, D) or (D,), ...], 'type': int}; insertion order is the verify order.
See similarity_f for what the two vector shapes mean.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

COLUMNS = ('query', 'matched_1', 'matched_3', 'matched_10', 'answer')


def similarity_f(pairs: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """:63-67, literally: unsqueeze(0) + cat + cosine_similarity over dim 1.

    The shape of the stored vectors decides what this computes.  Flat (D,) vectors stack to (P, D) and give the cosine
    of each pair.  The reference's own pipeline stores the model output as it comes, shape (1, D) (:199-201), so the
    stack is (P, 1, D), dim 1 has length one, and cosine_similarity degenerates to x*y / max(|x*y|, eps) = the SIGN
    agreement of every coordinate: the score of a pair is a (P, D) tensor of 0/1 values and mean / max are taken over
    pairs AND coordinates.  The restatement keeps that behaviour (it is what produced the reference's TSVs)."""
    t1 = torch.cat([p[0].unsqueeze(0) for p in pairs], dim=0)
    t2 = torch.cat([p[1].unsqueeze(0) for p in pairs], dim=0)
    return (F.cosine_similarity(t1, t2) + 1) / 2


def mean_strategy(v1: List[torch.Tensor], v2: List[torch.Tensor]) -> float:
    scores = similarity_f([(i, j) for i in v1 for j in v2])
    return torch.mean(scores).clamp(min=0.0).item()


def max_strategy(v1: List[torch.Tensor], v2: List[torch.Tensor]) -> float:
    scores = similarity_f([(i, j) for i in v1 for j in v2])
    return torch.max(scores).item()


def calc_scores(init_db: Dict[str, Any], extra_db: Dict[str, Any], strategy: str = 'mean', top: int = 100) -> List[tuple]:
    """Rows (query, matched_1, matched_3, matched_10, 'name,name,...') in enroll order; queries without head vectors or
    without a scorable verify folder produce no row (the reference back-fills those from preds.tsv)."""
    score_f = mean_strategy if strategy == 'mean' else max_strategy
    rows = []
    for name, enroll in init_db.items():
        v1 = enroll['head_vectors']
        ranked = []
        for name2, verify in extra_db.items():
            if verify['type'] != enroll['type']:
                continue
            if len(v1) != 0 and len(verify['head_vectors']) != 0:
                ranked.append((name2, score_f(v1, verify['head_vectors'])))
        ranked = sorted(ranked, key=lambda x: x[1], reverse=True)
        if ranked:
            answer = [ranked[i][0] for i in range(min(top, len(ranked)))]
            rows.append((str(name), ranked[0][1], float(np.mean([ranked[i][1] for i in range(3)])),
                         float(np.mean([ranked[i][1] for i in range(10)])), ','.join(answer)))
    return rows


def synth_db(n_sets: int, dim: int, seed: int, n_ids: int, prefix: str, max_vec: int = 4, noise: float = 0.6, flat: bool = False):
    """Deterministic synthetic db: set s shows identity s % n_ids of type 1 + (identity % 2); 0..max_vec head vectors
    (a few sets are empty, as when the detector finds no head), each = identity centre + noise."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(n_ids, dim, generator=torch.Generator().manual_seed(12345))
    db = {}
    for s in range(n_sets):
        ident = s % n_ids
        nvec = int(torch.randint(0, max_vec + 1, (1,), generator=g).item())
        vecs = [(centres[ident] + noise * torch.randn(dim, generator=g)).reshape((dim,) if flat else (1, dim)) for _ in range(nvec)]
        db[f'{prefix}{s:04d}'] = {'head_vectors': vecs, 'type': 1 + ident % 2}
    return db


# ---------------------------------------------------------------------------------------------------------------------
# Head + body ensemble (generate_tsv_to_reproduce1.py of the reference)
# ---------------------------------------------------------------------------------------------------------------------
ENSEMBLE_THRESHOLDS = (0.9069641, 0.985643)          # generate_tsv_to_reproduce1.py:106, indexed by type - 1 (dog, cat)


def calc_scores_ensemble(init_db: Dict[str, Any], extra_db: Dict[str, Any], top: int = 100) -> List[tuple]:
    """generate_tsv_to_reproduce1.py:88-120.  Every folder carries 'head_vectors' AND 'body_vectors'.  Per (enroll, verify)
    pair of the same type: s0 = mean-strategy score of the head sets (0 if either is empty), s1 = the same for the body sets;
    pairs with s0 + s1 == 0 are skipped; the pair's score is s1 if the enroll folder has no head vector, or if s0 == 0 and s1
    exceeds the type's threshold - else s0 (:106-107).  Then as calc_scores: stable descending sort, first 100 names,
    matched_1 / 3 / 10."""
    rows = []
    for name, enroll in init_db.items():
        v1, v1_body = enroll['head_vectors'], enroll['body_vectors']
        ranked = []
        for name2, verify in extra_db.items():
            if verify['type'] != enroll['type']:
                continue
            s0 = mean_strategy(v1, verify['head_vectors']) if (len(v1) != 0 and len(verify['head_vectors']) != 0) else 0
            s1 = mean_strategy(v1_body, verify['body_vectors']) if (len(v1_body) != 0 and len(verify['body_vectors']) != 0) else 0
            if s0 + s1 == 0:
                continue
            score = s1 if len(v1) == 0 or (s0 == 0 and s1 > ENSEMBLE_THRESHOLDS[enroll['type'] - 1]) else s0
            ranked.append((name2, score))
        ranked = sorted(ranked, key=lambda x: x[1], reverse=True)
        if ranked:
            answer = [ranked[i][0] for i in range(min(top, len(ranked)))]
            rows.append((str(name), ranked[0][1], float(np.mean([ranked[i][1] for i in range(3)])),
                         float(np.mean([ranked[i][1] for i in range(10)])), ','.join(answer)))
    return rows


def synth_db_ensemble(n_sets: int, dim: int, seed: int, n_ids: int, prefix: str, max_vec: int = 3, noise: float = 0.15, flat: bool = True):
    """synth_db with a second, independent 'body_vectors' set per folder.  Low noise: body scores of matching identities clear
    the ensemble thresholds (0.907 / 0.986), so every branch of the rule is taken; some folders have no head, some no body."""
    g = torch.Generator().manual_seed(seed)
    head_c = torch.randn(n_ids, dim, generator=torch.Generator().manual_seed(12345))
    body_c = torch.randn(n_ids, dim, generator=torch.Generator().manual_seed(54321))
    db = {}
    for s in range(n_sets):
        ident = s % n_ids
        nh = int(torch.randint(0, max_vec + 1, (1,), generator=g).item())
        nb = int(torch.randint(0, max_vec + 1, (1,), generator=g).item())
        shape = (dim,) if flat else (1, dim)
        db[f'{prefix}{s:04d}'] = {
            'head_vectors': [(head_c[ident] + noise * torch.randn(dim, generator=g)).reshape(shape) for _ in range(nh)],
            'body_vectors': [(body_c[ident] + (0.05 if s % 3 else noise) * torch.randn(dim, generator=g)).reshape(shape) for _ in range(nb)],
            'type': 1 + ident % 2}
    return db
