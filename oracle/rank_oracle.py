"""CPU restatement of the gallery-matching / Recall@K path.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def similarity_f(pairs: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """configs/dog_fe/fe_dogs_config.py:89-93: stack both sides, (cosine_similarity + 1) / 2."""
    a = torch.stack([p[0] for p in pairs], dim=0)
    b = torch.stack([p[1] for p in pairs], dim=0)
    return (F.cosine_similarity(a, b) + 1) / 2


def restore_dataset_order(outputs: List[Dict[str, torch.Tensor]]):
    """engine/controller.py:51-56: concatenate the per-batch dicts and undo the loader order."""
    emb = torch.cat([o['emb'] for o in outputs], dim=0)
    classes = torch.cat([o['label'] for o in outputs], dim=0)
    index = torch.cat([o['index'] for o in outputs], dim=0)
    order = torch.argsort(index)
    return emb[order], classes[order]


def recall_at_k_loop(emb: torch.Tensor, classes: torch.Tensor, ks: Iterable[int] = (10, 100)) -> Dict[str, float]:
    """The O(N^2) leave-one-out loop of engine/controller.py:77-91 (same as :143-160), statement
    for statement: every other embedding is scored against query j with similarity_f, a full
    descending argsort ranks them, hit@k = any same-class among the first k, and the denominator
    counts the queries that have at least one same-class partner."""
    ks = list(ks)
    n = classes.shape[0]
    tally = {k: [0, 0] for k in ks}
    for j in range(n):
        keep = [i for i in range(n) if i != j]
        others = emb[keep, :]
        scores = similarity_f([(emb[j], others[i]) for i in range(len(keep))])
        ranked_classes = classes[torch.as_tensor(keep)][torch.argsort(scores, descending=True)]
        for k in ks:
            tally[k][0] += int((classes[j] == ranked_classes[:k]).sum().item() != 0)
            tally[k][1] += int((classes[j] == ranked_classes).sum().item() != 0)
    return {f'Recall@K={k}': (hit / valid if valid else float('nan')) for k, (hit, valid) in tally.items()}


# ---------------------------------------------------------------------------------------------
# Deterministic specification of the top-k the CUDA path must reproduce bit-exactly.
#
# The reference's torch.argsort(descending=True) is not stable and its fp32 cosine depends on the
# summation order, so ties / near-ties are under-determined there.  The new path
# DEFINES: score(q, g) = <q, g> / (max(|q|, 1e-8) * max(|g|, 1e-8)) evaluated in fp64 from the fp32
# embeddings (F.cosine_similarity's formula, configs/dog_fe/fe_dogs_config.py:93); candidates are
# ranked by (score descending, gallery index ascending).  The (s + 1) / 2 map of similarity_f is
# monotone and does not change the ranking.
# ---------------------------------------------------------------------------------------------

def cosine_f64(q: np.ndarray, g: np.ndarray) -> np.ndarray:
    q64, g64 = q.astype(np.float64), g.astype(np.float64)
    nq = np.maximum(np.sqrt((q64 * q64).sum(axis=1)), 1e-8)
    ng = np.maximum(np.sqrt((g64 * g64).sum(axis=1)), 1e-8)
    out = np.empty((q.shape[0], g.shape[0]), dtype=np.float64)
    for i in range(q.shape[0]):           # row-at-a-time elementwise sum: identical gallery rows
        out[i] = (g64 * q64[i][None, :]).sum(axis=1) / (nq[i] * ng)   # give identical scores
    return out


def topk_spec(q: np.ndarray, g: np.ndarray, k: int, exclude_self_offset: int = -1):
    """Exact top-k under the defined order.  q/g are the raw fp32 embeddings.  If
    exclude_self_offset >= 0, gallery row (offset + i) is skipped for query i (leave-one-out form,
    engine/controller.py:80).  Returns (idx int32 [Q,k], score fp64 [Q,k]); missing slots
    (fewer than k candidates) are -1 / -inf."""
    sc = cosine_f64(q, g)
    if exclude_self_offset >= 0:
        for i in range(q.shape[0]):
            j = exclude_self_offset + i
            if 0 <= j < g.shape[0]:
                sc[i, j] = -np.inf
    idx = np.full((q.shape[0], k), -1, dtype=np.int32)
    val = np.full((q.shape[0], k), -np.inf, dtype=np.float64)
    cols = np.arange(g.shape[0])
    for i in range(q.shape[0]):
        order = np.lexsort((cols, -sc[i]))[:k]
        order = order[sc[i, order] > -np.inf]
        idx[i, :len(order)], val[i, :len(order)] = order, sc[i, order]
    return idx, val


def recall_from_topk(top_idx: np.ndarray, q_classes: np.ndarray, g_classes: np.ndarray,
                     ks: Iterable[int], exclude_self_offset: int = -1) -> Dict[str, float]:
    """hit@k / valid exactly as engine/controller.py:86-90 counts them, from ranked indices."""
    out = {}
    nq = top_idx.shape[0]
    counts = {}
    for c in g_classes.tolist():
        counts[c] = counts.get(c, 0) + 1
    valid = 0
    for i in range(nq):
        same = counts.get(int(q_classes[i]), 0)
        if exclude_self_offset >= 0 and 0 <= exclude_self_offset + i < len(g_classes) \
                and g_classes[exclude_self_offset + i] == q_classes[i]:
            same -= 1
        valid += int(same > 0)
    for k in ks:
        hit = 0
        for i in range(nq):
            ids = top_idx[i, :k]
            ids = ids[ids >= 0]
            hit += int((g_classes[ids] == q_classes[i]).any())
        out[f'Recall@K={k}'] = hit / valid if valid else float('nan')
    return out


def gallery_match_vectorised(q: torch.Tensor, g: torch.Tensor, k: int, chunk: int = 4096):
    """Vectorised fp32 CPU form used as the timed CPU baseline at scale:
    F.normalize -> chunked matmul -> topk.  Same ranking as the loop up to fp ties."""
    qn, gn = F.normalize(q), F.normalize(g)
    idx = torch.empty(q.shape[0], k, dtype=torch.int64)
    val = torch.empty(q.shape[0], k)
    for s in range(0, q.shape[0], chunk):
        sc = qn[s:s + chunk] @ gn.t()
        v, i = sc.topk(k, dim=1)
        idx[s:s + chunk], val[s:s + chunk] = i, v
    return idx, val
