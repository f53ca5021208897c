"""Multi-vector query != gallery scoring on the gallery kernels: the arithmetic behind the reference's submission TSV
(generate_tsv_to_reproduce2.py:63-136).

Every enroll / verify folder holds a SET of head embeddings; the reference scores a pair of folders by building all
|v1| x |v2| pairs, running similarity_f over them and taking the mean (:70-77), in a Python double loop over folders.
similarity_f is bilinear in a per-vector transform u(v), so the mean over all pairs is ONE dot product of set means:

    mean_{i,j} (u(a_i) . u(b_j) + 1) / 2  =  (mean_i u(a_i) . mean_j u(b_j) + 1) / 2

and the whole table is a [n_enroll x D] x [n_verify x D]^T product with a top-100 per row: exactly the fused cosine GEMM +
top-k of csrc/gallery.cu, run in "dot" mode (gallery rows enter un-normalised, their norm is passed as 1).

What u is depends on how the vectors are stored, because the reference's similarity_f is shape sensitive (:63-67, see
oracle/tsv_oracle.py): flat (D,) vectors give the pair cosine, u(v) = v / |v|; the (1, D) rows the reference's pipeline
actually stores (:199-201) make cosine_similarity run over a length-1 axis, i.e. compare coordinate SIGNS, and the mean
is taken over pairs and coordinates: u(v) = sign(v) / sqrt(D).  Both are supported (`layout`), default = what the stored
shape implies, as in the reference.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import abi, gallery, ops

COLUMNS = ('query', 'matched_1', 'matched_3', 'matched_10', 'answer')


def _layout_of(db_list: Sequence[Dict[Any, Any]]) -> str:
    for db in db_list:
        for entry in db.values():
            for v in entry['head_vectors']:
                return 'row' if v.dim() == 2 else 'flat'
    return 'flat'


def set_means(db: Dict[Any, Any], layout: str, device, field: str = 'head_vectors') -> Tuple[List[Any], torch.Tensor, torch.Tensor, torch.Tensor]:
    """names, types [n] (int64), counts [n] (int64), mean of u(v) per folder [n, D] fp32 on `device` (zeros for empty sets)."""
    names = list(db.keys())
    types = torch.tensor([int(db[n]['type']) for n in names], dtype=torch.int64, device=device)
    counts = torch.tensor([len(db[n][field]) for n in names], dtype=torch.int64, device=device)
    vecs = [v.reshape(-1) for n in names for v in db[n][field]]
    if not vecs:
        return names, types, counts, torch.zeros(len(names), 0, device=device)
    x = torch.stack(vecs).to(device=device, dtype=torch.float32)
    dim = x.shape[1]
    if layout == 'flat':
        u = x / x.norm(dim=1, keepdim=True).clamp_min(1e-8)             # F.cosine_similarity's eps
    else:
        u = torch.sign(x) / float(dim) ** 0.5
    owner = torch.repeat_interleave(torch.arange(len(names), device=device), counts)
    mean = torch.zeros(len(names), dim, device=device, dtype=torch.float64).index_add_(0, owner, u.double())
    mean = mean / counts.clamp_min(1).unsqueeze(1)
    return names, types, counts, mean.float()


def _dot_topk(q: torch.Tensor, g: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k gallery rows per query by the plain dot product q . g (scores fp64), on the fused cosine GEMM + top-k kernel:
    the gallery enters as it is (fp16 copy for the tensor-core pass, norm 1 for the exact re-rank), the query's own norm
    only rescales its row and is multiplied back."""
    qu, qn = gallery.prepare(q)
    g16 = g.to(torch.float16).contiguous()
    ones = torch.ones(g.shape[0], device=g.device, dtype=torch.float64)
    idx, score = gallery.cosine_topk(q, g, k, q_prepared=(qu, qn), g_prepared=(g16, ones))
    return idx, score * qn.clamp_min(1e-8).unsqueeze(1)


def calc_scores(init_db: Dict[Any, Any], extra_db: Dict[Any, Any], strategy: str = 'mean', top: int = 100,
                layout: Optional[str] = None, strict: bool = True, device='cuda') -> List[tuple]:
    """generate_tsv_to_reproduce2.py:90-120 for a whole table at once.  Rows (query name, matched_1, matched_3, matched_10,
    'name,name,...') in enroll order; folders without head vectors, or without a same-type verify folder that has some,
    produce no row.  strict: like the reference, a query with fewer than 10 scorable verify folders raises IndexError."""
    abi.require_device()
    if strategy not in ('mean', 'max'):
        raise ValueError(f'unknown strategy {strategy!r}')
    layout = layout or _layout_of([init_db, extra_db])
    device = torch.device(device)
    qn, qt, qc, qm = set_means(init_db, layout, device)
    gn, gt, gc, gm = set_means(extra_db, layout, device)
    rows: Dict[int, tuple] = {}
    for t in sorted(set(qt.tolist()) & set(gt.tolist())):
        qsel = torch.nonzero((qt == t) & (qc > 0)).flatten()
        gsel = torch.nonzero((gt == t) & (gc > 0)).flatten()
        if qsel.numel() == 0 or gsel.numel() == 0:
            continue
        if strict and gsel.numel() < 10:
            raise IndexError('list index out of range')      # the reference's np.mean([l[i][1] for i in range(10)])
        k = min(max(top, 10), int(gsel.numel()), 100)
        if strategy == 'mean':
            idx, dot = _dot_topk(qm[qsel].contiguous(), gm[gsel].contiguous(), k)
            score = ((dot + 1.0) / 2.0).clamp_min(0.0)
        else:
            idx, score = _max_strategy_topk(init_db, extra_db, qn, gn, qsel, gsel, layout, k, device)
        # Ties: set means of sign vectors take few distinct values and the reference's fp32 arithmetic keeps equal scores
        # exactly equal, ordering them by verify position (stable sort); here equal scores differ by float noise (~1e-10).
        # Consecutive scores closer than 1e-8 are chained into one tie group, groups keep their order, members go by index.
        ng = int(gsel.numel())
        new_group = torch.zeros_like(idx, dtype=torch.int64)
        new_group[:, 1:] = ((score[:, :-1] - score[:, 1:]) > 1e-8).to(torch.int64)
        key = torch.cumsum(new_group, dim=1) * (ng + 1) + idx.long().clamp_min(0)
        key = torch.where(idx >= 0, key, torch.full_like(key, torch.iinfo(torch.int64).max))
        order = torch.argsort(key, dim=1)
        idx, score = torch.gather(idx, 1, order).cpu(), torch.gather(score, 1, order).cpu()
        gsel_l = gsel.tolist()
        for r, qi in enumerate(qsel.tolist()):
            valid = idx[r] >= 0
            sc = score[r][valid].tolist()
            names = [gn[gsel_l[j]] for j in idx[r][valid].tolist()]
            m3 = sum(sc[:3]) / len(sc[:3])
            m10 = sum(sc[:10]) / len(sc[:10])
            rows[qi] = (str(getattr(qn[qi], 'name', qn[qi])), sc[0], m3, m10,
                        ','.join(str(getattr(n, 'name', n)) for n in names[:top]))
    return [rows[i] for i in sorted(rows)]


def _dense_dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a [n, D] . b [m, D]^T to ~fp32 accuracy on the fp16 tensor-core GEMM: x = hi + lo with hi = fp16(x), lo = fp16(2^11 (x - hi));
    a . b = hi.hi + 2^-11 (hi.lo + lo.hi) (the lo.lo term is below 2^-22).  Two launches of the tcgen05 GEMM with fp32 output;
    the second contracts over the concatenation [a_hi | a_lo] x [b_lo | b_hi]."""
    n, m = a.shape[0], b.shape[0]
    pad_n, pad_m = (-n) % 8, (-m) % 8

    def split(x, pad):
        if pad:
            x = torch.cat([x, torch.zeros(pad, x.shape[1], device=x.device, dtype=x.dtype)])
        hi = x.to(torch.float16)
        lo = ((x - hi.float()) * 2048.0).to(torch.float16)
        return hi.contiguous(), lo.contiguous()
    ah, al = split(a, pad_n)
    bh, bl = split(b, pad_m)
    main = ops.gemm_tn(ah, bh, out_fp32=True)
    corr = ops.gemm_tn(torch.cat([ah, al], 1).contiguous(), torch.cat([bl, bh], 1).contiguous(), out_fp32=True)
    return (main + corr * (1.0 / 2048.0))[:n, :m]


ENSEMBLE_THRESHOLDS = (0.9069641, 0.985643)      # generate_tsv_to_reproduce1.py:106, by pet type (1 = dog, 2 = cat)


def calc_scores_ensemble(init_db: Dict[Any, Any], extra_db: Dict[Any, Any], top: int = 100, layout: Optional[str] = None,
                         strict: bool = True, thresholds=ENSEMBLE_THRESHOLDS, device='cuda') -> List[tuple]:
    """Head + body ensemble of generate_tsv_to_reproduce1.py:88-120 for a whole table at once.  Per (enroll, verify) pair of one
    pet type: s0 / s1 = the mean-strategy scores of the head / body sets (0 when either set is empty); pairs with s0 + s1 == 0 are
    skipped; the score is s1 when the enroll folder has no head vector or (s0 == 0 and s1 > thresholds[type - 1]), else s0.
    The rule is not separable, so both [enroll x verify] score tables are formed densely - two set-mean products on the tcgen05
    GEMM at split precision (_dense_dot) - combined element-wise on the device and ranked with a stable sort (ties keep the
    verify order, as the reference's sorted() does)."""
    abi.require_device()
    layout = layout or _layout_of([init_db, extra_db])
    device = torch.device(device)
    qn, qt, qh, qhm = set_means(init_db, layout, device, 'head_vectors')
    _, _, qb, qbm = set_means(init_db, layout, device, 'body_vectors')
    gn, gt, gh, ghm = set_means(extra_db, layout, device, 'head_vectors')
    _, _, gb, gbm = set_means(extra_db, layout, device, 'body_vectors')
    rows: Dict[int, tuple] = {}
    for t in sorted(set(qt.tolist()) & set(gt.tolist())):
        qsel = torch.nonzero(qt == t).flatten()
        gsel = torch.nonzero(gt == t).flatten()
        if qsel.numel() == 0 or gsel.numel() == 0:
            continue

        def table(qm, qc, gm, gc):
            if qm.shape[1] == 0 or gm.shape[1] == 0:
                return torch.zeros(qsel.numel(), gsel.numel(), device=device)
            sc = ((_dense_dot(qm[qsel].contiguous(), gm[gsel].contiguous()) + 1.0) / 2.0).clamp_min(0.0)
            have = (qc[qsel] > 0).unsqueeze(1) & (gc[gsel] > 0).unsqueeze(0)
            return torch.where(have, sc, torch.zeros_like(sc))
        s0, s1 = table(qhm, qh, ghm, gh), table(qbm, qb, gbm, gb)
        thr = float(thresholds[int(t) - 1])
        use_body = (qh[qsel] == 0).unsqueeze(1) | ((s0 == 0) & (s1 > thr))
        score = torch.where(use_body, s1, s0)
        listed = (s0 + s1) != 0
        score = torch.where(listed, score, torch.full_like(score, -1.0))
        sc_sorted, order = torch.sort(score, dim=1, descending=True, stable=True)
        n_listed = listed.sum(1)
        if strict and bool(((n_listed > 0) & (n_listed < 10)).any()):
            raise IndexError('list index out of range')      # the reference's np.mean([l[i][1] for i in range(10)])
        k = min(max(top, 10), int(gsel.numel()))
        sc_sorted, order, n_listed = sc_sorted[:, :k].cpu(), order[:, :k].cpu(), n_listed.cpu()
        gsel_l = gsel.tolist()
        for r, qi in enumerate(qsel.tolist()):
            n = min(int(n_listed[r]), k)
            if n == 0:
                continue
            sc = sc_sorted[r, :n].tolist()
            names = [gn[gsel_l[j]] for j in order[r, :n].tolist()]
            rows[qi] = (str(getattr(qn[qi], 'name', qn[qi])), sc[0], sum(sc[:3]) / len(sc[:3]), sum(sc[:10]) / len(sc[:10]),
                        ','.join(str(getattr(nm, 'name', nm)) for nm in names[:top]))
    return [rows[i] for i in sorted(rows)]


def _max_strategy_topk(init_db, extra_db, qn, gn, qsel, gsel, layout, k, device):
    """max_strategy_cal_scores (:80-87): best pair of every (enroll, verify) folder pair.  Not bilinear, so the all-pairs
    score matrix of the member vectors is formed (tcgen05 GEMM on unit rows) and folded by a segmented max."""
    if layout != 'flat':
        raise NotImplementedError('max strategy on (1, D) rows compares coordinate signs (almost always 1.0); only flat vectors are supported')

    def members(db, names, sel):
        vecs, owner = [], []
        for j, i in enumerate(sel.tolist()):
            for v in db[names[i]]['head_vectors']:
                vecs.append(v.reshape(-1))
                owner.append(j)
        x = torch.stack(vecs).to(device=device, dtype=torch.float32)
        return x / x.norm(dim=1, keepdim=True).clamp_min(1e-8), torch.tensor(owner, device=device)
    qa, qo = members(init_db, qn, qsel)
    ga, go = members(extra_db, gn, gsel)
    nq, ng = int(qsel.numel()), int(gsel.numel())
    pad = (-ga.shape[0]) % 8                    # the GEMM wants its N (gallery member count) in multiples of 8: zero rows,
    if pad:                                     # owned by a dummy folder that is dropped afterwards
        ga = torch.cat([ga, torch.zeros(pad, ga.shape[1], device=device)])
        go = torch.cat([go, torch.full((pad,), ng, device=device, dtype=go.dtype)])
    cos = ops.gemm_tn(qa.to(torch.float16).contiguous(), ga.to(torch.float16).contiguous(), out_fp32=True)   # [members_q, members_g]
    best = torch.full((nq, cos.shape[1]), -2.0, device=device).scatter_reduce_(0, qo.unsqueeze(1).expand_as(cos), cos, 'amax')
    best = torch.full((nq, ng + 1), -2.0, device=device).scatter_reduce_(1, go.unsqueeze(0).expand_as(best), best, 'amax')[:, :ng]
    score, idx = torch.sort((best.double() + 1.0) / 2.0, dim=1, descending=True, stable=True)
    return idx[:, :k].to(torch.int32), score[:, :k]


def create_table(db: Dict[Any, Tuple[Dict[Any, Any], Dict[Any, Any]]], ensemble: bool = False, **kw):
    """:123-136: one calc_scores per big folder, rows concatenated into a DataFrame with the submission's columns.
    ensemble: the head + body rule of generate_tsv_to_reproduce1.py instead of the head-only scoring of script 2."""
    import pandas as pd
    rows: List[tuple] = []
    for big_folder in db:
        rows.extend((calc_scores_ensemble if ensemble else calc_scores)(*db[big_folder], **kw))
    return pd.DataFrame(data=rows, columns=COLUMNS)


def write_tsv(df, path) -> None:
    """:228: tab-separated, header, no index."""
    df.to_csv(str(path), index=False, sep='\t')


def backfill(pred_scores_path, preds_path) -> None:
    """:234-247: queries the scoring produced no row for take the row of preds.tsv; order follows preds.tsv."""
    import pandas as pd
    df1 = pd.read_csv(str(pred_scores_path), sep='\t')
    df2 = pd.read_csv(str(preds_path), sep='\t')
    d1 = {row['query']: row for _, row in df1.iterrows()}
    d2 = {row['query']: row for _, row in df2.iterrows()}
    merged = [d1[q] if q in d1 else row for q, row in d2.items()]
    pd.DataFrame(merged, columns=df1.columns).to_csv(str(pred_scores_path), index=False, sep='\t')
