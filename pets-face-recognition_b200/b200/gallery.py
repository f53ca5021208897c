"""Gallery matching on the B200 kernels (csrc/gallery.cu): cosine top-k and Recall@K (candR@k).

Replaces the O(N^2) Python loop of the reference's Controller.test_epoch_end / _evaluate
(engine/controller.py:77-91, :143-160) and the all-pairs scoring of generate_tsv_to_reproduce2.py:63-119.
Ranking is by cosine (similarity_f's (cos + 1) / 2 is monotone), under the deterministic order
(score desc, gallery index asc) - see oracle/rank_oracle.py:topk_spec.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

from . import abi
from .abi import check, lib, ptr, stream_ptr


def prepare(emb: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 [n, dim] -> (unit fp16 rows [n, dim], fp64 norms [n])."""
    emb = emb.contiguous().float()
    n, dim = emb.shape
    unit = torch.empty(n, dim, device=emb.device, dtype=torch.float16)
    norm = torch.empty(n, device=emb.device, dtype=torch.float64)
    check(lib().b200_gallery_prepare(ptr(emb), ptr(unit), ptr(norm), n, dim, stream_ptr()), 'gallery_prepare')
    return unit, norm


_G_SCALE = 64.0      # power of two: lifts the centred gallery rows well into the fp16 normal range; ranking is scale-invariant


class Prepared:
    """fp16 rows of one embedding set for the tensor-core pass + what the exactness certificate needs (csrc/gallery.cu).
    A gallery defines the frame (its mean unit row mu and the reflection that puts mu on the first axis) and is stored as
    scale * H (g^ - mu); queries matched against it are stored as H q^ in the SAME frame (`frame_of=` the gallery's Prepared)."""

    def __init__(self, emb: torch.Tensor, as_gallery: bool, frame_of: 'Prepared' = None):
        emb = emb.contiguous().float()
        n, dim = emb.shape
        dev = emb.device
        L = lib()
        self.rows = torch.empty(n, dim, device=dev, dtype=torch.float16)
        self.norm = torch.empty(n, device=dev, dtype=torch.float64)
        self.err = torch.empty(n, 4, device=dev, dtype=torch.float32)
        self.stats = torch.zeros(4, device=dev, dtype=torch.float32)
        self.frame, self.scale = None, 1.0
        if as_gallery and n > 0:
            mean = torch.empty(dim, device=dev, dtype=torch.float32)
            partial = torch.empty(L.b200_unit_row_mean_blocks(n), dim, device=dev, dtype=torch.float32)
            check(L.b200_unit_row_mean(ptr(emb), n, dim, ptr(mean), ptr(partial), stream_ptr()), 'unit_row_mean')
            self.frame = torch.empty(2 * dim + 1, device=dev, dtype=torch.float32)
            check(L.b200_gallery_frame(ptr(mean), dim, ptr(self.frame), stream_ptr()), 'gallery_frame')
            self.scale = _G_SCALE
        elif frame_of is not None:
            self.frame = frame_of.frame
        check(L.b200_gallery_prepare_ex(ptr(emb), ptr(self.frame), int(as_gallery), self.scale, ptr(self.rows), ptr(self.norm),
                                        ptr(self.err), ptr(self.stats), n, dim, stream_ptr()), 'gallery_prepare_ex')


def cosine_topk(q: torch.Tensor, g: torch.Tensor, k: int, exclude_self_offset: Optional[int] = None, g_index_base: int = 0,
                q_prepared=None, g_prepared=None, return_uncertified: bool = False):
    """Top-k gallery rows per query: (idx int32 [nq, k] (+ g_index_base, -1 = none), score fp64 [nq, k]).
    exclude_self_offset = o skips gallery row (o + i) for query i (leave-one-out, engine/controller.py:80).

    Default path: centred fp16 gallery rows + the exactness certificate (queries it cannot prove are re-done by an exact
    scan on the device; return_uncertified=True also returns how many that were).  Passing legacy (rows, norm) tuples as
    q_prepared / g_prepared runs the uncertified kernel pair (the dot-mode scoring of b200/multivector.py)."""
    abi.require_device()
    q = q.contiguous().float()
    g = g.contiguous().float()
    nq, dim = q.shape
    ng = g.shape[0]
    idx = torch.empty(nq, k, device=q.device, dtype=torch.int32)
    score = torch.empty(nq, k, device=q.device, dtype=torch.float64)
    wsb = lib().b200_cosine_topk_workspace_bytes(nq, ng, dim, k)
    ws = torch.empty(wsb, device=q.device, dtype=torch.uint8)
    off = abi.NO_EXCLUDE if exclude_self_offset is None else int(exclude_self_offset)
    legacy = isinstance(q_prepared, tuple) or isinstance(g_prepared, tuple)
    if legacy:
        qu, qn = q_prepared if q_prepared is not None else prepare(q)
        gu, gn = g_prepared if g_prepared is not None else prepare(g)
        check(lib().b200_cosine_topk(ptr(q), ptr(qu), ptr(qn), nq, ptr(g), ptr(gu), ptr(gn), ng, dim, k, off, g_index_base,
                                     ptr(idx), ptr(score), ptr(ws), wsb, stream_ptr()), 'cosine_topk')
        return (idx, score, 0) if return_uncertified else (idx, score)
    gp = g_prepared if g_prepared is not None else Prepared(g, as_gallery=True)
    qp = q_prepared if q_prepared is not None else Prepared(q, as_gallery=False, frame_of=gp)
    unc = torch.empty(1 + nq, device=q.device, dtype=torch.int32)
    check(lib().b200_cosine_topk_certified(ptr(q), ptr(qp.rows), ptr(qp.norm), ptr(qp.err), nq, ptr(g), ptr(gp.rows), ptr(gp.norm),
                                           ptr(gp.frame), gp.scale, ptr(gp.stats), ng, dim, k, off, g_index_base, ptr(idx), ptr(score),
                                           ptr(unc), ptr(ws), wsb, stream_ptr()), 'cosine_topk_certified')
    return (idx, score, unc[0]) if return_uncertified else (idx, score)


def topk_merge(scores: torch.Tensor, idx: torch.Tensor, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge `lists` already-scored top lists per query.  scores fp64 / idx int32: [lists, nq, k_in]."""
    lists, nq, k_in = scores.shape
    out_idx = torch.empty(nq, k_out, device=scores.device, dtype=torch.int32)
    out_score = torch.empty(nq, k_out, device=scores.device, dtype=torch.float64)
    check(lib().b200_topk_merge(ptr(scores.contiguous()), ptr(idx.contiguous()), nq, lists, k_in, k_out, ptr(out_idx), ptr(out_score),
                                stream_ptr()), 'topk_merge')
    return out_idx, out_score


def pair_similarity(emb: torch.Tensor, idx1: torch.Tensor, idx2: torch.Tensor) -> torch.Tensor:
    """(cos(emb[idx1[p]], emb[idx2[p]]) + 1) / 2 for every pair, fp32 [n_pairs] on the device: the pair-verification scores of
    engine/controller.py:60-68 without building 2 x n_pairs Python tensors."""
    abi.require_device()
    emb = emb.contiguous().float()
    idx1 = idx1.to(device=emb.device, dtype=torch.int64).contiguous()
    idx2 = idx2.to(device=emb.device, dtype=torch.int64).contiguous()
    if idx1.numel() and (int(torch.max(torch.maximum(idx1, idx2))) >= emb.shape[0] or int(torch.min(torch.minimum(idx1, idx2))) < 0):
        raise abi.B200Error('pair_similarity: pair index out of range')
    out = torch.empty(idx1.numel(), device=emb.device, dtype=torch.float32)
    check(lib().b200_pair_similarity(ptr(emb), emb.shape[0], emb.shape[1], ptr(idx1), ptr(idx2), idx1.numel(), ptr(out), stream_ptr()),
          'pair_similarity')
    return out


def recall_hits(top_idx: torch.Tensor, q_class: torch.Tensor, g_class: torch.Tensor, ks: Iterable[int]) -> torch.Tensor:
    """hits[i] = #queries with a same-class gallery row among their first ks[i] candidates (engine/controller.py:86-87)."""
    ks = list(ks)
    kst = torch.tensor(ks, dtype=torch.int32, device=top_idx.device)
    hits = torch.zeros(len(ks), dtype=torch.int64, device=top_idx.device)
    check(lib().b200_recall_hits(ptr(top_idx.contiguous()), top_idx.shape[0], top_idx.shape[1], ptr(q_class.contiguous().long()),
                                 ptr(g_class.contiguous().long()), ptr(kst), len(ks), ptr(hits), stream_ptr()), 'recall_hits')
    return hits


def valid_queries(q_class: torch.Tensor, g_class: torch.Tensor, leave_one_out: bool) -> torch.Tensor:
    """Denominator of engine/controller.py:88: queries with at least one same-class gallery row (other than themselves)."""
    uniq, counts = torch.unique(g_class, return_counts=True)
    pos = torch.searchsorted(uniq, q_class).clamp_(max=uniq.numel() - 1)
    same = torch.where(uniq[pos] == q_class, counts[pos], torch.zeros_like(counts[pos]))
    if leave_one_out:
        same = same - 1
    return (same > 0).sum()


def recall_at_k(emb: torch.Tensor, classes: torch.Tensor, ks: Iterable[int] = (10, 100)) -> Dict[str, float]:
    """Leave-one-out Recall@K over one embedding set, the metric Controller.test_epoch_end prints
    (engine/controller.py:77-91).  k larger than N-1 behaves like the reference (all others ranked)."""
    ks = list(ks)
    n = emb.shape[0]
    kmax = max(1, min(max(ks), 100, n - 1))
    if max(ks) > 100:
        raise abi.B200Error('Recall@K is built for K <= 100 (the reference uses 5, 10, 100)')
    idx, _ = cosine_topk(emb, emb, kmax, exclude_self_offset=0)
    hits = recall_hits(idx, classes, classes, ks).tolist()
    valid = int(valid_queries(classes, classes, True).item())
    return {f'Recall@K={k}': (h / valid if valid else float('nan')) for k, h in zip(ks, hits)}


def recall_at_k_rows(emb_all: torch.Tensor, classes_all: torch.Tensor, rows: torch.Tensor, ks: Iterable[int] = (10, 100), group=None):
    """Leave-one-out Recall@K of the WHOLE set `emb_all` (present on every rank) with the queries split over the ranks: this
    rank ranks emb_all[rows] against everything (each query skipping its own row), hit / valid counts are all-reduced.  The
    queries are gathered into a contiguous block in front of the gallery so that leave-one-out is the kernel's offset form."""
    import torch.distributed as dist
    ks = list(ks)
    n = emb_all.shape[0]
    kmax = max(1, min(max(ks), 100, n - 1))
    if max(ks) > 100:
        raise abi.B200Error('Recall@K is built for K <= 100 (the reference uses 5, 10, 100)')
    dev = emb_all.device
    rows = rows.to(dev).long()
    if rows.numel() > 0:
        # gallery = [this rank's rows first, everything else after]: query i excludes gallery row i
        mask = torch.ones(n, dtype=torch.bool, device=dev)
        mask[rows] = False
        order = torch.cat([rows, mask.nonzero().flatten()])
        g, gc = emb_all[order].contiguous(), classes_all[order].contiguous()
        idx = cosine_topk(g[:rows.numel()], g, kmax, exclude_self_offset=0)[0]
        hits = recall_hits(idx, gc[:rows.numel()], gc, ks).to(torch.int64)
        uniq, counts = torch.unique(classes_all, return_counts=True)
        valid = ((counts[torch.searchsorted(uniq, gc[:rows.numel()])] - 1) > 0).sum().to(torch.int64)
    else:
        hits = torch.zeros(len(ks), dtype=torch.int64, device=dev)
        valid = torch.zeros((), dtype=torch.int64, device=dev)
    tot = torch.cat([hits, valid.reshape(1)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(tot, group=group)
    tot = tot.tolist()
    return {f'Recall@K={k}': (h / tot[-1] if tot[-1] else float('nan')) for k, h in zip(ks, tot[:-1])}


# ---------------------------------------------------------------------------------------------------------
# multi-GPU: one process per GPU, torch.distributed for the plumbing (NCCL on GPUs; gloo in the CPU tests)
# ---------------------------------------------------------------------------------------------------------
def _all_gather_rows(x: torch.Tensor, group=None) -> Tuple[torch.Tensor, list]:
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], device=x.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    pad[:x.shape[0]] = x
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0), sizes


def recall_at_k_sharded(emb_local, classes_local, ks=(10, 100), group=None, topk_fn=None, hits_fn=None) -> Dict[str, float]:
    """Leave-one-out Recall@K with the embedding set sharded by rows over the ranks: all-gather the
    embeddings, every rank ranks ITS OWN queries against everything (no merge needed), all-reduce the counts.
    topk_fn / hits_fn default to the CUDA kernels; the CPU gloo tests inject the oracle to exercise the host logic."""
    import torch.distributed as dist
    ks = list(ks)
    rank = dist.get_rank(group)
    emb_all, sizes = _all_gather_rows(emb_local.contiguous().float(), group)
    cls_all, _ = _all_gather_rows(classes_local.contiguous().long(), group)
    start = sum(sizes[:rank])
    kmax = max(1, min(max(ks), 100, emb_all.shape[0] - 1))
    topk_fn = topk_fn or (lambda q, g, k, off: cosine_topk(q, g, k, exclude_self_offset=off)[0])
    hits_fn = hits_fn or recall_hits
    if emb_local.shape[0] > 0:
        idx = topk_fn(emb_local.contiguous().float(), emb_all, kmax, start)
        hits = hits_fn(idx, classes_local.long(), cls_all, ks).to(torch.int64)
        uniq, counts = torch.unique(cls_all, return_counts=True)
        pos = torch.searchsorted(uniq, classes_local.long())
        valid = ((counts[pos] - 1) > 0).sum().to(torch.int64)
    else:
        hits = torch.zeros(len(ks), dtype=torch.int64, device=emb_local.device)
        valid = torch.zeros((), dtype=torch.int64, device=emb_local.device)
    tot = torch.cat([hits.to(emb_local.device), valid.reshape(1).to(emb_local.device)])
    dist.all_reduce(tot, group=group)
    tot = tot.tolist()
    return {f'Recall@K={k}': (h / tot[-1] if tot[-1] else float('nan')) for k, h in zip(ks, tot[:-1])}


def cosine_topk_gallery_sharded(q_local, g_local, k, group=None, topk_fn=None, merge_fn=None):
    """BASELINE config 4 layout: the gallery is sharded by rows, queries are sharded too.  All-gather the queries,
    local fused top-k against the local gallery shard (indices offset to global), all-gather the partial lists and
    merge the lists of this rank's own queries.  Returns (idx int32 [nq_local, k] global gallery rows, score fp64)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    q_all, qsizes = _all_gather_rows(q_local.contiguous().float(), group)
    gcount = torch.tensor([g_local.shape[0]], device=g_local.device, dtype=torch.int64)
    gs = [torch.zeros_like(gcount) for _ in range(world)]
    dist.all_gather(gs, gcount, group=group)
    g_base = sum(int(s.item()) for s in gs[:rank])
    topk_fn = topk_fn or (lambda q, g, kk, base: cosine_topk(q, g, kk, g_index_base=base))
    merge_fn = merge_fn or topk_merge
    kk = min(k, max(1, g_local.shape[0]))
    idx, score = topk_fn(q_all, g_local.contiguous().float(), kk, g_base)
    if kk < k:       # pad short lists so that every rank contributes the same shape
        pad_i = torch.full((idx.shape[0], k - kk), -1, dtype=idx.dtype, device=idx.device)
        pad_s = torch.full((idx.shape[0], k - kk), float('-inf'), dtype=score.dtype, device=idx.device)
        idx, score = torch.cat([idx, pad_i], 1), torch.cat([score, pad_s], 1)
    idx_all = [torch.empty_like(idx) for _ in range(world)]
    sc_all = [torch.empty_like(score) for _ in range(world)]
    dist.all_gather(idx_all, idx.contiguous(), group=group)
    dist.all_gather(sc_all, score.contiguous(), group=group)
    q0 = sum(qsizes[:rank])
    q1 = q0 + qsizes[rank]
    return merge_fn(torch.stack([s[q0:q1] for s in sc_all]), torch.stack([i[q0:q1] for i in idx_all]), k)
