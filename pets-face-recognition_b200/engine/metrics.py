"""torchmetrics-free pair-verification metrics used by Controller (reference engine/controller.py:60-75,
:106-133,:162-180,:205-211 call torchmetrics.AUROC / ROC / AveragePrecision / StatScores; those packages are not part of
this build).  Every function works on the device its inputs live on: with CUDA scores (b200.gallery.pair_similarity) the
sort / scans stay on the GPU and only the final scalars cross to the host."""
import torch


def _sorted_ends(scores: torch.Tensor, labels: torch.Tensor):
    """Scores descending (stable), labels in that order, and the last position of every run of equal scores."""
    scores = scores.detach().double().flatten()
    labels = labels.detach().to(scores.device).flatten().bool()
    order = torch.argsort(scores, descending=True, stable=True)
    s, l = scores[order], labels[order]
    distinct = torch.nonzero(s[1:] != s[:-1]).flatten()
    ends = torch.cat([distinct, torch.tensor([s.numel() - 1], device=s.device)])
    return s, l, ends


def roc(scores: torch.Tensor, labels: torch.Tensor):
    """fpr, tpr, thresholds with thresholds descending over the distinct scores (sklearn / torchmetrics convention,
    with the leading (0, 0) point at threshold max + 1)."""
    s, l, ends = _sorted_ends(scores, labels)
    tps = torch.cumsum(l.double(), 0)[ends]
    fps = (ends + 1).double() - tps
    p, n = l.sum().double().clamp_min(1), (~l).sum().double().clamp_min(1)
    zero = torch.zeros(1, dtype=torch.float64, device=s.device)
    tpr = torch.cat([zero, tps / p])
    fpr = torch.cat([zero, fps / n])
    thr = torch.cat([s[ends][:1] + 1, s[ends]])
    return fpr, tpr, thr


def auroc(scores, labels) -> float:
    fpr, tpr, _ = roc(scores, labels)
    return float(torch.trapz(tpr, fpr))


def average_precision(scores, labels) -> float:
    """sum_n (R_n - R_{n-1}) P_n over the distinct score thresholds (ties form one threshold)."""
    s, l, ends = _sorted_ends(scores, labels)
    l = l.double()
    tp = torch.cumsum(l, 0)[ends]
    prec = tp / (ends + 1).double()
    rec = tp / l.sum().clamp_min(1)
    prev = torch.cat([torch.zeros(1, dtype=torch.float64, device=s.device), rec[:-1]])
    return float(((rec - prev) * prec).sum())


def stat_scores(scores, labels, thr):
    pred = scores.detach().flatten() >= thr
    lab = labels.detach().to(pred.device).flatten().bool()
    counts = torch.stack([(pred & lab).sum(), (pred & ~lab).sum(), (~pred & ~lab).sum(), (~pred & lab).sum()]).tolist()
    return tuple(int(c) for c in counts)
