"""torchmetrics-free pair-verification metrics used by Controller (reference engine/controller.py:60-75,
:106-133,:162-180,:205-211 call torchmetrics.AUROC/ROC/...; those packages are not part of this build)."""
import torch


def roc(scores: torch.Tensor, labels: torch.Tensor):
    """fpr, tpr, thresholds with thresholds descending over the distinct scores (sklearn / torchmetrics convention,
    with the leading (0, 0) point at threshold max + 1)."""
    scores = scores.detach().double().cpu().flatten()
    labels = labels.detach().cpu().flatten().bool()
    order = torch.argsort(scores, descending=True, stable=True)
    s, l = scores[order], labels[order]
    distinct = torch.nonzero(s[1:] != s[:-1]).flatten()
    ends = torch.cat([distinct, torch.tensor([s.numel() - 1])])
    tps = torch.cumsum(l.double(), 0)[ends]
    fps = (ends + 1).double() - tps
    p, n = l.sum().double().clamp_min(1), (~l).sum().double().clamp_min(1)
    tpr = torch.cat([torch.zeros(1, dtype=torch.float64), tps / p])
    fpr = torch.cat([torch.zeros(1, dtype=torch.float64), fps / n])
    thr = torch.cat([s[ends][:1] + 1, s[ends]])
    return fpr, tpr, thr


def auroc(scores, labels) -> float:
    fpr, tpr, _ = roc(scores, labels)
    return float(torch.trapz(tpr, fpr))


def average_precision(scores, labels) -> float:
    """sum_n (R_n - R_{n-1}) P_n over the distinct score thresholds (ties form one threshold)."""
    scores = scores.detach().double().cpu().flatten()
    labels = labels.detach().cpu().flatten().bool()
    order = torch.argsort(scores, descending=True, stable=True)
    s, l = scores[order], labels[order].double()
    distinct = torch.nonzero(s[1:] != s[:-1]).flatten()
    ends = torch.cat([distinct, torch.tensor([s.numel() - 1])])
    tp = torch.cumsum(l, 0)[ends]
    prec = tp / (ends + 1).double()
    rec = tp / l.sum().clamp_min(1)
    prev = torch.cat([torch.zeros(1, dtype=torch.float64), rec[:-1]])
    return float(((rec - prev) * prec).sum())


def stat_scores(scores, labels, thr):
    pred = scores.detach().cpu().flatten() >= thr
    lab = labels.detach().cpu().flatten().bool()
    tp = int((pred & lab).sum()); fp = int((pred & ~lab).sum())
    tn = int((~pred & ~lab).sum()); fn = int((~pred & lab).sum())
    return tp, fp, tn, fn
