"""CPU restatement of the margin heads, the loss and the optimizer step.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch


def _unit_rows(x: torch.Tensor) -> torch.Tensor:
    # F.normalize(x) as used at losses/large_margin.py:32,71: x / max(||x||_2, 1e-12), dim=1
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def arcface_logits(emb, weight, label, s=64.0, m=0.5, easy_margin=False, clamp_sine=False):
    """ArcMarginProduct.forward, losses/large_margin.py:69-84 (defaults s=64, m=0.5 come from
    the wrapper, losses/__init__.py:15-20,24).

    clamp_sine=False reproduces the reference exactly, including its NaN when fp rounding
    makes |cos| > 1 (sqrt of a negative, :72).  clamp_sine=True is the documented deviation
    of the CUDA path (1 - cos^2 clamped at 0)."""
    cos = _unit_rows(emb) @ _unit_rows(weight).t()
    one_minus = 1.0 - cos * cos
    if clamp_sine:
        one_minus = one_minus.clamp_min(0.0)
    sin = torch.sqrt(one_minus)
    phi = cos * math.cos(m) - sin * math.sin(m)
    if easy_margin:
        phi = torch.where(cos > 0, phi, cos)
    else:
        phi = torch.where(cos > math.cos(math.pi - m), phi, cos - math.sin(math.pi - m) * m)
    hot = torch.zeros_like(cos)
    hot[torch.arange(cos.shape[0]), label.long()] = 1.0
    return s * (hot * phi + (1.0 - hot) * cos)


def cosface_logits(emb, weight, label, s=64.0, m=0.5):
    """AddMarginProduct.forward, losses/large_margin.py:30-40."""
    cos = _unit_rows(emb) @ _unit_rows(weight).t()
    hot = torch.zeros_like(cos)
    hot[torch.arange(cos.shape[0]), label.long()] = 1.0
    return s * (hot * (cos - m) + (1.0 - hot) * cos)


def focal_loss(logits, label, gamma=0.0):
    """FocalLoss.forward with alpha=None, losses/losses.py:22-28.  gamma=0 is mean CE."""
    lse = torch.logsumexp(logits, dim=1)
    nll = lse - logits[torch.arange(logits.shape[0]), label.long()]
    p = torch.exp(-nll)
    return ((1.0 - p) ** gamma * nll).mean()


def metric_learning_forward(backbone, emb_weight, img, label=None, *, s=64.0, m=0.5,
                            arc_margin=True, easy_margin=False, gamma=0.0, clamp_sine=False):
    """SoftmaxBasedMetricLearning.forward, losses/__init__.py:37-46.
    ``backbone`` is a callable img -> (B, E)."""
    if isinstance(img, (list, tuple)):
        emb = torch.cat([backbone(i) for i in img], dim=0)
    else:
        emb = backbone(img)
    if label is None:
        return emb
    if arc_margin:
        logits = arcface_logits(emb, emb_weight, label, s, m, easy_margin, clamp_sine)
    else:
        logits = cosface_logits(emb, emb_weight, label, s, m)
    return {'loss': focal_loss(logits, label, gamma), 'emb': emb, 'logits': logits}


def sgd_momentum_step(params: List[torch.Tensor], grads: List[torch.Tensor],
                      bufs: List[Optional[torch.Tensor]], lr: float, momentum: float = 0.9,
                      weight_decay: float = 0.0) -> List[torch.Tensor]:
    """torch.optim.SGD(momentum=0.9, dampening=0, nesterov=False) as configured at
    configs/dog_fe/fe_dogs_config.py:123-133: g += wd*p ; buf = g on the first step, else
    momentum*buf + g ; p -= lr*buf.  Updates ``params`` in place, returns the new buffers."""
    out = []
    for p, g, b in zip(params, grads, bufs):
        g = g + weight_decay * p if weight_decay != 0 else g
        b = g.clone() if b is None else momentum * b + g
        p.sub_(lr * b)
        out.append(b)
    return out


def multistep_lr(base_lr: float, epoch: int, milestones: Iterable[int] = (35, 45), gamma: float = 0.1) -> float:
    """MultiStepLR stepped once per epoch, configs/dog_fe/fe_dogs_config.py:132."""
    return base_lr * gamma ** sum(1 for ms in milestones if epoch >= ms)
