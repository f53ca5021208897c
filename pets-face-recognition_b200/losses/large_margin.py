"""Large-margin heads - drop-in for the reference's losses/large_margin.py (same class names, ctor
arguments, ``weight`` parameter and its Xavier init), computing on the B200 kernels:
unit-row normalisation -> tcgen05 cosine GEMM with the margin + scale epilogue (csrc/gemm.cu EpiMargin).

Deviation (documented in DESIGN.md): the reference takes sqrt(1 - cos^2) unclamped and yields NaN rows
when rounding makes |cos| > 1 (losses/large_margin.py:72); here 1 - cos^2 is clamped at 0.
"""
import math

import torch
import torch.nn as nn
from torch.nn import Parameter

from b200 import ops


class _MarginProduct(nn.Module):
    kind = 0

    def __init__(self, in_features, out_features, s, m, easy_margin=False):
        super().__init__()
        self.in_features, self.out_features, self.s, self.m = in_features, out_features, s, m
        self.easy_margin = easy_margin
        self.weight = Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, input, label):
        """Margin logits (B, out_features), as the reference's forward returns them.  Not differentiable on its own:
        training goes through SoftmaxBasedMetricLearning, which fuses head + loss + their gradient."""
        with torch.no_grad():
            _, logits = ops.margin_head(input, self.weight, label, self.s, self.m, self.kind, self.easy_margin, 0.0)
        return logits


class AddMarginProduct(_MarginProduct):
    """cos(theta) - m  (reference losses/large_margin.py:10-40)."""
    kind = 1

    def __init__(self, in_features, out_features, s=30.0, m=0.40, device=None, **_):
        super().__init__(in_features, out_features, s, m)
        self.device = device


class ArcMarginProduct(_MarginProduct):
    """cos(theta + m)  (reference losses/large_margin.py:44-84)."""
    kind = 0

    def __init__(self, in_features, out_features, s=30.0, m=0.50, easy_margin=False, **_):
        super().__init__(in_features, out_features, s, m, easy_margin)
        self.cos_m, self.sin_m = math.cos(m), math.sin(m)
        self.th, self.mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
