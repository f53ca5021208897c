"""FocalLoss - drop-in for the reference's losses/losses.py:7-28.  With the configs' default gamma=0 and
alpha=None it is exactly mean cross-entropy.  Inside SoftmaxBasedMetricLearning the arithmetic runs fused with the margin
head (csrc/arcface.cu: margin_ce_kernel); called on its own - `criterion(logits, target)` as the reference's forward allows -
it runs the stand-alone kernel of the same file (focal_rows_kernel), differentiable wrt the logits."""
import torch.nn as nn

from b200 import ops
from b200.abi import B200Error


class FocalLoss(nn.Module):
    def __init__(self, num_class: int, gamma=0, eps=1e-7, alpha=None):
        super().__init__()
        if alpha:
            raise B200Error('FocalLoss(alpha=...) (learned per-class scale) is not built; no shipped config uses it')
        self.gamma, self.eps, self.adaptive_flag = gamma, eps, False

    def forward(self, input, target):
        """(B, C) logits, (B,) int64 targets -> scalar mean focal loss (losses/losses.py:22-28)."""
        return ops.focal_loss(input, target, float(self.gamma))
