"""FocalLoss - drop-in for the reference's losses/losses.py:7-28.  With the configs' default gamma=0 and
alpha=None it is exactly mean cross-entropy.  The arithmetic runs fused with the margin head
(csrc/arcface.cu: margin_ce_kernel); this module only carries the hyper-parameters."""
import torch.nn as nn

from b200.abi import B200Error


class FocalLoss(nn.Module):
    def __init__(self, num_class: int, gamma=0, eps=1e-7, alpha=None):
        super().__init__()
        if alpha:
            raise B200Error('FocalLoss(alpha=...) (learned per-class scale) is not built; no shipped config uses it')
        self.gamma, self.eps, self.adaptive_flag = gamma, eps, False

    def forward(self, input, target):
        raise B200Error('FocalLoss runs fused inside SoftmaxBasedMetricLearning on the B200 path')
