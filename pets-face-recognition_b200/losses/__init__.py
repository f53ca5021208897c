"""Drop-in for the reference's losses/__init__.py: SoftmaxBasedMetricLearning = backbone -> margin head -> loss."""
import torch
import torch.nn as nn

from b200 import ops
from b200.abi import B200Error

from .large_margin import ArcMarginProduct, AddMarginProduct
from .losses import FocalLoss


class SoftmaxBasedMetricLearning(nn.Module):
    """Same constructor and attributes (.module, .add_margin, .focal_loss) as the reference
    (losses/__init__.py:8-35); forward(img, label=None) returns the embeddings without a label and
    {'loss', 'emb', 'logits'} with one (:37-46)."""

    def __init__(self, model: nn.Module, num_class, embedding_size=512, s=64.0, m=0.5, is_focal=False, loss_kwargs=None,
                 arc_margin=False, easy_margin=False):
        super().__init__()
        if arc_margin:
            self.add_margin = ArcMarginProduct(embedding_size, num_class, s=s, m=m, easy_margin=easy_margin)
        else:
            self.add_margin = AddMarginProduct(embedding_size, num_class, s=s, m=m)
        loss_kwargs = loss_kwargs or {}
        if is_focal:
            self.focal_loss = FocalLoss(num_class=num_class, **loss_kwargs)
            self._gamma = float(self.focal_loss.gamma)
        else:
            if loss_kwargs:
                raise B200Error(f'CrossEntropyLoss options {sorted(loss_kwargs)} are not built on the fused path')
            self.focal_loss = nn.CrossEntropyLoss()
            self._gamma = 0.0
        self.module = model
        self.softmax = nn.Softmax(dim=1)

    def forward(self, img, label=None, **__):
        if isinstance(img, (list, tuple)):
            tensor = torch.cat([self.module(i) for i in img], dim=0)
        else:
            tensor = self.module(img)
        if label is None:
            return tensor
        h = self.add_margin
        loss, logits = ops.margin_head(tensor, h.weight, label, h.s, h.m, h.kind, h.easy_margin, self._gamma)
        return {'loss': loss, 'emb': tensor, 'logits': logits}


class DummyWrapper(nn.Module):
    def __init__(self, model, *_, **__):
        super().__init__()
        self.module = model

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
