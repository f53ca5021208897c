"""The metric-learning wrapper of the FE path: backbone -> large-margin head -> loss, with the reference's class name,
constructor keywords and attributes (losses/__init__.py:8-46 there), so configs and `Controller.load_state_dict` keys
(`model_loss.module.*`, `model_loss.add_margin.weight`) carry over.  Head and loss run as one fused native op."""
import torch
from torch import nn

from b200 import ops
from b200.abi import B200Error

from .large_margin import AddMarginProduct, ArcMarginProduct
from .losses import FocalLoss


class SoftmaxBasedMetricLearning(nn.Module):
    """`module` is the backbone, `add_margin` the ArcFace (arc_margin=True) or CosFace head, `focal_loss` the criterion
    (FocalLoss when is_focal, else plain cross entropy).  forward(img) -> embeddings; forward(img, label) ->
    {'loss', 'emb', 'logits'}; `img` may be a list of image batches (their embeddings are concatenated)."""

    def __init__(self, model: nn.Module, num_class, embedding_size=512, s=64.0, m=0.5, is_focal=False, loss_kwargs=None,
                 arc_margin=False, easy_margin=False):
        super().__init__()
        head_cls, head_kw = (ArcMarginProduct, {'easy_margin': easy_margin}) if arc_margin else (AddMarginProduct, {})
        self.add_margin = head_cls(embedding_size, num_class, s=s, m=m, **head_kw)
        options = dict(loss_kwargs or {})
        if is_focal:
            self.focal_loss = FocalLoss(num_class=num_class, **options)
            gamma = float(self.focal_loss.gamma)
        elif options:
            raise B200Error(f'CrossEntropyLoss options {sorted(options)} are not built on the fused path')
        else:
            self.focal_loss, gamma = nn.CrossEntropyLoss(), 0.0
        self._gamma = gamma
        self.module = model
        self.softmax = nn.Softmax(dim=1)

    def _embed(self, img):
        if isinstance(img, (list, tuple)):
            return torch.cat([self.module(part) for part in img], dim=0)
        return self.module(img)

    def forward(self, img, label=None, **__):
        emb = self._embed(img)
        if label is None:
            return emb
        head = self.add_margin
        loss, logits = ops.margin_head(emb, head.weight, label, head.s, head.m, head.kind, head.easy_margin, self._gamma)
        return {'loss': loss, 'emb': emb, 'logits': logits}


class DummyWrapper(nn.Module):
    """Pass-through wrapper with the `.module` attribute the configs expect."""

    def __init__(self, model, *_, **__):
        super().__init__()
        self.module = model

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
