"""ResNet-50 behind the reference's `config.model()` hook (configs/dog_fe/fe_dogs_config.py:96-109):

    model_ = models.resnet50()                # same module tree / state-dict keys as torchvision.models.resnet50
    model_.fc = torch.nn.Linear(2048, 512)

The modules only hold parameters and buffers (so torchvision checkpoints load unchanged and the reference's optimizer groups by
parameter name - `'fc' in name` - work); `forward` runs the whole network on the B200 kernels (b200/convnet.py).  No eager
fallback: a CPU tensor or a missing library raises."""
import torch
import torchvision
from torchvision.models.resnet import Bottleneck

from b200.convnet import ConvNetEngine, ConvNetFunction


class ResNet(torchvision.models.ResNet):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._convnet = None

    @property
    def convnet(self) -> ConvNetEngine:
        if self._convnet is None:
            object.__setattr__(self, "_convnet", ConvNetEngine(self))      # not a sub-module: it only points back at this one
        return self._convnet

    def _forward_impl(self, x):
        params = [p for p in self.parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return ConvNetFunction.apply(self.convnet, x, *params)
        return self.convnet.forward(x, False)


def resnet50(pretrained=False, **kwargs):
    """pretrained weights are not downloadable here (no network): load a torchvision state dict with load_state_dict instead"""
    if pretrained:
        raise RuntimeError('resnet50(pretrained=True) needs a download; build the model and load_state_dict a torchvision checkpoint')
    return ResNet(Bottleneck, [3, 4, 6, 3], **kwargs)
