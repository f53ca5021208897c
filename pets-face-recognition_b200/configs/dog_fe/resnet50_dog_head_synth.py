"""dog head FE config with the backbone the reference's FE configs ship: torchvision's ResNet-50 with a 512-d `fc`
(configs/dog_fe/fe_dogs_config.py:96-109), here `models.resnet50` - the same module tree and state-dict keys, run on the B200
kernels (b200/convnet.py).  Everything else (synthetic identities, ArcFace + focal loss, the reference's SGD groups split on
`'fc' in name`, MultiStepLR, pair / similarity hooks) is the Swin-T synthetic config's, loaded from the file next to this one."""
from importlib.util import module_from_spec, spec_from_file_location
from pathlib import Path

import torch

from models import resnet50

_spec = spec_from_file_location('swin_t_dog_head_synth', Path(__file__).with_name('swin_t_dog_head_synth.py'))
_base = module_from_spec(_spec)
_spec.loader.exec_module(_base)
globals().update({k: v for k, v in vars(_base).items() if not k.startswith('__')})


def model():
    model_ = resnet50()          # pretrained=True in the reference: load_state_dict a torchvision checkpoint when one is at hand
    model_.fc = torch.nn.Linear(2048, 512)
    return model_


run_name = 'ResNet-50 dog head, synthetic'
