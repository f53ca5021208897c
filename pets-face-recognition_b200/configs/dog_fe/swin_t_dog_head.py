"""dog head FE on real identity folders: Swin-T + ArcFace with the reference's recipe (configs/dog_fe/fe_dogs_config.py:
user-level 50/50 split, 10k + 10k verification pairs, SGD momentum 0.9 with lr 5e-3 backbone / 1e-2 head, wd 1e-4 on the
ArcFace weight, MultiStepLR [35, 45], 50 epochs, batch 64 / 20).  Data roots come from the environment:

    PETS_DOGS_ROOT        identity folders (>= 2 images each)                         [required]
    PETS_DOGS_EXTRA_ROOT  extra training identities (>= 3 images each)                [optional]
    PETS_PAIRS            genuine pairs to draw (default 10000; as many impostors)
    PETS_GPU_AUGMENT      1: the train augmentation runs on the GPU over the uint8 batch (data_loading/gpu_augment.py)
"""
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader

from configs.fe_real_data import build
from losses import SoftmaxBasedMetricLearning
from models import swin_t

seed = 123
torch.manual_seed(seed)
np.random.seed(seed)

if 'PETS_DOGS_ROOT' not in os.environ:
    raise Exception('set PETS_DOGS_ROOT to the folder of identity folders (see the docstring of this config)')
globals().update(build(os.environ['PETS_DOGS_ROOT'], os.environ.get('PETS_DOGS_EXTRA_ROOT'), seed, int(os.environ.get('PETS_PAIRS', 10000)),
                       gpu_augment=os.environ.get('PETS_GPU_AUGMENT', '0') == '1'))

n_epochs, train_batch_size, test_batch_size = 50, 64, 20
thrs = np.linspace(0.5, 0.99, 6)
far_thr = [0.1, 0.05, 0.03, 0.01, 0.005, 0.001]
k = [5, 10, 100]


def similarity_f(pairs):
    left = torch.cat([a.unsqueeze(0) for a, _ in pairs], dim=0)
    right = torch.cat([b.unsqueeze(0) for _, b in pairs], dim=0)
    return (F.cosine_similarity(left, right) + 1) / 2


similarity_f.b200_kind = 'cosine01'      # lets the Controller score the verification pairs on the device (b200_pair_similarity)


def model():
    return swin_t(num_classes=512)


def loss(config, model_):
    return SoftmaxBasedMetricLearning(model=model_, num_class=config.n_train_classes, embedding_size=512, is_focal=True, arc_margin=True)


def optimizer(model_):
    named = list(model_.module.named_parameters())
    groups = [{'lr': 5e-3, 'params': [p for n, p in named if 'fc' not in n]},
              {'lr': 1e-2, 'params': [p for n, p in named if 'fc' in n]},
              {'lr': 1e-2, 'params': model_.add_margin.parameters(), 'weight_decay': 1e-4}]
    optim = torch.optim.SGD(groups, 0.01, momentum=0.9)
    return [optim], [torch.optim.lr_scheduler.MultiStepLR(optim, milestones=[35, 45], gamma=0.1)]


def train_dataloader():
    return DataLoader(train, train_batch_size, shuffle=True, drop_last=True, num_workers=4)      # noqa: F821  (from build())


def val_dataloader():
    return DataLoader(val, test_batch_size, num_workers=0)                                       # noqa: F821


trainer_kwargs = dict(benchmark=True, precision='bf16')
output = Path('results')
output.mkdir(exist_ok=True)
experiment_name, run_name = 'Dogs', 'Swin-T dog head'
device = 'cuda:0'
distributed_train = not isinstance(device, str)
world_size = len(device) if distributed_train else None
