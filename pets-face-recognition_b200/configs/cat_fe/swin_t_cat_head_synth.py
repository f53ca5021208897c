"""cat head FE config: Swin-T backbone + ArcFace head on synthetic data.

Same module-as-config format and the same keys as the reference's configs/cat_fe/*.py (e.g.
configs/dog_fe/fe_dogs_config.py:14-162): model(), loss(), optimizer(), *_dataloader(), pair_generator(),
similarity_f(), k, thrs, far_thr, n_epochs, batch sizes, device.  Differences: the backbone is models.swin_t
(the hot path this build accelerates) through the reference's generic model() hook, and the datasets are the
synthetic stand-ins (no image folders offline).  Sizes can be scaled with SYNTH_* environment variables.
"""
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader

from data_loading import SyntheticRecDataset, SyntheticPairs
from losses import SoftmaxBasedMetricLearning
from models import swin_t

seed = 123 + 1000
torch.manual_seed(seed)
np.random.seed(seed)

_n_train_ids = int(os.environ.get('SYNTH_TRAIN_IDS', 1000))
_n_val_ids = int(os.environ.get('SYNTH_VAL_IDS', 200))
_per_id = int(os.environ.get('SYNTH_PER_ID', 4))

train = SyntheticRecDataset(_n_train_ids, _per_id, seed=seed)
val = SyntheticRecDataset(_n_val_ids, _per_id, seed=seed + 1, start_class=_n_train_ids)
__pair_gen = SyntheticPairs(val, int(os.environ.get('SYNTH_PAIRS', 2000)), 1, seed)

n_epochs = int(os.environ.get('SYNTH_EPOCHS', 50))
train_batch_size = int(os.environ.get('SYNTH_BATCH', 64))
test_batch_size = 20

thrs = np.linspace(0.5, 0.99, 6)
far_thr = [0.1, 0.05, 0.03, 0.01, 0.005, 0.001]
k = [5, 10, 100]


def pair_generator(idx):
    if idx == 0:
        return 'Val', __pair_gen
    if idx == 1:
        return 'Val 1', __pair_gen
    raise Exception


def similarity_f(pairs):
    t1 = torch.cat([i[0].unsqueeze(0) for i in pairs], dim=0)
    t2 = torch.cat([i[1].unsqueeze(0) for i in pairs], dim=0)
    return (F.cosine_similarity(t1, t2) + 1) / 2


similarity_f.b200_kind = 'cosine01'      # lets the Controller score the verification pairs on the device (b200_pair_similarity)


def model():
    return swin_t(num_classes=512)


def loss(config, model_):
    _ = config
    return SoftmaxBasedMetricLearning(model=model_, num_class=_n_train_ids, embedding_size=512, is_focal=True, arc_margin=True)


def optimizer(model_):
    params1 = [p for i, p in model_.module.named_parameters() if 'fc' not in i]
    params2 = [p for i, p in model_.module.named_parameters() if 'fc' in i]
    d = [
        {'lr': 10 ** -2 / 2, 'params': params1},
        {'lr': 10 ** -2, 'params': params2},
        {'lr': 10 ** -2, 'params': model_.add_margin.parameters(), 'weight_decay': 1 * (10 ** -4)}
    ]
    optim = torch.optim.SGD(d, 0.01, momentum=0.9)
    sched = torch.optim.lr_scheduler.MultiStepLR(optim, milestones=[35, 45], gamma=0.1)
    return [optim], [sched]


def train_dataloader():
    return DataLoader(train, train_batch_size, shuffle=True, drop_last=True, num_workers=int(os.environ.get('SYNTH_WORKERS', 4)))


def val_dataloader():
    return DataLoader(val, test_batch_size, num_workers=0)


trainer_kwargs = dict(benchmark=True, precision='bf16')

output = Path('results')
output.mkdir(exist_ok=True)
experiment_name = 'cats'
run_name = 'Swin-T cat head, synthetic'

# devices
device = 'cuda:0'
distributed_train = not isinstance(device, str)
world_size = len(device) if distributed_train else None
