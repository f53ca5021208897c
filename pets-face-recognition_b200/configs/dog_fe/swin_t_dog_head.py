"""dog head FE config on real identity folders: Swin-T backbone + ArcFace head.

The reference's configs/dog_fe/fe_dogs_config.py with the backbone swapped for models.swin_t: the same augmentation stack
(:17-32), user-level 50/50 train / validation split by a seeded permutation (:41-47), optional extra training set with class
ids after the first (:51-63), 10,000 + 10,000 verification pairs over the validation users (:65), SGD groups and MultiStepLR
(:123-133), loaders (:136-141).  Dataset roots come from the environment (the reference hard-codes ../pets_datasets/...):

    PETS_DOGS_ROOT        identity folders (>= 2 images each)                         [required]
    PETS_DOGS_EXTRA_ROOT  extra training identities (>= 3 images each)                [optional]
    PETS_PAIRS            genuine pairs to draw (default 10000; impostors the same number)
"""
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
import torchvision
from torch.utils.data import ConcatDataset, DataLoader

from data_loading import PairGenerator, RecDataset, RecSubset, simple_init_dataset
from losses import SoftmaxBasedMetricLearning
from models import swin_t

seed = 123
torch.manual_seed(seed)
np.random.seed(seed)

train_augmentation = torchvision.transforms.Compose([
    torchvision.transforms.ToPILImage(),
    torchvision.transforms.RandomAdjustSharpness(0, 0.1),
    torchvision.transforms.RandomAutocontrast(0.3),
    torchvision.transforms.RandomCrop((220, 220)),
    torchvision.transforms.Resize((224, 224)),
    torchvision.transforms.RandomRotation(5),
    torchvision.transforms.ToTensor(),
])

val_augmentation = torchvision.transforms.Compose([
    torchvision.transforms.ToPILImage(),
    torchvision.transforms.ToTensor(),
])

if 'PETS_DOGS_ROOT' not in os.environ:
    raise Exception('set PETS_DOGS_ROOT to the folder of identity folders (see the docstring of this config)')

dataset = RecDataset(Path(os.environ['PETS_DOGS_ROOT']), None, 2, init_dataset_method=simple_init_dataset)

perm = np.random.RandomState(seed).permutation(dataset.get_users())
tr_size = 0.5
train_users = [perm[i] for i in range(int(len(perm) * tr_size))]
val_users = [perm[i] for i in range(int(len(perm) * tr_size), len(perm))]
train_indices = [j for i in train_users for j in dataset.uid_to_indices[i]]
val_indices = [j for i in val_users for j in dataset.uid_to_indices[i]]
assert len(set(train_indices) & set(val_indices)) == 0

train = RecSubset(dataset, train_indices, train_augmentation)
n_train_classes = len(train_users)
if os.environ.get('PETS_DOGS_EXTRA_ROOT'):
    dataset3 = RecDataset(Path(os.environ['PETS_DOGS_EXTRA_ROOT']), None, 3, init_dataset_method=simple_init_dataset,
                          start_class=len(train_users))
    n_train_classes += len(dataset3.get_users())
    train = ConcatDataset((train, RecSubset(dataset3, list(range(len(dataset3))), train_augmentation)))
val = RecSubset(dataset, val_indices, val_augmentation)
for a, b in enumerate(train_users):
    dataset.label_map[b] = a

__n_pairs = int(os.environ.get('PETS_PAIRS', 10000))
__pair_gen = PairGenerator(dataset, __n_pairs, 1, None, seed, val_users)

n_epochs = 50
train_batch_size = 64
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
    return SoftmaxBasedMetricLearning(model=model_, num_class=n_train_classes, embedding_size=512, is_focal=True, arc_margin=True)


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
    return DataLoader(train, train_batch_size, shuffle=True, drop_last=True, num_workers=4)


def val_dataloader():
    return DataLoader(val, test_batch_size, num_workers=0)


trainer_kwargs = dict(benchmark=True, precision='bf16')

output = Path('results')
output.mkdir(exist_ok=True)
experiment_name = 'Dogs'
run_name = 'Swin-T dog head'

# devices
device = 'cuda:0'
distributed_train = not isinstance(device, str)
world_size = len(device) if distributed_train else None
