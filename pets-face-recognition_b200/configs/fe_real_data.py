"""Builder of the real-data FE configs: everything the reference's configs/dog_fe/fe_dogs_config.py sets up at import time
(:17-65), as one function of the dataset roots, so that the per-species config modules stay declarative.

Returned names are the reference's config keys: train / val datasets, train_users / val_users, the pair generator hook,
augmentations.  Split: users permuted with RandomState(seed), first half trains; class ids of the training users are
re-numbered 0..n-1 in permutation order (label_map), an optional extra dataset continues the numbering (start_class)."""
from pathlib import Path

import numpy as np
import torchvision.transforms as T
from torch.utils.data import ConcatDataset

from data_loading import PairGenerator, RecDataset, RecSubset, simple_init_dataset


def augmentations():
    """Training: random sharpness removal (p 0.1), autocontrast (p 0.3), 220-crop resized back to 224, rotation within
    5 degrees; validation: the image as it is.  Both end in ToTensor (float CHW in [0, 1])."""
    train = T.Compose([T.ToPILImage(), T.RandomAdjustSharpness(0, 0.1), T.RandomAutocontrast(0.3), T.RandomCrop((220, 220)),
                       T.Resize((224, 224)), T.RandomRotation(5), T.ToTensor()])
    val = T.Compose([T.ToPILImage(), T.ToTensor()])
    return train, val


def uint8_chw(img):
    """Dataset item -> uint8 CHW tensor, untouched (the GPU augmentation and the backbone's first kernel take raw pixels)."""
    import torch
    return img if torch.is_tensor(img) else torch.as_tensor(np.asarray(img)).permute(2, 0, 1).contiguous()


def build(root, extra_root=None, seed=123, n_pairs=10000, train_fraction=0.5, min_images=2, extra_min_images=3, gpu_augment=False):
    """gpu_augment: the training items stay uint8 and un-augmented in the DataLoader; the returned dict carries
    `gpu_train_augmentation`, which engine.Controller.training_step applies to the whole batch on the device."""
    train_aug, val_aug = augmentations()
    if gpu_augment:
        train_aug = uint8_chw
    base = RecDataset(Path(root), None, min_images, init_dataset_method=simple_init_dataset)
    order = np.random.RandomState(seed).permutation(base.get_users())
    cut = int(len(order) * train_fraction)
    train_users, val_users = [order[i] for i in range(cut)], [order[i] for i in range(cut, len(order))]
    train_idx = [j for u in train_users for j in base.uid_to_indices[u]]
    val_idx = [j for u in val_users for j in base.uid_to_indices[u]]
    assert not set(train_idx) & set(val_idx)
    parts = [RecSubset(base, train_idx, train_aug)]
    n_classes = len(train_users)
    if extra_root:
        extra = RecDataset(Path(extra_root), None, extra_min_images, init_dataset_method=simple_init_dataset, start_class=n_classes)
        n_classes += len(extra.get_users())
        parts.append(RecSubset(extra, list(range(len(extra))), train_aug))
    for new_id, user in enumerate(train_users):
        base.label_map[user] = new_id
    pairs = PairGenerator(base, n_pairs, 1, None, seed, val_users)

    def pair_generator(idx):
        if idx in (0, 1):
            return ('Val', 'Val 1')[idx], pairs
        raise Exception

    return dict(dataset=base, train=parts[0] if len(parts) == 1 else ConcatDataset(parts), val=RecSubset(base, val_idx, val_aug),
                train_users=train_users, val_users=val_users, n_train_classes=n_classes, pair_generator=pair_generator,
                train_augmentation=train_aug, val_augmentation=val_aug,
                **({'gpu_train_augmentation': _gpu_aug()} if gpu_augment else {}))


def _gpu_aug():
    from data_loading.gpu_augment import GpuTrainAugmentation
    return GpuTrainAugmentation(size=224, crop=220, p_sharp=0.1, p_autocontrast=0.3, degrees=5.0)
