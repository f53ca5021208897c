"""Identity-folder datasets of the FE path: same classes, constructor arguments, attributes and item format as the reference
(data_loading/dataset.py:13-142, :189-202), so the reference's FE configs (configs/dog_fe/fe_dogs_config.py:34-63) run
unchanged on real data.

Directory format: `root/<identity folder>/<image files>`; an identity folder may carry a `card.json` whose
`pet.animal` field says dog (1) / cat (2) - `init_dataset` keeps the folders of one type with at least `min_number`
readable images, `simple_init_dataset` keeps every folder with enough files.  Images are .jpg / .png / .JPG / jpeg (decoded
to RGB uint8 HWC) or .npy arrays.  Items are {'x': image, 'label': class id + start_class, 'index': dataset index}.

B200 note: with `val_augmentation=uint8_chw` (below) items stay uint8 CHW, a batch crosses PCIe at one byte per sample and
the /255 of ToTensor happens inside the first gather kernel (csrc/elementwise.cu: patch_gather_image_u8).
"""
from __future__ import annotations

import json
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

CARD = 'card.json'
IMAGE_SUFFIXES = ('.jpg', '.png', '.JPG', 'jpeg')      # compared with the last four characters of the file name


def uint8_chw(img: np.ndarray) -> torch.Tensor:
    """HWC uint8 image -> CHW uint8 tensor (the fast input layout of the B200 path; the model divides by 255)."""
    return torch.from_numpy(np.array(img, dtype=np.uint8, copy=True)).permute(2, 0, 1).contiguous()     # PIL's buffer is read-only


def _card_type(folder: Path) -> int:
    with open(folder / CARD, 'r', encoding='utf-8') as fp:
        return int(json.load(fp)['pet']['animal'])


def _content(folder: Path):
    return [f for f in folder.iterdir() if f.name != CARD]


def check_dir(path, type_, min_number) -> bool:
    """An identity folder qualifies when its card says `type_` and it holds at least `min_number` files besides the card."""
    folder = Path(path)
    return folder.is_dir() and _card_type(folder) == type_ and len(_content(folder)) >= min_number


def _opens(path, preprocessor) -> bool:
    try:
        pixels = np.asarray(Image.open(path))
        if preprocessor:
            preprocessor(pixels)
        return True
    except Exception:
        return False


def check(paths, preprocessor=None):
    """The paths that open as images (and survive the preprocessor, if any)."""
    return [p for p in paths if _opens(p, preprocessor)]


def init_dataset(path, type_=1, min_number=3, preprocessor=None, paths_to_exclude=None):
    """Folders of pet type `type_` (card.json) with at least `min_number` readable, non-excluded images."""
    skip = {Path(p).resolve() for p in (paths_to_exclude or ())}
    found = {}
    for folder in (f for f in Path(path).iterdir() if check_dir(f, type_, min_number)):
        images = check([f for f in _content(folder) if f.resolve() not in skip], preprocessor)
        if len(images) >= min_number:
            found[folder] = images
    return found


def simple_init_dataset(path, type_, min_number, *_, **__):
    """Every folder with at least `min_number` files, whatever they are (no card, no decoding check)."""
    listing = {folder: list(folder.iterdir()) for folder in Path(path).iterdir()}
    return {folder: files for folder, files in listing.items() if len(files) >= min_number}


class RecDataset(Dataset):
    def __init__(self, path, type_, min_number, preprocessor=None, train_augmentation=None, val_augmentation=None,
                 init_dataset_method=init_dataset, paths_to_exclude=None, val_indices=None, start_class=0):
        self.user_to_paths = init_dataset_method(path, type_, min_number, preprocessor, paths_to_exclude)
        self.preprocessor = preprocessor
        self.start_class = start_class
        self.train_augmentation = train_augmentation
        self.val_augmentation = val_augmentation
        # identities numbered in folder-name order; samples ordered by (folder name, file name)
        folders = sorted(set(self.user_to_paths), key=lambda f: str(f.name))
        self.uid_to_user = dict(enumerate(folders))
        self.user_to_uid = {folder: uid for uid, folder in self.uid_to_user.items()}
        samples = sorted(((folder, f) for folder, files in self.user_to_paths.items() for f in files),
                         key=lambda s: (str(s[0].name), str(s[1].name)))
        self.index_to_uid = {i: self.user_to_uid[folder] for i, (folder, _) in enumerate(samples)}
        self.index_to_path = {i: f for i, (_, f) in enumerate(samples)}
        by_uid = defaultdict(list)
        for i, uid in self.index_to_uid.items():
            by_uid[uid].append(i)
        self.uid_to_indices = dict(by_uid)
        self.val_indices = val_indices
        self.label_map = {uid: k for k, uid in enumerate(self.uid_to_user.keys())}

    def _read(self, path: Path) -> np.ndarray:
        tail = path.name[-4:]
        if tail in IMAGE_SUFFIXES:
            return np.asarray(Image.open(path).convert('RGB'))
        if tail == '.npy':
            return np.load(path)
        raise Exception('Unsupported file format')

    def __getitem__(self, item):
        if item < 0:
            item += len(self)
        img = self._read(self.index_to_path[item])
        label = self.label_map[self.index_to_uid[item]] + self.start_class
        if self.preprocessor:
            img = self.preprocessor(img)
        is_train_item = self.val_indices is None or item not in self.val_indices
        if is_train_item and self.train_augmentation:
            img = self.train_augmentation(img)
        elif self.val_augmentation:
            img = self.val_augmentation(img)
        return {'x': img, 'label': label, 'index': item}

    def __len__(self):
        return len(self.index_to_path)

    def get_users(self):
        return list(self.user_to_uid.values())

    @property
    def val_indices(self):
        return self._val_indices

    @val_indices.setter
    def val_indices(self, value):
        self._val_indices = set(value) if value is not None else None


class RecSubset(Dataset):
    """A view on `indices` of a dataset, with an optional extra transform of 'x'."""

    def __init__(self, dataset, indices, transform=None):
        self.dataset, self.indices, self.transform = dataset, indices, transform

    def __getitem__(self, item):
        data = self.dataset[self.indices[item]]
        if self.transform:
            data['x'] = self.transform(data['x'])
        return data

    def __len__(self):
        return len(self.indices)
