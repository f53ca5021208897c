"""Synthetic stand-ins for the reference's RecDataset / PairGenerator (data_loading/dataset.py:67-142,
data_loading/pairs.py:31-108), which read identity-per-directory image folders that are not available offline and
are out of scope for this build.  They keep the item / attribute interface the
Controller consumes: items are {'x': float image in [0,1], 'label': int, 'index': int}; the pair object exposes
`corrected_indices`, `labels` and `__len__`."""
import numpy as np
import torch
from torch.utils.data import Dataset


class SyntheticRecDataset(Dataset):
    def __init__(self, n_identities: int, per_identity: int, image_size: int = 224, noise: float = 0.15, seed: int = 123,
                 start_class: int = 0):
        g = torch.Generator().manual_seed(seed)
        self.n_identities, self.per_identity, self.image_size, self.noise = n_identities, per_identity, image_size, noise
        # low-resolution identity templates, upsampled on access: same-identity images share structure
        self.templates = torch.rand(n_identities, 3, 14, 14, generator=g)
        self.seeds = torch.randint(0, 2 ** 31 - 1, (n_identities * per_identity,), generator=g)
        self.labels = (torch.arange(n_identities * per_identity) % n_identities) + start_class
        self.start_class = start_class
        self.uid_to_indices = {int(u) + start_class: [int(i) for i in torch.nonzero(self.labels == u + start_class).flatten()]
                               for u in range(n_identities)}

    def get_users(self):
        return sorted(self.uid_to_indices)

    def __len__(self):
        return self.labels.numel()

    def __getitem__(self, i):
        ident = int(self.labels[i]) - self.start_class
        g = torch.Generator().manual_seed(int(self.seeds[i]))
        base = torch.nn.functional.interpolate(self.templates[ident][None], size=self.image_size, mode='bilinear',
                                               align_corners=False)[0]
        x = (base + self.noise * torch.randn(3, self.image_size, self.image_size, generator=g)).clamp_(0, 1)
        return {'x': x, 'label': int(self.labels[i]), 'index': int(i)}


class SyntheticPairs:
    """Seeded genuine / impostor index pairs over a dataset's items (the role of PairGenerator; indices are already
    0..N-1 here, so `corrected_indices` is the identity correction)."""

    def __init__(self, dataset: SyntheticRecDataset, gen_number: int, gen_ratio: float = 1.0, seed: int = 123):
        rand = np.random.RandomState(seed)
        labels = dataset.labels.numpy()
        n = len(labels)
        pairs = []
        users = [u for u, idx in dataset.uid_to_indices.items() if len(idx) > 1]
        for _ in range(gen_number):
            idx = dataset.uid_to_indices[users[rand.randint(len(users))]]
            a, b = rand.choice(len(idx), 2, replace=False)
            pairs.append((idx[a], idx[b], 1))
        for _ in range(int(gen_number * gen_ratio)):
            while True:
                a, b = rand.randint(n), rand.randint(n)
                if labels[a] != labels[b]:
                    break
            pairs.append((int(a), int(b), 0))
        self.pairs = pairs

    def __len__(self):
        return len(self.pairs)

    @property
    def labels(self):
        return np.array([int(i) for _, _, i in self.pairs])

    @property
    def indices(self):
        return [(i, j) for i, j, _ in self.pairs]

    @property
    def corrected_indices(self):
        return self.indices
