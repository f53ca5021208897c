"""Verification pairs over an identity dataset: the reference's PairGenerator (data_loading/pairs.py:10-108) - same
constructor, pickle cache format ([pairs, correction]), properties, and the same pairs for the same seed (the draws from
numpy's RandomState happen in the same order with the same arguments).

Genuine pairs: every identity of `usr_list` with at least two images contributes ordered pairs (i, j), i != j, in proportion
to its share of all possible ones.  Impostor pairs: (image of the identity, image of another listed identity), the quota in
proportion to len(dataset) * images - min(len(dataset), images).  `correction` maps the dataset indices of the listed
identities to their rank among them: the position of the image in an embedding matrix extracted from that subset
(`corrected_indices`, used by Controller for pair scoring, engine/controller.py:62).
"""
from __future__ import annotations

import pickle
from pathlib import Path

import numpy as np
from torch.utils.data import Dataset


class PairGenerator(Dataset):
    def __init__(self, dataset, gen_number=None, gen_ratio=1, path=None, random_seed=None, usr_list=None):
        self.dataset = dataset
        if path is not None and Path(path).exists():
            with open(path, 'rb') as f:
                self.pairs, self.correction = pickle.load(f)
        else:
            self.generate_pairs(gen_number, gen_ratio, path, random_seed, usr_list)

    def __getitem__(self, item):
        first, second, label = self.pairs[item]
        return {'x1': self.dataset[first]['x'], 'x2': self.dataset[second]['x'], 'label': int(label)}

    def __len__(self):
        return len(self.pairs)

    def generate_pairs(self, gen_number, gen_ratio, path, random_seed, usr_list):
        rng = np.random.RandomState(random_seed)
        n_total = len(self.dataset)
        users = set(usr_list)
        members = {u: idx for u, idx in self.dataset.uid_to_indices.items() if u in users}     # dataset order kept

        def genuine_capacity(idx):
            return len(idx) * len(idx) - len(idx)

        def impostor_capacity(idx):
            return n_total * len(idx) - min(n_total, len(idx))

        max_gen = sum(genuine_capacity(idx) for idx in members.values())
        max_imp = sum(impostor_capacity(idx) for idx in members.values())
        if gen_number is None:
            gen_number = max_gen
        else:
            assert gen_number <= max_gen, f'{gen_number} greater than {max_gen}'
        imp_number = int(gen_number * gen_ratio)
        assert imp_number <= max_imp, f'{imp_number} greater than {max_imp}'

        genuine = []
        for u, idx in members.items():
            if len(idx) < 2:
                continue
            cap = genuine_capacity(idx)
            quota = min(round(cap / max_gen * gen_number), cap)
            candidates = [(a, b) for a in idx for b in idx if a != b]
            genuine.extend(candidates[k] for k in rng.choice(len(candidates), quota, replace=False))

        listed = {i for idx in members.values() for i in idx}
        impostor = []
        for u, idx in members.items():
            cap = impostor_capacity(idx)
            quota = min(round(cap * imp_number / max_imp), cap)
            others = listed - set(idx)
            candidates = [(a, b) for a in idx for b in others]
            impostor.extend(candidates[k] for k in rng.choice(len(candidates), quota, replace=False))

        # dataset index -> rank among the listed images.  (As in the reference, the smallest listed index maps to 0 and every
        # later one to its index minus the number of unlisted indices below it.)
        correction = {i: 0 for i in listed}
        skipped, previous = 0, None
        for i in sorted(correction):
            if previous is None:
                skipped = i
            else:
                skipped += i - previous - 1
                correction[i] = i - skipped
            previous = i

        pairs = [(a, b, 1) for a, b in genuine] + [(a, b, 0) for a, b in impostor]
        if path is not None:
            with open(path, 'wb') as f:
                pickle.dump([pairs, correction], f)
        self.pairs, self.correction = pairs, correction

    @property
    def labels(self):
        return np.array([int(label) for _, _, label in self.pairs])

    @property
    def indices(self):
        return [(a, b) for a, b, _ in self.pairs]

    @property
    def corrected_indices(self):
        return [(self.correction[a], self.correction[b]) for a, b, _ in self.pairs]
