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

import itertools
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
        wanted = set(usr_list)
        members = {u: idx for u, idx in self.dataset.uid_to_indices.items() if u in wanted}     # dataset order kept
        n_total = len(self.dataset)
        gen_cap = {u: len(idx) * (len(idx) - 1) for u, idx in members.items()}                  # ordered pairs (i, j), i != j
        imp_cap = {u: n_total * len(idx) - min(n_total, len(idx)) for u, idx in members.items()}
        max_gen, max_imp = sum(gen_cap.values()), sum(imp_cap.values())
        if gen_number is None:
            gen_number = max_gen
        assert gen_number <= max_gen, f'{gen_number} greater than {max_gen}'
        imp_number = int(gen_number * gen_ratio)
        assert imp_number <= max_imp, f'{imp_number} greater than {max_imp}'

        def sample(candidates, quota):
            picked = rng.choice(len(candidates), quota, replace=False)      # one draw per identity, in dataset order
            return [candidates[k] for k in picked]

        listed = {i for idx in members.values() for i in idx}
        pairs = []
        for u, idx in members.items():                                      # genuine pairs first (label 1) ...
            if len(idx) > 1:
                quota = min(round(gen_cap[u] / max_gen * gen_number), gen_cap[u])
                pairs += [(a, b, 1) for a, b in sample(list(itertools.permutations(idx, 2)), quota)]
        for u, idx in members.items():                                      # ... then the impostors (label 0)
            quota = min(round(imp_cap[u] * imp_number / max_imp), imp_cap[u])
            pairs += [(a, b, 0) for a, b in sample(list(itertools.product(idx, listed - set(idx))), quota)]

        self.pairs, self.correction = pairs, self._rank_among(listed)
        if path is not None:
            with open(path, 'wb') as f:
                pickle.dump([self.pairs, self.correction], f)

    @staticmethod
    def _rank_among(listed):
        """dataset index -> position among the listed images.  As in the reference the smallest listed index maps to 0 and
        every later one to its index minus the number of unlisted indices below it."""
        ranks, skipped, previous = {}, 0, None
        for i in sorted(listed):
            skipped = i if previous is None else skipped + (i - previous - 1)
            ranks[i] = 0 if previous is None else i - skipped
            previous = i
        return ranks

    @property
    def labels(self):
        return np.array([int(label) for _, _, label in self.pairs])

    @property
    def indices(self):
        return [(a, b) for a, b, _ in self.pairs]

    @property
    def corrected_indices(self):
        return [(self.correction[a], self.correction[b]) for a, b, _ in self.pairs]
