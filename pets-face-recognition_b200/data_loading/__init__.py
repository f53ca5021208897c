from .dataset import RecDataset, RecSubset, init_dataset, simple_init_dataset, uint8_chw
from .pairs import PairGenerator
from .synthetic import SyntheticRecDataset, SyntheticPairs

__all__ = ['RecDataset', 'RecSubset', 'PairGenerator', 'init_dataset', 'simple_init_dataset', 'uint8_chw', 'SyntheticRecDataset', 'SyntheticPairs']
