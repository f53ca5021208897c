from .synthetic import SyntheticRecDataset, SyntheticPairs

__all__ = ['SyntheticRecDataset', 'SyntheticPairs']
