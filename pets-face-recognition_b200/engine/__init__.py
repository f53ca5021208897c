from .controller import Controller
from .trainer import Trainer

__all__ = ['Controller', 'Trainer']
