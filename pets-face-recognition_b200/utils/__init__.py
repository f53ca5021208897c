"""Config loading and trainer construction with the reference's names and config keys (utils/__init__.py:13-134 there),
re-implemented without pytorch-lightning.

A config is a Python module; `get_config(path)` executes it and exposes its public, non-module globals both as attributes
and as mapping items (`cfg.n_epochs`, `cfg['n_epochs']`, `cfg.get('callbacks')`, iteration over the names).  `Config` is a
process-wide singleton as in the reference (a second `get_config` replaces it); `get_dict_wrapper` gives an independent
`DictWrapper` (the TSV scripts load two configs side by side).  `configure_trainer` maps `config.device` / `world_size`
to the GPU list and passes the reference's keyword set to `engine.Trainer`; `get_strategy` answers with the string 'ddp'
where the reference builds a DDPPlugin.
"""
import os
import types
from importlib.util import module_from_spec, spec_from_file_location
from typing import List, Optional, Union

import torch

from engine import Trainer


class DictWrapper:
    """Attribute + mapping view of a config namespace.  Unknown attributes fall through to the underlying dict, which is
    what makes `cfg.get(...)`, `cfg.items()` and `cfg.keys()` work."""

    def __init__(self, d=None):
        for name, value in (d or {}).items():
            if not isinstance(value, types.ModuleType):
                setattr(self, name, value)

    # mapping protocol over the instance dict
    def __getitem__(self, name):
        return self.__dict__[name] if name in self.__dict__ else getattr(self, name)

    def __setitem__(self, name, value):
        setattr(self, name, value)

    def __iter__(self):
        return iter(vars(self))

    def __len__(self):
        return len(vars(self))

    def __getattr__(self, name):          # only reached when normal lookup fails
        return getattr(vars(self), name)

    def __repr__(self):
        return f'{type(self).__name__}: {vars(self)!r}'


class Config(DictWrapper):
    """The one live configuration of the process: constructing it again without arguments hands back the same object."""
    _live = None

    def __new__(cls, *args, **kwargs):
        if cls._live is None:
            cls._live = super().__new__(cls)
            cls._live._filled = False
        return cls._live

    def __init__(self, d=None):
        if not self.__dict__.get('_filled'):
            super().__init__(d)
            self.__dict__['_filled'] = True

    def __iter__(self):
        return (k for k in vars(self) if k != '_filled')

    def __len__(self):
        return len(vars(self)) - 1

    def __repr__(self):
        return 'Config: ' + repr({k: self[k] for k in self})

    @classmethod
    def _forget(cls):
        cls._live = None


def _public_globals(path) -> dict:
    if not os.path.exists(path):
        raise AssertionError(path)
    spec = spec_from_file_location('config', path)
    module = module_from_spec(spec)
    spec.loader.exec_module(module)
    return {name: getattr(module, name) for name in dir(module) if not name.startswith('_')}


def get_dict_wrapper(path) -> DictWrapper:
    return DictWrapper(_public_globals(path))


def get_config(path) -> Config:
    values = _public_globals(path)
    Config._forget()
    return Config(values)


def get_gpus(world_size=1) -> Union[int, List[int]]:
    """The first `world_size` devices a tensor can actually be created on."""
    visible = torch.cuda.device_count()
    assert world_size <= visible, f'Only {visible} are visible'
    assert world_size >= 0
    usable: List[int] = []
    for index in range(visible if world_size else 0):
        try:
            torch.zeros(5, 5, device=f'cuda:{index}')
        except RuntimeError:
            continue
        usable.append(index)
        if len(usable) == world_size:
            return usable
    if world_size == 0:
        return 0
    raise Exception(f'Cannot access {world_size} gpus')


def parse_gpus(cfg) -> Union[int, List[int]]:
    """config.device: 'cpu' -> 0, 'cuda' -> first usable GPU, 'cuda:N' -> [N] (0 / [0] if that index does not exist),
    a list -> those GPUs (distributed_train, world_size must match); an int world_size -> that many GPUs."""
    if cfg.get('distributed_train'):
        if isinstance(cfg.device, list):
            assert cfg.world_size == len(cfg.device), 'Not enough GPUs'
            return cfg.device
        return cfg.world_size
    if cfg.device == 'cpu':
        return 0
    if cfg.device == 'cuda':
        return get_gpus()
    index = int(cfg.device.split(':')[-1])
    if index < torch.cuda.device_count():
        return [index]
    return [0] if torch.cuda.is_available() else 0


def is_main_process() -> bool:
    """Global rank 0.  The reference (utils/__init__.py:105-111) compares the env STRING with the int 0 and so answers False
    whenever LOCAL_RANK is set; that only worked because Lightning spawned the workers after the check.  Here every rank is
    launched torchrun-style with RANK / LOCAL_RANK already set, so main-ness is decided from the rank itself."""
    for name in ('RANK', 'NODE_RANK', 'LOCAL_RANK'):
        if name in os.environ:
            try:
                return int(os.environ[name]) == 0
            except ValueError:
                return False
    return True


def get_strategy(config) -> Optional[str]:
    return 'ddp' if config.get('distributed_train', False) else None


def configure_trainer(config, lightning_logger, lightning_log_dir=None) -> Trainer:
    kwargs = dict(config.get('trainer_kwargs', {}))
    return Trainer(gpus=parse_gpus(config), default_root_dir=lightning_log_dir, strategy=get_strategy(config),
                   max_epochs=config.n_epochs, logger=False if lightning_logger is None else lightning_logger,
                   enable_checkpointing=True, callbacks=config.get('callbacks'), **kwargs)
