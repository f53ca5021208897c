"""Config loader and trainer factory - drop-in for the reference's utils/__init__.py:13-134 (same names, same config
keys), without pytorch-lightning: get_strategy returns the string 'ddp' instead of a DDPPlugin object."""
import inspect
import os
from contextlib import suppress
from importlib.util import module_from_spec, spec_from_file_location
from typing import List, Optional, Union

import torch

from engine import Trainer


class DictWrapper:
    def __init__(self, d=None):
        if d:
            for k in d:
                if not inspect.ismodule(d[k]):
                    self.__setattr__(k, d[k])

    def __getitem__(self, index):
        return self.__getattribute__(index)

    def __setitem__(self, key, value):
        return self.__setattr__(key, value)

    def __iter__(self):
        return iter(self.__dict__)

    def __len__(self):
        return len(self.__dict__)

    def __repr__(self):
        return 'DictWrapper: ' + repr(self.__dict__)

    def __getattr__(self, item):
        return getattr(self.__dict__, item)      # .get / .items / .keys fall through to the dict


class _SingletonBase(type):
    _instances = {}

    def __call__(cls, *args, **kwargs):
        if cls not in cls._instances:
            cls._instances[cls] = super().__call__(*args, **kwargs)
        return cls._instances[cls]


class Config(DictWrapper, metaclass=_SingletonBase):
    def __repr__(self):
        return 'Config: ' + repr(self.__dict__)


def _exec_config(path):
    assert os.path.exists(path), path
    spec = spec_from_file_location('config', path)
    config = module_from_spec(spec)
    spec.loader.exec_module(config)
    return {k: getattr(config, k) for k in dir(config) if not k.startswith('_')}


def get_dict_wrapper(path) -> DictWrapper:
    return DictWrapper(_exec_config(path))


def get_config(path) -> Config:
    config = _exec_config(path)
    if Config in _SingletonBase._instances:
        del _SingletonBase._instances[Config]
    return Config(config)


def get_gpus(world_size=1) -> Union[int, List[int]]:
    assert world_size <= torch.cuda.device_count(), f'Only {torch.cuda.device_count()} are visible'
    assert world_size >= 0
    if world_size == 0:
        return 0
    gpus = []
    for i in range(torch.cuda.device_count()):
        with suppress(RuntimeError):
            torch.zeros(5, 5, device=f'cuda:{i}')
            gpus.append(i)
        if len(gpus) >= world_size:
            return gpus
    raise Exception(f'Cannot access {world_size} gpus')


def parse_gpus(cfg) -> Union[int, List[int]]:
    if cfg.get('distributed_train'):
        if isinstance(cfg.device, list):
            gpus = cfg.device
            assert cfg.world_size == len(gpus), 'Not enough GPUs'
        else:
            gpus = cfg.world_size
    elif cfg.device == 'cpu':
        gpus = 0
    elif cfg.device == 'cuda':
        gpus = get_gpus()
    else:
        gpus = [int(cfg.device.split(':')[-1])]
        gpus = list({i for i in gpus if i < torch.cuda.device_count()})
        if len(gpus) == 0:
            gpus = 0 if not torch.cuda.is_available() else [0]
    return gpus


def is_main_process() -> bool:
    return all(os.environ.get(i, 0) == 0 for i in ('NODE_RANK', 'LOCAL_RANK'))


def get_strategy(config) -> Optional[str]:
    if config.get('distributed_train', False):
        return 'ddp'
    return None


def configure_trainer(config, lightning_logger, lightning_log_dir=None) -> Trainer:
    return Trainer(
        gpus=parse_gpus(config),
        default_root_dir=lightning_log_dir,
        strategy=get_strategy(config),
        max_epochs=config.n_epochs,
        logger=lightning_logger if lightning_logger is not None else False,
        enable_checkpointing=True,
        callbacks=config.get('callbacks'),
        **config.get('trainer_kwargs', {})
    )
