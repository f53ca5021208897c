"""Shared body of the eval_fe_*_head_sgd.py entry points: the reference's sequence (eval_fe_dog_head_sgd.py:15-25) -
load the config, build the Controller, load a plain state_dict with strict=False (released checkpoints ship without the
ArcFace weight), run Trainer.test - parameterised by environment variables instead of edited paths."""
import os
import warnings
from pathlib import Path

import torch


def run_fe_eval(default_config: str) -> dict:
    """FE_CONFIG: config module path (default: the synthetic config of the entry point); FE_CKPT: checkpoint to load;
    FE_LOG_DIR: trainer root directory (default ./results, as in the reference)."""
    from engine import Controller
    from utils import configure_trainer, get_config
    warnings.simplefilter('ignore')
    config = get_config(Path(os.environ.get('FE_CONFIG', default_config)))
    controller = Controller(config=config)
    checkpoint = os.environ.get('FE_CKPT')
    if checkpoint:
        controller.load_state_dict(torch.load(Path(checkpoint)), strict=False)
    trainer = configure_trainer(config, False, Path(os.environ.get('FE_LOG_DIR', 'results')))
    metrics = trainer.test(controller)
    print('Completed!')
    return metrics
