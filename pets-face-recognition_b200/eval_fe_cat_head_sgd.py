"""Eval entry point - same call sequence as the reference's eval_fe_cat_head_sgd.py:15-25
(get_config -> Controller -> load_state_dict(strict=False) -> configure_trainer -> trainer.test)."""
import os
import warnings
from pathlib import Path

import torch

from engine import Controller
from utils import configure_trainer, get_config

if __name__ == '__main__':
    warnings.simplefilter('ignore')
    lightning_logger = False
    checkpoint_path = Path('results')
    cfg_path = Path(os.environ.get('FE_CONFIG', 'configs/cat_fe/swin_t_cat_head_synth.py'))
    config = get_config(cfg_path)
    controller = Controller(config=config)
    ckpt = os.environ.get('FE_CKPT')
    if ckpt:
        controller.load_state_dict(torch.load(Path(ckpt)), strict=False)
    trainer = configure_trainer(config, lightning_logger, checkpoint_path)
    trainer.test(controller)
    print('Completed!')
