"""Train entry point - same CLI and call sequence as the reference's main.py:18-93
(get_config -> Controller -> configure_trainer -> trainer.fit).  MLflow is optional (not installed here):
metrics go to stdout and, when a config sets `metrics_jsonl`, to a JSON-lines file."""
import argparse
import json
import shutil
import warnings
from datetime import datetime
from pathlib import Path

from engine import Controller
from utils import is_main_process, configure_trainer, get_config


class JsonlLogger:
    def __init__(self, path):
        self.path = Path(path)

    def log_metrics(self, metrics, step):
        with self.path.open('a') as f:
            f.write(json.dumps({'step': step, **{k: float(v) for k, v in metrics.items()}}) + '\n')


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument('-c', '--config', required=True, type=Path, help='Path to config file')
    return parser.parse_args()


if __name__ == '__main__':
    warnings.simplefilter('ignore')
    args = parse_args()
    config = get_config(args.config)

    checkpoint_path = None
    logger = None
    if is_main_process():
        restime = datetime.now().strftime('%Y%m%d-%H%M%S')
        run_output_root = Path(config.output) / restime
        config.output = run_output_root
        checkpoint_path = run_output_root / 'checkpoints'
        config.checkpoint_path = checkpoint_path
        config.img_dir = run_output_root / 'img'
        checkpoint_path.mkdir(parents=True, exist_ok=True)
        config.img_dir.mkdir(exist_ok=True)
        shutil.copy2(args.config, run_output_root)
        logger = JsonlLogger(run_output_root / 'metrics.jsonl')

    controller = Controller(config=config)
    trainer = configure_trainer(config, logger, checkpoint_path)
    trainer.fit(controller)
    print('Completed!')
