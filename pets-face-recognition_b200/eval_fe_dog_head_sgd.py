"""Evaluate the dog-head feature extractor (pair metrics + Recall@K=10/100): drop-in for the reference's
eval_fe_dog_head_sgd.py.  Paths come from FE_CONFIG / FE_CKPT / FE_LOG_DIR (see engine/evaluate.py)."""
from engine.evaluate import run_fe_eval

if __name__ == '__main__':
    run_fe_eval('configs/dog_fe/swin_t_dog_head_synth.py')
