"""Evaluate the cat-head feature extractor (pair metrics + Recall@K=10/100): drop-in for the reference's
eval_fe_cat_head_sgd.py.  Paths come from FE_CONFIG / FE_CKPT / FE_LOG_DIR (see engine/evaluate.py)."""
from engine.evaluate import run_fe_eval

if __name__ == '__main__':
    run_fe_eval('configs/cat_fe/swin_t_cat_head_synth.py')
