"""Submission TSV from per-folder head embeddings (reference generate_tsv_to_reproduce2.py:90-136, :228-247).

The reference script also detects and aligns heads and runs the FE models per image (:140-226); that front end is outside
this build's hot path (detection / preprocessing are out of scope), so the database of embeddings is
taken as input:

    python generate_tsv_to_reproduce2.py --db embeddings.pt --out pred_scores_test2.tsv [--preds preds.tsv]

`embeddings.pt` (torch.save) maps each big folder to (init_db, extra_db); a db maps a folder name to
{'head_vectors': [tensor, ...], 'type': 1 (dog) | 2 (cat)} exactly as process_base builds it (:31-52).  Scoring runs on the
B200 gallery kernels (b200/multivector.py); --synthetic N writes a table for N synthetic enroll folders instead.
"""
import argparse
from pathlib import Path

import torch

from b200 import multivector
from b200.multivector import backfill, calc_scores, create_table, write_tsv  # noqa: F401  (the reference's function names)


def synthetic_db(n_enroll: int, seed: int = 123, dim: int = 512):
    g = torch.Generator().manual_seed(seed)
    n_ids = max(4, n_enroll)
    centres = torch.randn(n_ids, dim, generator=g)

    def make(n, prefix):
        db = {}
        for s in range(n):
            ident = s % n_ids
            nvec = int(torch.randint(1, 5, (1,), generator=g).item())
            db[Path(f'{prefix}{s:05d}')] = {'head_vectors': [(centres[ident] + 0.6 * torch.randn(dim, generator=g)).reshape(1, dim) for _ in range(nvec)],
                                            'type': 1 + ident % 2}
        return db
    return {Path('synthetic'): (make(n_enroll, 'q'), make(8 * n_enroll, 'g'))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--db')
    ap.add_argument('--synthetic', type=int, default=0)
    ap.add_argument('--out', default='pred_scores_test2.tsv')
    ap.add_argument('--preds', default=None, help='preds.tsv whose rows fill the queries that got no prediction')
    ap.add_argument('--strategy', default='mean', choices=['mean', 'max'])
    args = ap.parse_args()
    db = synthetic_db(args.synthetic) if args.synthetic else torch.load(args.db, weights_only=False)   # trusted input: keys are pathlib.Path objects
    df = create_table(db, strategy=args.strategy)
    write_tsv(df, args.out)
    if args.preds:
        backfill(args.out, args.preds)
    print(f'{len(df)} rows -> {args.out}')


if __name__ == '__main__':
    main()
