"""Submission TSV from per-folder head AND body embeddings (reference generate_tsv_to_reproduce1.py:88-136, :228-247): the
head + body ensemble rule - the body score stands in when the enroll folder has no head vector, or when the head score is 0 and the
body score clears the pet type's threshold (0.9069641 dogs / 0.985643 cats, :106).

As for script 2, the reference's front end (head / body detection, alignment, the FE models per crop, :140-226) is outside this
build's hot path; the database of embeddings is the input:

    python generate_tsv_to_reproduce1.py --db embeddings.pt --out pred_scores_test1.tsv [--preds preds.tsv]

`embeddings.pt` (torch.save) maps each big folder to (init_db, extra_db); a db maps a folder name to
{'head_vectors': [...], 'body_vectors': [...], 'type': 1 (dog) | 2 (cat)} as process_base builds it (:31-60).  Both score tables
run on the tcgen05 GEMM (b200/multivector.py: calc_scores_ensemble).
"""
import argparse

import torch

from b200.multivector import backfill, calc_scores_ensemble, create_table, write_tsv  # noqa: F401


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--db', required=True)
    ap.add_argument('--out', default='pred_scores_test1.tsv')
    ap.add_argument('--preds', default=None, help='preds.tsv whose rows fill the queries that got no prediction')
    args = ap.parse_args()
    db = torch.load(args.db, weights_only=False)   # trusted input: keys are pathlib.Path objects
    df = create_table(db, ensemble=True)
    write_tsv(df, args.out)
    if args.preds:
        backfill(args.out, args.preds)
    print(f'{len(df)} rows -> {args.out}')


if __name__ == '__main__':
    main()
