"""Generate tests/golden/dataset_pairs.json by RUNNING THE REFERENCE's RecDataset / PairGenerator (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_dataset.py

data_loading/dataset.py imports albumentations and pipe (not installed); neither takes part in RecDataset: they are replaced
by minimal stand-ins (pipe.where as the `iterable | where(f)` filter it is).  The tree comes from dataset_tree.build_tree.
"""
import importlib.util
import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.dont_write_bytecode = True
from dataset_tree import build_tree      # noqa: E402

REF = Path('/root/reference')


class _Where:
    def __init__(self, f):
        self.f = f

    def __ror__(self, it):
        return (x for x in it if self.f(x))


def load_reference():
    sys.modules['pipe'] = types.SimpleNamespace(where=_Where)
    sys.modules['albumentations'] = types.SimpleNamespace(bbox_rot90=None, keypoint_rot90=None)
    pkg = types.ModuleType('ref_dl')
    pkg.__path__ = [str(REF / 'data_loading')]
    sys.modules['ref_dl'] = pkg
    mods = {}
    for name in ('dataset', 'pairs'):
        spec = importlib.util.spec_from_file_location(f'ref_dl.{name}', REF / 'data_loading' / f'{name}.py')
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f'ref_dl.{name}'] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods['dataset'], mods['pairs']


def main():
    ds_mod, pairs_mod = load_reference()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        roots = {True: build_tree(Path(tmp) / 'cards'), False: build_tree(Path(tmp) / 'plain', with_cards=False)}
        for case, kw in {'typed_dogs': dict(type_=1, min_number=2), 'typed_cats': dict(type_=2, min_number=1),
                         'simple': dict(type_=None, min_number=3, init_dataset_method=ds_mod.simple_init_dataset)}.items():
            root = roots[case != 'simple']
            ds = ds_mod.RecDataset(root, start_class=5, **kw)
            items = [ds[i] for i in range(len(ds))]
            rec = {'len': len(ds), 'paths': [str(ds.index_to_path[i].relative_to(root)) for i in range(len(ds))],
                   'uids': [ds.index_to_uid[i] for i in range(len(ds))], 'labels': [it['label'] for it in items],
                   'shapes': [list(np.asarray(it['x']).shape) for it in items], 'sums': [int(np.asarray(it['x']).astype(np.int64).sum()) for it in items],
                   'users': ds.get_users(), 'uid_to_user': {str(k): v.name for k, v in ds.uid_to_user.items()}}
            users = ds.get_users()[::2] if case != 'typed_cats' else ds.get_users()
            n_gen = sum(len(i) * len(i) - len(i) for u, i in ds.uid_to_indices.items() if u in set(users))
            if n_gen:
                pg = pairs_mod.PairGenerator(ds, gen_number=max(1, n_gen // 2), gen_ratio=1.5, random_seed=11, usr_list=users)
                rec['pairs'] = {'users': users, 'gen_number': max(1, n_gen // 2), 'pairs': [list(map(int, p)) for p in pg.pairs],
                                'correction': {str(k): int(v) for k, v in pg.correction.items()},
                                'corrected': [list(map(int, p)) for p in pg.corrected_indices], 'labels': pg.labels.tolist()}
            out[case] = rec
            print(case, rec['len'], 'items', len(rec.get('pairs', {}).get('pairs', [])), 'pairs')
    (HERE / 'dataset_pairs.json').write_text(json.dumps(out, indent=0))


if __name__ == '__main__':
    main()
