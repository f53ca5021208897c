"""The identity-folder datasets and the verification-pair generator (data_loading/dataset.py, pairs.py) against the outputs
of the reference's own classes on the same little folder tree (tests/golden/dataset_pairs.json, written by
tests/golden/make_golden_dataset.py): item order, labels, decoded pixels, users, and - for the same seed - the very same pairs."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent / 'golden'))
from dataset_tree import build_tree      # noqa: E402


@pytest.fixture(scope='module')
def trees(tmp_path_factory):
    tmp = tmp_path_factory.mktemp('trees')
    return {True: build_tree(tmp / 'cards'), False: build_tree(tmp / 'plain', with_cards=False)}


@pytest.mark.parametrize('case', ['typed_dogs', 'typed_cats', 'simple'])
def test_rec_dataset_and_pairs_match_reference(golden_dir, trees, case):
    from data_loading import PairGenerator, RecDataset, simple_init_dataset
    g = json.loads((golden_dir / 'dataset_pairs.json').read_text())[case]
    kw = {'typed_dogs': dict(type_=1, min_number=2), 'typed_cats': dict(type_=2, min_number=1),
          'simple': dict(type_=None, min_number=3, init_dataset_method=simple_init_dataset)}[case]
    root = trees[case != 'simple']
    ds = RecDataset(root, start_class=5, **kw)
    assert len(ds) == g['len'] > 0
    assert [str(ds.index_to_path[i].relative_to(root)) for i in range(len(ds))] == g['paths']
    assert [ds.index_to_uid[i] for i in range(len(ds))] == g['uids']
    items = [ds[i] for i in range(len(ds))]
    assert [it['label'] for it in items] == g['labels'] and [it['index'] for it in items] == list(range(len(ds)))
    assert [list(np.asarray(it['x']).shape) for it in items] == g['shapes']
    assert [int(np.asarray(it['x']).astype(np.int64).sum()) for it in items] == g['sums']
    assert ds.get_users() == g['users'] and {str(k): v.name for k, v in ds.uid_to_user.items()} == g['uid_to_user']
    assert ds[-1]['index'] == len(ds) - 1
    p = g['pairs']
    pg = PairGenerator(ds, gen_number=p['gen_number'], gen_ratio=1.5, random_seed=11, usr_list=p['users'])
    assert [list(map(int, q)) for q in pg.pairs] == p['pairs']
    assert {str(k): int(v) for k, v in pg.correction.items()} == p['correction']
    assert [list(map(int, q)) for q in pg.corrected_indices] == p['corrected'] and pg.labels.tolist() == p['labels']
    assert set(pg[0]) == {'x1', 'x2', 'label'} and len(pg) == len(p['pairs'])


def test_pair_cache_subset_and_uint8_path(trees, tmp_path):
    from data_loading import PairGenerator, RecDataset, RecSubset, simple_init_dataset, uint8_chw
    ds = RecDataset(trees[False], None, 3, init_dataset_method=simple_init_dataset, val_augmentation=uint8_chw,
                    train_augmentation=lambda im: torch.zeros(1), val_indices=[0, 1, 2])
    x = ds[0]['x']
    assert x.dtype == torch.uint8 and x.dim() == 3 and x.shape[0] == 3           # validation items: uint8 CHW for the u8 gather
    assert ds[5]['x'].shape == (1,)                                              # training items take the training augmentation
    sub = RecSubset(ds, [2, 0], transform=lambda t: t.float() / 255)
    assert len(sub) == 2 and sub[1]['index'] == 0 and sub[0]['x'].dtype == torch.float32
    cache = tmp_path / 'pairs.pkl'
    a = PairGenerator(ds, gen_number=6, gen_ratio=1, path=cache, random_seed=3, usr_list=ds.get_users())
    b = PairGenerator(ds, path=cache)                                            # second construction reads the pickle
    assert cache.exists() and a.pairs == b.pairs and a.correction == b.correction
    with pytest.raises(AssertionError):
        PairGenerator(ds, gen_number=10 ** 6, random_seed=3, usr_list=ds.get_users())
    stray = ds.uid_to_user[0] / 'notes.txt'                                        # a file the reader does not know
    stray.write_text('x')
    try:
        bad = RecDataset(trees[False], None, 3, init_dataset_method=simple_init_dataset)
        with pytest.raises(Exception, match='Unsupported file format'):
            [bad[i] for i in range(len(bad))]
    finally:
        stray.unlink()


def test_real_data_config_loads_on_a_folder_tree(tmp_path, monkeypatch):
    """configs/dog_fe/swin_t_dog_head.py (the reference's fe_dogs_config.py on models.swin_t): split by users, pair
    generator over the validation users, loaders yielding the reference's item format through its torchvision augmentations."""
    from PIL import Image
    from utils import get_config
    root = tmp_path / 'dogs'
    rng = np.random.RandomState(0)
    for u in range(8):
        d = root / f'dog_{u:02d}'
        d.mkdir(parents=True)
        for j in range(3):
            Image.fromarray(rng.randint(0, 256, (230, 240, 3)).astype(np.uint8)).save(d / f'{j}.jpg')
    monkeypatch.setenv('PETS_DOGS_ROOT', str(root))
    monkeypatch.setenv('PETS_PAIRS', '12')
    monkeypatch.chdir(tmp_path)
    cfg = get_config(Path(__file__).resolve().parents[1] / 'pets-face-recognition_b200' / 'configs/dog_fe/swin_t_dog_head.py')
    assert len(cfg.train_users) == 4 and len(cfg.val_users) == 4 and cfg.n_train_classes == 4
    name, pg = cfg.pair_generator(0)
    assert name == 'Val' and len(pg) == 24 and pg.labels.sum() == 12
    assert max(max(p) for p in pg.corrected_indices) < len(cfg.val)
    item = cfg.train[0]
    assert item['x'].shape == (3, 224, 224) and item['x'].dtype == torch.float32 and 0 <= item['label'] < 4
    batch = next(iter(cfg.val_dataloader()))
    assert batch['x'].shape == (12, 3, 230, 240) and set(batch) == {'x', 'label', 'index'}
    assert getattr(cfg.similarity_f, 'b200_kind', None) == 'cosine01'
