"""Deterministic little identity-folder tree shared by make_golden_dataset.py (reference run) and tests/test_dataset_cpu.py."""
import json
from pathlib import Path

import numpy as np
from PIL import Image


def build_tree(root: Path, seed: int = 7, with_cards: bool = True):
    """12 identity folders (names not in creation order), 1..5 files each (.jpg / .png / .npy mixed), with_cards: a card.json with
    the pet type (1 dog / 2 cat) per folder, plus one folder whose files are not images."""
    rng = np.random.RandomState(seed)
    root.mkdir(parents=True, exist_ok=True)
    names = [f'pet_{(i * 7) % 12:02d}' for i in range(12)]
    for k, name in enumerate(names):
        d = root / name
        d.mkdir()
        if with_cards:
            (d / 'card.json').write_text(json.dumps({'pet': {'animal': 1 + (k % 3 == 0)}}), encoding='utf-8')
        for j in range(1 + (k * 3) % 5):
            img = rng.randint(0, 256, (16 + j, 20, 3)).astype(np.uint8)
            kind = (k + j) % 3
            if kind == 0:
                Image.fromarray(img).save(d / f'img_{4 - j}.jpg')
            elif kind == 1:
                Image.fromarray(img).save(d / f'img_{4 - j}.png')
            else:
                np.save(d / f'img_{4 - j}.npy', img)
    if with_cards:          # a folder whose files do not open as images: init_dataset's check() must drop it
        broken = root / 'pet_zz'
        broken.mkdir()
        (broken / 'card.json').write_text(json.dumps({'pet': {'animal': 1}}), encoding='utf-8')
        for j in range(3):
            (broken / f'note_{j}.jpg').write_text('not an image')
    return root
