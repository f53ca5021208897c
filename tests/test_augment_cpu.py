"""The augmentation oracle (oracle/augment_oracle.py) pinned against the library the reference calls: torchvision.transforms
over Pillow (configs/dog_fe/fe_dogs_config.py:17-26), byte for byte, step by step and as the whole Compose under a shared seed."""
import numpy as np
import pytest
import torch

tv = pytest.importorskip('torchvision')
import torchvision.transforms as T                     # noqa: E402
import torchvision.transforms.functional as F          # noqa: E402

from oracle import augment_oracle as A                 # noqa: E402


def _img(seed, size=224):
    g = torch.Generator().manual_seed(seed)
    base = torch.nn.functional.interpolate(torch.rand(1, 3, 9, 9, generator=g), size=size, mode='bicubic')[0]
    x = (base * 0.7 + 0.15 + 0.25 * torch.rand(3, size, size, generator=g)).clamp(0, 1)
    return (x * 255).to(torch.uint8)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_each_step_equals_pillow(seed):
    chw = _img(seed)
    pil = F.to_pil_image(chw)
    hwc = chw.permute(1, 2, 0).numpy()
    assert np.array_equal(A.smooth(hwc), np.asarray(F.adjust_sharpness(pil, 0)))
    assert np.array_equal(A.autocontrast(hwc), np.asarray(F.autocontrast(pil)))
    crop = hwc[3:223, 1:221]
    assert np.array_equal(A.resize_bilinear(crop, 224), np.asarray(F.resize(F.crop(pil, 3, 1, 220, 220), [224, 224])))
    for angle in (-4.99, -0.3, 0.0, 1.7, 5.0):
        assert np.array_equal(A.rotate_nearest(hwc, angle), np.asarray(F.rotate(pil, angle)))


def test_whole_compose_under_a_shared_seed():
    """data_loading.gpu_augment.draw_params draws in torchvision's order: same seed -> same parameters -> same bytes."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / 'pets-face-recognition_b200'))
    from data_loading.gpu_augment import draw_params
    compose = T.Compose([T.ToPILImage(), T.RandomAdjustSharpness(0, 0.5), T.RandomAutocontrast(0.5), T.RandomCrop((220, 220)),
                         T.Resize((224, 224)), T.RandomRotation(5)])
    imgs = [_img(10 + i) for i in range(6)]
    torch.manual_seed(77)
    want = [np.asarray(compose(x)) for x in imgs]
    torch.manual_seed(77)
    params = draw_params(6, 224, 220, 0.5, 0.5, 5.0)
    assert any(p[0] for p in params) and any(p[1] for p in params) and not all(p[0] for p in params)
    for x, p, w in zip(imgs, params, want):
        assert np.array_equal(A.train_augment(x.permute(1, 2, 0).numpy(), *p), w)
