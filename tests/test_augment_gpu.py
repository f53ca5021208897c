"""csrc/augment.cu against the oracle and against torchvision / Pillow itself (byte-exact: integer / byte work), full size."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _img(seed, size=224):
    g = torch.Generator().manual_seed(seed)
    base = torch.nn.functional.interpolate(torch.rand(1, 3, 9, 9, generator=g), size=size, mode='bicubic')[0]
    x = (base * 0.7 + 0.15 + 0.25 * torch.rand(3, size, size, generator=g)).clamp(0, 1)
    return (x * 255).to(torch.uint8)


def test_kernel_equals_oracle_for_every_flag_combination():
    from data_loading.gpu_augment import GpuTrainAugmentation
    from oracle import augment_oracle as A
    imgs = torch.stack([_img(i) for i in range(8)])
    params = [(bool(i & 1), bool(i & 2), (7 * i) % 5, (3 * i) % 5, [-5.0, -2.25, -0.01, 0.0, 0.4, 1.7, 3.3, 4.999][i]) for i in range(8)]
    out = GpuTrainAugmentation()(imgs.cuda(), params=params).cpu()
    for i, p in enumerate(params):
        want = A.train_augment(imgs[i].permute(1, 2, 0).numpy(), *p)
        got = out[i].permute(1, 2, 0).numpy()
        assert np.array_equal(got, want), (i, p, int((got != want).sum()))
    flat = torch.full((2, 3, 224, 224), 117, dtype=torch.uint8)                   # hi == lo: autocontrast leaves the image alone
    o = GpuTrainAugmentation()(flat.cuda(), params=[(True, True, 0, 0, 0.0), (False, True, 4, 4, 0.0)]).cpu()
    assert torch.equal(o, flat)


def test_batch_equals_torchvision_compose_under_a_shared_seed():
    """The same seed on the host -> the same draws as torchvision's transforms make image by image -> the same bytes as the
    reference's Compose (configs/dog_fe/fe_dogs_config.py:17-26 with its own probabilities), 64 images at once."""
    tv = pytest.importorskip('torchvision')
    import torchvision.transforms as T
    from data_loading.gpu_augment import GpuTrainAugmentation
    compose = T.Compose([T.ToPILImage(), T.RandomAdjustSharpness(0, 0.1), T.RandomAutocontrast(0.3), T.RandomCrop((220, 220)),
                         T.Resize((224, 224)), T.RandomRotation(5)])
    imgs = torch.stack([_img(100 + i) for i in range(64)])
    torch.manual_seed(2024)
    want = np.stack([np.asarray(compose(x)) for x in imgs])
    torch.manual_seed(2024)
    got = GpuTrainAugmentation()(imgs.cuda()).cpu().permute(0, 2, 3, 1).numpy()
    assert np.array_equal(got, want), int((got != want).sum())


def test_controller_applies_the_config_hook(monkeypatch):
    from data_loading.gpu_augment import GpuTrainAugmentation
    from engine import Controller
    seen = {}

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    class Loss(torch.nn.Module):
        def forward(self, x, label=None):
            seen['x'] = x
            return {'loss': x.float().mean()}
    cfg = Cfg(model=lambda: torch.nn.Identity(), loss=lambda c, m: Loss(), gpu_train_augmentation=GpuTrainAugmentation())
    ctl = Controller(cfg)
    x = torch.stack([_img(3), _img(4)]).cuda()
    torch.manual_seed(5)
    ctl.training_step({'x': x, 'label': torch.zeros(2, dtype=torch.long).cuda()}, 0)
    torch.manual_seed(5)
    assert torch.equal(seen['x'], GpuTrainAugmentation()(x)) and seen['x'].dtype == torch.uint8
