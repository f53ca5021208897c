"""The reference's training augmentation (configs/dog_fe/fe_dogs_config.py:17-26) on the GPU, for a whole uint8 batch at once
(csrc/augment.cu).  SURVEY.md 8f-3: at > 12 k images/s per GPU four PIL workers cannot feed the step.

    aug = GpuTrainAugmentation()                       # crop 220 -> 224, p_sharp 0.1, p_autocontrast 0.3, +-5 degrees
    x = aug(batch_u8_on_device)                        # uint8 [B, 3, 224, 224] -> uint8 [B, 3, 224, 224]

The random draws are made on the host with torch's global generator in the ORDER torchvision's transforms make them for one
image after the other (sharpness flag, autocontrast flag, crop top, crop left, angle), so under the same seed the parameters -
and with them every output byte - equal those of the reference's Compose applied image by image.  `params=` takes explicit
draws (tests).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from b200 import abi
from b200.abi import check, lib, ptr, stream_ptr

_PRECISION_BITS = 32 - 8 - 2          # libImaging/Resample.c


def resize_coefficients(in_size: int, out_size: int) -> List[Tuple[int, int, int, int]]:
    """PIL's bilinear coefficients for in_size -> out_size (libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc):
    per output index the first source index and up to three fixed-point weights (zero padded)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    if int(math.ceil(support)) * 2 + 1 > 3:
        raise abi.B200Error('resize_coefficients: down-scaling by more than 1 needs more than three taps (not built)')
    ss = 1.0 / filterscale
    out = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = []
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * ss)
            k.append(1.0 - t if t < 1.0 else 0.0)
        ww = sum(k)                                    # C accumulates in the same left-to-right order
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        kk = [int(-0.5 + w * (1 << _PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << _PRECISION_BITS)) for w in k]
        kk += [0] * (3 - len(kk))
        if len(kk) > 3:
            raise abi.B200Error('resize_coefficients: more than three taps')
        out.append((xmin, kk[0], kk[1], kk[2]))
    return out


def rotation_fixed_point(angle: float, size: int) -> Tuple[int, ...]:
    """PIL Image.rotate(angle, NEAREST, expand=False) about the image centre -> the six 16.16 fixed-point numbers of
    libImaging/Geometry.c: affine_fixed, which walks xin = a2 + a0 x + a1 y, yin = a5 + a3 x + a4 y."""
    angle = angle % 360.0
    rad = -math.radians(angle)
    m = [round(math.cos(rad), 15), round(math.sin(rad), 15), 0.0, round(-math.sin(rad), 15), round(math.cos(rad), 15), 0.0]
    cx = cy = size / 2
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy

    def fix(v):
        v = v * 65536.0 + 0.5
        return int(v) if v >= 0.0 else int(math.floor(v))
    a2 = m[2] + m[0] * 0.5 + m[1] * 0.5
    a5 = m[5] + m[3] * 0.5 + m[4] * 0.5
    return fix(m[0]), fix(m[1]), fix(a2), fix(m[3]), fix(m[4]), fix(a5)


def draw_params(n: int, size: int = 224, crop: int = 220, p_sharp: float = 0.1, p_autocontrast: float = 0.3, degrees: float = 5.0):
    """Per image, in torchvision's order: RandomAdjustSharpness (torch.rand(1) < p), RandomAutocontrast (torch.rand(1) < p),
    RandomCrop.get_params (two torch.randint draws: top, left), RandomRotation.get_params (uniform_(-d, d))."""
    out = []
    for _ in range(n):
        sharp = bool(torch.rand(1).item() < p_sharp)
        ac = bool(torch.rand(1).item() < p_autocontrast)
        if size == crop:
            top = left = 0
        else:
            top = int(torch.randint(0, size - crop + 1, size=(1,)).item())
            left = int(torch.randint(0, size - crop + 1, size=(1,)).item())
        angle = float(torch.empty(1).uniform_(-degrees, degrees).item())
        out.append((sharp, ac, top, left, angle))
    return out


class GpuTrainAugmentation:
    def __init__(self, size: int = 224, crop: int = 220, p_sharp: float = 0.1, p_autocontrast: float = 0.3, degrees: float = 5.0):
        self.size, self.crop, self.p_sharp, self.p_ac, self.degrees = size, crop, p_sharp, p_autocontrast, degrees
        self._coef_host = torch.tensor(resize_coefficients(crop, size), dtype=torch.int32)
        self._coef = {}

    def __call__(self, x: torch.Tensor, params: Optional[Sequence[Tuple[bool, bool, int, int, float]]] = None) -> torch.Tensor:
        if not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise abi.B200Error('GpuTrainAugmentation takes a CUDA uint8 [B, 3, S, S] batch (no CPU fallback)')
        B, _, H, W = x.shape
        if H != self.size:
            raise abi.B200Error(f'GpuTrainAugmentation(size={self.size}) got {H} x {W} images')
        params = list(params) if params is not None else draw_params(B, self.size, self.crop, self.p_sharp, self.p_ac, self.degrees)
        arr = (abi.AugParams * B)()
        for i, (sharp, ac, top, left, angle) in enumerate(params):
            arr[i].sharpen, arr[i].autocontrast, arr[i].crop_y, arr[i].crop_x = int(sharp), int(ac), int(top), int(left)
            if angle % 360.0 == 0:
                rot = (65536, 0, 32768, 0, 65536, 32768)      # PIL's fast path returns a copy: the identity walk
            else:
                rot = rotation_fixed_point(angle, self.size)
            for j in range(6):
                arr[i].rot[j] = rot[j]
        prm = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory().to(x.device, non_blocking=True)
        coef = self._coef.get(x.device)
        if coef is None:
            coef = self._coef[x.device] = self._coef_host.to(x.device).contiguous()
        x = x.contiguous()
        out = torch.empty_like(x)
        scratch = torch.empty(B * 6, dtype=torch.uint8, device=x.device)
        check(lib().b200_augment_train(ptr(x), ptr(out), ptr(prm), ptr(coef), B, H, W, self.crop, self.size, ptr(scratch), stream_ptr()),
              'augment_train')
        return out
