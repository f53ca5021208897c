"""CPU restatement (numpy) of the reference's training augmentation, configs/dog_fe/fe_dogs_config.py:17-26:
ToPILImage -> RandomAdjustSharpness(0, 0.1) -> RandomAutocontrast(0.3) -> RandomCrop(220) -> Resize(224) -> RandomRotation(5).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The arithmetic lives in a third-party dependency of the reference, not under
/root/reference: Pillow (installed 12.2.0; reached through torchvision.transforms, installed 0.26.0 - the reference pins
`torchvision` without a version, requirements.txt).  Each function restates the published algorithm of the Pillow routine it
names and is pinned against Pillow itself by tests/test_augment_cpu.py (Pillow is importable on both boxes).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def smooth(img: np.ndarray) -> np.ndarray:
    """ImageEnhance.Sharpness(img).enhance(0) == img.filter(ImageFilter.SMOOTH): libImaging/Filter.c ImagingFilter3x3 with the
    kernel (1 1 1 / 1 5 1 / 1 1 1) / 13 in float32, offset 0.5, truncation; the one-pixel border is copied.  img: uint8 [H, W, 3]."""
    k1, k5 = np.float32(1.0) / np.float32(13.0), np.float32(5.0) / np.float32(13.0)
    f = img.astype(np.float32)
    out = img.copy()
    H, W = img.shape[:2]

    def row(r, a, b, c):          # (in[x-1] * a + in[x] * b) + in[x+1] * c, float32 throughout
        return (f[r, 0:W - 2] * a + f[r, 1:W - 1] * b) + f[r, 2:W] * c
    for y in range(1, H - 1):
        ss = np.full((W - 2, 3), np.float32(0.5), dtype=np.float32)
        ss = ss + row(y + 1, k1, k1, k1)
        ss = ss + row(y, k1, k5, k1)
        ss = ss + row(y - 1, k1, k1, k1)
        out[y, 1:W - 1] = np.where(ss <= 0, 0, np.where(ss >= 255, 255, ss.astype(np.int32))).astype(np.uint8)
    return out


def autocontrast(img: np.ndarray) -> np.ndarray:
    """ImageOps.autocontrast(img) with cutoff 0: per channel lo / hi = first / last populated histogram bin, lut[i] =
    clamp(int(i * 255.0 / (hi - lo) - lo * 255.0 / (hi - lo))) in Python floats (int() truncates toward zero)."""
    out = img.copy()
    for c in range(img.shape[2]):
        lo, hi = int(img[..., c].min()), int(img[..., c].max())
        if hi <= lo:
            continue
        scale = 255.0 / (hi - lo)
        offset = -lo * scale
        lut = np.array([min(255, max(0, int(i * scale + offset))) for i in range(256)], dtype=np.uint8)
        out[..., c] = lut[img[..., c]]
    return out


def bilinear_coefficients(in_size: int, out_size: int):
    """libImaging/Resample.c: precompute_coeffs (bilinear filter, support 1) + normalize_coeffs_8bpc."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ss = 1.0 / filterscale
    bounds, coefs = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = []
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * ss)
            k.append(1.0 - t if t < 1.0 else 0.0)
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        coefs.append([int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS)) for w in k])
        bounds.append(xmin)
    return bounds, coefs


def resize_bilinear(img: np.ndarray, out_size: int) -> np.ndarray:
    """Image.resize((S, S), BILINEAR) on an 8-bit RGB image: horizontal pass to uint8, then vertical pass to uint8
    (ImagingResampleHorizontal_8bpc / Vertical_8bpc: ss = 1 << 21; ss += pixel * k; out = clip8(ss >> 22))."""
    H, W = img.shape[:2]
    bx, cx = bilinear_coefficients(W, out_size)
    by, cy = bilinear_coefficients(H, out_size)
    src = img.astype(np.int64)
    tmp = np.empty((H, out_size, 3), dtype=np.int64)
    for x in range(out_size):
        acc = np.full((H, 3), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for i, k in enumerate(cx[x]):
            acc += src[:, bx[x] + i] * k
        tmp[:, x] = np.clip(acc >> PRECISION_BITS, 0, 255)
    out = np.empty((out_size, out_size, 3), dtype=np.uint8)
    for y in range(out_size):
        acc = np.full((out_size, 3), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j, k in enumerate(cy[y]):
            acc += tmp[by[y] + j] * k
        out[y] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def rotate_nearest(img: np.ndarray, angle: float) -> np.ndarray:
    """Image.rotate(angle) (NEAREST, expand=False, centre = image centre, fill 0): Image.rotate's matrix (cos / sin rounded to 15
    places) walked by libImaging/Geometry.c: affine_fixed in 16.16 fixed point."""
    angle = angle % 360.0
    if angle == 0:
        return img.copy()
    H, W = img.shape[:2]
    rad = -math.radians(angle)
    m = [round(math.cos(rad), 15), round(math.sin(rad), 15), 0.0, round(-math.sin(rad), 15), round(math.cos(rad), 15), 0.0]
    cx, cy = W / 2, H / 2
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy

    def fix(v):
        v = v * 65536.0 + 0.5
        return int(v) if v >= 0.0 else int(math.floor(v))
    a0, a1, a3, a4 = fix(m[0]), fix(m[1]), fix(m[3]), fix(m[4])
    a2, a5 = fix(m[2] + m[0] * 0.5 + m[1] * 0.5), fix(m[5] + m[3] * 0.5 + m[4] * 0.5)
    ys, xs = np.mgrid[0:H, 0:W]
    xin = (a2 + a1 * ys + a0 * xs) >> 16
    yin = (a5 + a4 * ys + a3 * xs) >> 16
    ok = (xin >= 0) & (xin < W) & (yin >= 0) & (yin < H)
    out = np.zeros_like(img)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def train_augment(img: np.ndarray, sharpen: bool, ac: bool, top: int, left: int, angle: float, crop: int = 220, size: int = 224) -> np.ndarray:
    """uint8 [H, W, 3] -> uint8 [size, size, 3] for one set of draws."""
    x = smooth(img) if sharpen else img
    x = autocontrast(x) if ac else x
    x = x[top:top + crop, left:left + crop]
    x = resize_bilinear(x, size)
    return rotate_nearest(x, angle)
