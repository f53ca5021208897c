"""Deterministic synthetic weights, images and embeddings.

There is no network for datasets or checkpoints, so every parity test and bench runs on
seeded synthetic data.  Weights are generated per state-dict key from a generator seeded with
(seed, crc32(key)): any process (the build container that ran the reference to produce
tests/golden/, the GPU box that checks against it) regenerates identical tensors without
shipping 110 MB of parameters.  All generation happens on the CPU generator.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device='cpu')
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int = 123) -> torch.Tensor:
    """Scale follows PyTorch's default initialisers of the reference modules (Linear:
    U(+-1/sqrt(fan_in)); pos_embedding: N(0,1), models/swin.py:95; ArcFace weight: Xavier-uniform,
    losses/large_margin.py:61), except LayerNorm affine, which is perturbed away from (1, 0) so
    that its gradients are exercised."""
    g = _gen(seed, key)
    if key.endswith('_mask'):
        raise ValueError('masks are structural, not random')
    if key.endswith('pos_embedding'):
        return torch.randn(shape, generator=g)
    if 'norm.' in key or key.startswith('mlp_head.0.') or '.mlp_head.0.' in key:
        base = 1.0 if key.endswith('weight') else 0.0
        return base + 0.1 * torch.randn(shape, generator=g)
    if key.endswith('add_margin.weight'):
        bound = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if key.endswith('.bias'):
        # bias bound uses the fan_in of the matching weight; recover it from the sibling's key
        # is not possible here, so callers pass fan_in through synth_state_dict.
        raise ValueError('use synth_state_dict for biases')
    bound = 1.0 / math.sqrt(shape[-1])
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 123, window_size: int = 7) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in shapes.items():
        if key.endswith('upper_lower_mask') or key.endswith('left_right_mask'):
            sd[key] = _mask(window_size, key.endswith('upper_lower_mask'))
        elif key.endswith('.bias') and 'norm.' not in key and 'mlp_head.0.' not in key:
            fan_in = shapes[key[:-len('bias')] + 'weight'][-1]
            g = _gen(seed, key)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:
            sd[key] = synth_tensor(key, shape, seed)
    return sd


def _mask(ws: int, upper_lower: bool) -> torch.Tensor:
    d = ws // 2
    idx = torch.arange(ws * ws)
    part = (idx // ws if upper_lower else idx % ws) >= ws - d
    return torch.zeros(ws * ws, ws * ws).masked_fill(part[:, None] != part[None, :], float('-inf'))


def synth_images(batch: int, seed: int = 123, hw: int = 224, channels: int = 3) -> torch.Tensor:
    """torch.rand in [0,1): the ToTensor range; the reference applies no mean/std normalisation
    (configs/dog_fe/fe_dogs_config.py:17-32)."""
    g = _gen(seed, f'images/{batch}/{hw}')
    return torch.rand(batch, channels, hw, hw, generator=g)


def synth_labels(batch: int, num_class: int, seed: int = 123) -> torch.Tensor:
    g = _gen(seed, f'labels/{batch}/{num_class}')
    return torch.randint(0, num_class, (batch,), generator=g, dtype=torch.int64)


def synth_embeddings(n_identities: int, per_identity: int, dim: int = 512, sigma: float = 1.0,
                     seed: int = 123):
    """Identity centres c_i ~ N(0, I)/|.|, sample = normalize(c_i + sigma * eps), eps ~ N(0, I/dim).
    Returns (emb fp32 [n_identities*per_identity, dim], classes int64), interleaved so that row
    r has class r % n_identities (same-class rows are far apart in memory)."""
    g = _gen(seed, f'emb/{n_identities}/{per_identity}/{dim}')
    centres = torch.randn(n_identities, dim, generator=g)
    centres = centres / centres.norm(dim=1, keepdim=True)
    eps = torch.randn(per_identity, n_identities, dim, generator=g) / math.sqrt(dim)
    emb = centres[None] + sigma * eps
    emb = emb / emb.norm(dim=2, keepdim=True)
    classes = torch.arange(n_identities).repeat(per_identity)
    return emb.reshape(-1, dim).contiguous(), classes
