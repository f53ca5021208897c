"""Functional CPU restatement of the reference Swin Transformer forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows /root/reference/models/swin.py
(berniwal variant) but is written as one functional pass over a plain state dict,
with the cyclic shift / window partition expressed as index arithmetic instead of
roll + einops rearranges, so that it shares no code shape with the reference and
doubles as the specification of the addressing the CUDA kernels use.

All arithmetic is done in the dtype of ``x`` (fp32 or fp64); autograd through this
function is the gradient oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class SwinSpec:
    """Constructor arguments of models/swin.py:197-198 (+ swin_t defaults :228)."""
    hidden_dim: int = 96
    layers: Tuple[int, ...] = (2, 2, 6, 2)
    heads: Tuple[int, ...] = (3, 6, 12, 24)
    channels: int = 3
    num_classes: int = 512
    head_dim: int = 32
    window_size: int = 7
    downscaling_factors: Tuple[int, ...] = (4, 2, 2, 2)

    def stage_dims(self):
        return [self.hidden_dim * m for m in (1, 2, 4, 8)]


def param_shapes(spec: SwinSpec) -> Dict[str, Tuple[int, ...]]:
    """State-dict template of SwinTransformer, in the
    registration order of models/swin.py:201-217."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    dims = spec.stage_dims()
    c_in = spec.channels
    ws = spec.window_size
    for s in range(4):
        c = dims[s]
        df = spec.downscaling_factors[s]
        inner = spec.head_dim * spec.heads[s]
        p = f'stage{s + 1}.'
        shapes[p + 'patch_partition.linear.weight'] = (c, c_in * df * df)
        shapes[p + 'patch_partition.linear.bias'] = (c,)
        for pair in range(spec.layers[s] // 2):
            for blk in range(2):
                q = f'{p}layers.{pair}.{blk}.'
                a = q + 'attention_block.fn.'
                shapes[a + 'norm.weight'] = (c,)
                shapes[a + 'norm.bias'] = (c,)
                if blk == 1:  # shifted block registers its masks first (models/swin.py:82-89)
                    shapes[a + 'fn.upper_lower_mask'] = (ws * ws, ws * ws)
                    shapes[a + 'fn.left_right_mask'] = (ws * ws, ws * ws)
                shapes[a + 'fn.pos_embedding'] = (2 * ws - 1, 2 * ws - 1)
                shapes[a + 'fn.to_qkv.weight'] = (3 * inner, c)
                shapes[a + 'fn.to_out.weight'] = (c, inner)
                shapes[a + 'fn.to_out.bias'] = (c,)
                m = q + 'mlp_block.fn.'
                shapes[m + 'norm.weight'] = (c,)
                shapes[m + 'norm.bias'] = (c,)
                shapes[m + 'fn.net.0.weight'] = (4 * c, c)
                shapes[m + 'fn.net.0.bias'] = (4 * c,)
                shapes[m + 'fn.net.2.weight'] = (c, 4 * c)
                shapes[m + 'fn.net.2.bias'] = (c,)
        c_in = c
    shapes['mlp_head.0.weight'] = (dims[3],)
    shapes['mlp_head.0.bias'] = (dims[3],)
    shapes['mlp_head.1.weight'] = (spec.num_classes, dims[3])
    shapes['mlp_head.1.bias'] = (spec.num_classes,)
    return shapes


def shift_mask(ws: int, upper_lower: bool, left_right: bool, dtype=torch.float32) -> torch.Tensor:
    """create_mask, models/swin.py:49-62, in closed form: with displacement d = ws // 2,
    token (r, c) of a window belongs to the 'wrapped' part iff r >= ws - d (upper/lower)
    resp. c >= ws - d (left/right); pairs straddling the boundary get -inf."""
    d = ws // 2
    idx = torch.arange(ws * ws)
    r, c = idx // ws, idx % ws
    m = torch.zeros(ws * ws, ws * ws, dtype=dtype)
    if upper_lower:
        part = r >= ws - d
        m = m.masked_fill(part[:, None] != part[None, :], float('-inf'))
    if left_right:
        part = c >= ws - d
        m = m.masked_fill(part[:, None] != part[None, :], float('-inf'))
    return m


def relative_bias(pos: torch.Tensor, ws: int) -> torch.Tensor:
    """models/swin.py:65-68,93-95,117-118: bias[i, j] = pos[r_j - r_i + ws-1, c_j - c_i + ws-1],
    one (2ws-1)^2 table shared by every head and window of the block."""
    idx = torch.arange(ws * ws)
    r, c = idx // ws, idx % ws
    dr = r[None, :] - r[:, None] + ws - 1
    dc = c[None, :] - c[:, None] + ws - 1
    return pos[dr, dc]


def _layer_norm(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)   # nn.LayerNorm default eps, models/swin.py:29


def attention_core(qkv, pos, heads, head_dim, ws, shifted):
    """The part of WindowAttention.forward between to_qkv and to_out, models/swin.py:102-130,133-134.
    qkv: (B, H, W, 3*heads*head_dim) on the UNSHIFTED pixel grid; returns (B, H, W, heads*head_dim) on
    the same grid.  roll(-d) means shifted[y] = x[(y + d) % H] (:83, :102-103); the inverse roll on the
    way out puts every token's result back at its source pixel, so the shift is pure addressing."""
    B, H, W, _ = qkv.shape
    nh, nw = H // ws, W // ws
    off = ws // 2 if shifted else 0
    ys = (torch.arange(H) + off) % H
    xs = (torch.arange(W) + off) % W
    g = qkv[:, ys][:, :, xs]                                        # gather the shifted grid
    g = g.reshape(B, nh, ws, nw, ws, 3, heads, head_dim)            # [q|k|v] chunk, then (h d) :107-113
    g = g.permute(5, 0, 6, 1, 3, 2, 4, 7).reshape(3, B, heads, nh * nw, ws * ws, head_dim)
    q, k, v = g[0], g[1], g[2]
    dots = (q @ k.transpose(-1, -2)) * (head_dim ** -0.5)           # :115, :77
    dots = dots + relative_bias(pos, ws).to(dots.dtype)             # :117-118
    if shifted:                                                     # :122-124
        ul = torch.isinf(shift_mask(ws, True, False))
        lr = torch.isinf(shift_mask(ws, False, True))
        widx = torch.arange(nh * nw)
        last_row = (widx // nw) == nh - 1                           # dots[:, :, -nw_w:]
        last_col = (widx % nw) == nw - 1                            # dots[:, :, nw_w-1::nw_w]
        banned = (last_row[:, None, None] & ul[None]) | (last_col[:, None, None] & lr[None])
        dots = dots.masked_fill(banned[None, None].to(dots.device), float('-inf'))
    attn = dots.softmax(dim=-1)                                     # :126
    out = attn @ v                                                  # :128
    out = out.reshape(B, heads, nh, nw, ws, ws, head_dim).permute(0, 2, 4, 3, 5, 1, 6)
    out = out.reshape(B, H, W, heads * head_dim)                    # :129-130, still on the shifted grid
    res = torch.empty_like(out)
    res[:, ys[:, None], xs[None, :]] = out                          # scatter back == cyclic_back_shift
    return res


def window_attention(x, sd, pfx, heads, head_dim, ws, shifted):
    """WindowAttention.forward, models/swin.py:101-135, on NHWC x (B, H, W, C)."""
    qkv = x @ sd[pfx + 'to_qkv.weight'].t()                       # :107, no bias (:91)
    res = attention_core(qkv, sd[pfx + 'pos_embedding'], heads, head_dim, ws, shifted)
    return res @ sd[pfx + 'to_out.weight'].t() + sd[pfx + 'to_out.bias']   # :131 (linear commutes with the roll)


def swin_block(x, sd, pfx, heads, head_dim, ws, shifted):
    """SwinBlock.forward, models/swin.py:149-152 with Residual :22-23 and PreNorm :32-33."""
    a = pfx + 'attention_block.fn.'
    x = x + window_attention(_layer_norm(x, sd[a + 'norm.weight'], sd[a + 'norm.bias']),
                             sd, a + 'fn.', heads, head_dim, ws, shifted)
    m = pfx + 'mlp_block.fn.'
    h = _layer_norm(x, sd[m + 'norm.weight'], sd[m + 'norm.bias'])
    h = F.gelu(h @ sd[m + 'fn.net.0.weight'].t() + sd[m + 'fn.net.0.bias'])   # erf GELU, :41
    return x + (h @ sd[m + 'fn.net.2.weight'].t() + sd[m + 'fn.net.2.bias'])


def patch_merge(x_nhwc, w, b, df):
    """PatchMerging.forward, models/swin.py:162-167.  nn.Unfold's feature order inside a
    df x df patch is channel-major: index = c*df*df + kh*df + kw."""
    B, H, W, C = x_nhwc.shape
    p = x_nhwc.reshape(B, H // df, df, W // df, df, C).permute(0, 1, 3, 5, 2, 4)
    p = p.reshape(B, H // df, W // df, C * df * df)
    return p @ w.t() + b


def swin_forward(sd: Dict[str, torch.Tensor], img: torch.Tensor, spec: SwinSpec = SwinSpec(),
                 prefix: str = '') -> torch.Tensor:
    """SwinTransformer.forward, models/swin.py:219-225: (B, 3, H, W) in [0,1] -> (B, num_classes)."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = img.permute(0, 2, 3, 1)                                     # NHWC throughout
    for s in range(4):
        p = f'stage{s + 1}.'
        x = patch_merge(x, sd[p + 'patch_partition.linear.weight'], sd[p + 'patch_partition.linear.bias'],
                        spec.downscaling_factors[s])
        for pair in range(spec.layers[s] // 2):                     # StageModule.forward :188-193
            for blk in range(2):
                x = swin_block(x, sd, f'{p}layers.{pair}.{blk}.', spec.heads[s], spec.head_dim,
                               spec.window_size, shifted=(blk == 1))
    x = x.mean(dim=(1, 2))                                          # :224
    x = _layer_norm(x, sd['mlp_head.0.weight'], sd['mlp_head.0.bias'])
    return x @ sd['mlp_head.1.weight'].t() + sd['mlp_head.1.bias']  # :214-217


def swin_flops_per_image(spec: SwinSpec = SwinSpec(), hw: int = 224) -> float:
    """2*MAC of every Linear and both attention matmuls (8.98 GFLOP for Swin-T)."""
    dims = spec.stage_dims()
    c_in, res, total = spec.channels, hw, 0.0
    for s in range(4):
        df = spec.downscaling_factors[s]
        res //= df
        t, c = res * res, dims[s]
        total += 2.0 * t * c_in * df * df * c
        inner = spec.head_dim * spec.heads[s]
        per_block = 2.0 * t * c * 3 * inner + 2.0 * t * inner * c + 2 * 2.0 * t * c * 4 * c
        per_block += 2 * 2.0 * t * (spec.window_size ** 2) * inner
        total += spec.layers[s] * per_block
        c_in = c
    total += 2.0 * dims[3] * spec.num_classes
    return total
