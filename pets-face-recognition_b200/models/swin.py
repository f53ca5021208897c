"""Swin Transformer feature extractor - B200-native drop-in for the reference's models/swin.py.

Same public surface as /root/reference/models/swin.py:196-241 (SwinTransformer keyword arguments,
swin_t/s/b/l factories) and the same module tree, so ``state_dict()`` has the reference's 168 keys in
the reference's order and checkpoints interchange with it.  The modules below
are parameter containers only: ``SwinTransformer.forward`` hands the whole network to the native plan
(b200/plan.py -> csrc/swin_plan.cu), which runs it as hand-written sm_100a kernels in bf16 with fp32
accumulation / statistics.  There is no eager PyTorch or CPU implementation behind it.
"""
import torch
from torch import nn

from b200.abi import B200Error
from b200.plan import SwinEngine, SwinFunction


def _no_eager(name):
    def forward(self, *_, **__):
        raise B200Error(f'{name} is a parameter container; run the whole SwinTransformer '
                        f'(the B200 path executes the network as one fused native plan)')
    return forward


def _shift_mask(window_size, upper_lower):
    # create_mask of the reference (models/swin.py:49-62) in closed form
    d = window_size // 2
    idx = torch.arange(window_size * window_size)
    part = (idx // window_size if upper_lower else idx % window_size) >= window_size - d
    return torch.zeros(window_size ** 2, window_size ** 2).masked_fill(part[:, None] != part[None, :], float('-inf'))


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn
    forward = _no_eager('Residual')


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn
    forward = _no_eager('PreNorm')


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, dim))
    forward = _no_eager('FeedForward')


class WindowAttention(nn.Module):
    def __init__(self, dim, heads, head_dim, shifted, window_size, relative_pos_embedding):
        super().__init__()
        if not relative_pos_embedding:
            raise B200Error('only relative_pos_embedding=True (the reference default) is built')
        inner_dim = head_dim * heads
        self.heads, self.scale, self.window_size, self.shifted = heads, head_dim ** -0.5, window_size, shifted
        self.relative_pos_embedding = relative_pos_embedding
        if shifted:   # non-trainable Parameters in the reference (models/swin.py:86-89); kept for state_dict parity,
            # the kernel derives the masks analytically (csrc/attention.cu: score_bias)
            self.upper_lower_mask = nn.Parameter(_shift_mask(window_size, True), requires_grad=False)
            self.left_right_mask = nn.Parameter(_shift_mask(window_size, False), requires_grad=False)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.pos_embedding = nn.Parameter(torch.randn(2 * window_size - 1, 2 * window_size - 1))
        self.to_out = nn.Linear(inner_dim, dim)
    forward = _no_eager('WindowAttention')


class SwinBlock(nn.Module):
    def __init__(self, dim, heads, head_dim, mlp_dim, shifted, window_size, relative_pos_embedding):
        super().__init__()
        self.attention_block = Residual(PreNorm(dim, WindowAttention(dim=dim, heads=heads, head_dim=head_dim, shifted=shifted,
                                                                     window_size=window_size,
                                                                     relative_pos_embedding=relative_pos_embedding)))
        self.mlp_block = Residual(PreNorm(dim, FeedForward(dim=dim, hidden_dim=mlp_dim)))
    forward = _no_eager('SwinBlock')


class PatchMerging(nn.Module):
    def __init__(self, in_channels, out_channels, downscaling_factor):
        super().__init__()
        self.downscaling_factor = downscaling_factor
        self.linear = nn.Linear(in_channels * downscaling_factor ** 2, out_channels)
    forward = _no_eager('PatchMerging')


class StageModule(nn.Module):
    def __init__(self, in_channels, hidden_dimension, layers, downscaling_factor, num_heads, head_dim, window_size,
                 relative_pos_embedding):
        super().__init__()
        assert layers % 2 == 0, 'Stage layers need to be divisible by 2 for regular and shifted block.'
        self.patch_partition = PatchMerging(in_channels=in_channels, out_channels=hidden_dimension,
                                            downscaling_factor=downscaling_factor)
        self.layers = nn.ModuleList([])
        for _ in range(layers // 2):
            self.layers.append(nn.ModuleList([
                SwinBlock(dim=hidden_dimension, heads=num_heads, head_dim=head_dim, mlp_dim=hidden_dimension * 4,
                          shifted=False, window_size=window_size, relative_pos_embedding=relative_pos_embedding),
                SwinBlock(dim=hidden_dimension, heads=num_heads, head_dim=head_dim, mlp_dim=hidden_dimension * 4,
                          shifted=True, window_size=window_size, relative_pos_embedding=relative_pos_embedding),
            ]))
    forward = _no_eager('StageModule')


class SwinTransformer(nn.Module):
    def __init__(self, *, hidden_dim, layers, heads, channels=3, num_classes=1000, head_dim=32, window_size=7,
                 downscaling_factors=(4, 2, 2, 2), relative_pos_embedding=True, image_size=224):
        super().__init__()
        dims = [hidden_dim, hidden_dim * 2, hidden_dim * 4, hidden_dim * 8]
        ins = [channels] + dims[:3]
        for i in range(4):
            setattr(self, f'stage{i + 1}', StageModule(in_channels=ins[i], hidden_dimension=dims[i], layers=layers[i],
                                                       downscaling_factor=downscaling_factors[i], num_heads=heads[i],
                                                       head_dim=head_dim, window_size=window_size,
                                                       relative_pos_embedding=relative_pos_embedding))
        self.mlp_head = nn.Sequential(nn.LayerNorm(hidden_dim * 8), nn.Linear(hidden_dim * 8, num_classes))
        self._spec = dict(img=image_size, channels=channels, hidden_dim=hidden_dim, layers=tuple(layers), heads=tuple(heads),
                          downscaling_factors=tuple(downscaling_factors), num_classes=num_classes, head_dim=head_dim,
                          window_size=window_size)
        self._engine = None

    def _trainable(self):
        return [p for n, p in self.named_parameters() if not n.endswith('_mask')]

    @property
    def engine(self) -> SwinEngine:
        if self._engine is None:
            self._engine = SwinEngine(self._spec, self._trainable())
        return self._engine

    def forward(self, img):
        if img.shape[-1] != self._spec['img'] or img.shape[-2] != self._spec['img']:
            # plans are specialised on the input size; rebuild for a new one (e.g. 256x256 body crops need /32 %7 == 0)
            self._spec['img'] = int(img.shape[-1])
            self._engine = None
        eng = self.engine
        params = eng.params
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            eng._ensure_flat(img.device)      # re-point .data BEFORE autograd records the parameters
            return SwinFunction.apply(eng, img, *params)
        return eng.forward(img, False)


def swin_t(hidden_dim=96, layers=(2, 2, 6, 2), heads=(3, 6, 12, 24), **kwargs):
    return SwinTransformer(hidden_dim=hidden_dim, layers=layers, heads=heads, **kwargs)


def swin_s(hidden_dim=96, layers=(2, 2, 18, 2), heads=(3, 6, 12, 24), **kwargs):
    return SwinTransformer(hidden_dim=hidden_dim, layers=layers, heads=heads, **kwargs)


def swin_b(hidden_dim=128, layers=(2, 2, 18, 2), heads=(4, 8, 16, 32), **kwargs):
    return SwinTransformer(hidden_dim=hidden_dim, layers=layers, heads=heads, **kwargs)


def swin_l(hidden_dim=192, layers=(2, 2, 18, 2), heads=(6, 12, 24, 48), **kwargs):
    return SwinTransformer(hidden_dim=hidden_dim, layers=layers, heads=heads, **kwargs)
