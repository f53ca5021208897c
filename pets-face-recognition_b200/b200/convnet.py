"""ResNet-50 feature extractor on the B200 kernels: the host side of csrc/convnet.cu + the tcgen05 GEMMs.

The reference's FE configs build `torchvision.models.resnet50(pretrained=True)` and replace `fc` by `Linear(2048, 512)`
(configs/dog_fe/fe_dogs_config.py:96-109).  `ConvNetEngine` runs that network - training forward / backward with
batch-statistics BatchNorm, or the eval forward with the running statistics - on bf16 NHWC activations stored as rows of a
padded grid (see csrc/convnet.cu): 1x1 convolutions are plain GEMMs over those rows, 3x3 convolutions are implicit GEMMs
(b200_gemm_taps: nine row-shifted contributions, no im2col matrix), stride-2 convolutions are computed on the fine grid and
sampled, every weight gradient is the MN-major split-K GEMM the Swin path uses (the 3x3 ones as nine shifted launches).

There is no CPU or eager fallback: a missing library or a non-CUDA input raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import abi, ops
from . import plan as _plan
from .abi import B200Error, check, lib, ptr, stream_ptr

bf16 = torch.bfloat16


def _shifts(W: int) -> List[int]:
    """row shift of tap (r, s) of a 3x3 window on a padded grid of width W + 2"""
    return [(r - 1) * (W + 2) + (s - 1) for r in range(3) for s in range(3)]


class _WeightCache:
    """bf16 GEMM layouts of the fp32 master weights, rebuilt when a parameter changes (optimizer step / load_state_dict)."""

    def __init__(self):
        self.store: Dict[int, Tuple[tuple, tuple]] = {}

    def get(self, p: torch.Tensor, kind: str) -> tuple:
        key = (_plan._weight_epoch, p._version, p.data_ptr(), kind)      # the fused optimizer writes through raw pointers: epoch
        hit = self.store.get(id(p))
        if hit is not None and hit[0] == key:
            return hit[1]
        w = p.detach()
        if kind == 'conv':                          # [Co, Ci, kh, kw] -> forward [Co, (r, s, ci)], data gradient [Ci, (r, s, co)]
            fwd = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(bf16).contiguous()
            dgrad = w.permute(1, 2, 3, 0).reshape(w.shape[1], -1).to(bf16).contiguous()
            val = (fwd, dgrad)
        elif kind == 'stem':                        # [64, 3, 7, 7] -> [64, 160]: (r, s, c) order, zero-padded from 147
            fwd = torch.zeros(w.shape[0], 160, device=w.device, dtype=bf16)
            fwd[:, :147] = w.permute(0, 2, 3, 1).reshape(w.shape[0], 147).to(bf16)
            val = (fwd, None)
        else:                                       # linear [out, in]
            val = (w.to(bf16).contiguous(), w.t().to(bf16).contiguous())
        self.store[id(p)] = (key, val)
        return val


class ConvNetEngine:
    def __init__(self, model: torch.nn.Module):
        self.model = model
        self.wcache = _WeightCache()
        self.saved = None
        self._token = 0
        self._shift_arrays: Dict[Tuple[int, int], C.Array] = {}
        self._frozen = False
        self._counters: List[torch.Tensor] = []
        self._eval_consts: Dict[int, tuple] = {}
        self._folded_cache: Dict[int, tuple] = {}

    # ------------------------------------------------------------------ small wrappers over the C ABI
    def _shift_array(self, W: int, sign: int):
        arr = self._shift_arrays.get((W, sign))
        if arr is None:
            arr = self._shift_arrays[(W, sign)] = (C.c_int * 9)(*[sign * s for s in _shifts(W)])
        return arr

    def _taps(self, a, w, W: int, sign: int, aux=None):
        arr = self._shift_array(W, sign)
        M, Cin = a.shape
        N = w.shape[0]
        out = torch.empty(M, N, device=a.device, dtype=bf16)
        check(lib().b200_gemm_taps(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, Cin, 9, arr, abi.EPI_RESID if aux is not None else abi.EPI_STORE,
                                   ptr(out), N, ptr(aux), aux.stride(0) if aux is not None else 0, stream_ptr()), 'gemm_taps')
        return out

    def _folded(self, conv, bn):
        """eval mode: (bf16 [Co, (r, s, ci)] weights with the BatchNorm scale folded in, fp32 shift), rebuilt when a tensor changes"""
        if bn.running_mean is None or bn.running_var is None:
            raise B200Error('eval-mode BatchNorm without running statistics (track_running_stats=False) is not built')
        key = (_plan._weight_epoch, conv.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
               conv.weight.data_ptr())
        hit = self._folded_cache.get(id(conv))
        if hit is not None and hit[0] == key:
            return hit[1]
        scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.float() + bn.eps)
        shift = (bn.bias.detach().float() - bn.running_mean.float() * scale).contiguous()
        w = (conv.weight.detach().float() * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(conv.weight.shape[0], -1).to(bf16).contiguous()
        self._folded_cache[id(conv)] = (key, (w, shift))
        return w, shift

    def _conv_bn(self, x, conv, bn, H: int, W: int, relu: bool, aux=None):
        """eval mode: convolution (1x1 or 3x3 on the (H, W) grid) + BatchNorm (+ residual) (+ ReLU) in one GEMM launch"""
        w, shift = self._folded(conv, bn)
        M, Cin = x.shape
        N = w.shape[0]
        taps = 9 if conv.kernel_size == (3, 3) else 0
        out = torch.empty(M, N, device=x.device, dtype=bf16)
        check(lib().b200_gemm_conv_bn(ptr(x), x.stride(0), ptr(w), w.stride(0), M, N, Cin, taps, self._shift_array(W, 1) if taps else None, ptr(shift),
                                      int(relu), H, W, ptr(aux), aux.stride(0) if aux is not None else 0, ptr(out), N, stream_ptr()), 'gemm_conv_bn')
        return out

    def _block_forward_eval(self, blk, x, B: int, H: int, W: int):
        """inference: four launches per Bottleneck (five with a down-sampling branch), no BatchNorm pass of its own"""
        stride = blk.conv2.stride[0]
        if blk.conv1.stride[0] != 1 or blk.conv2.kernel_size != (3, 3) or blk.conv2.groups != 1 or blk.conv2.dilation[0] != 1 or stride not in (1, 2):
            raise B200Error('ConvNetEngine supports the torchvision v1.5 Bottleneck (stride on the 3x3 convolution, no groups / dilation)')
        y1 = self._conv_bn(x, blk.conv1, blk.bn1, H, W, True)
        y2 = self._conv_bn(y1, blk.conv2, blk.bn2, H, W, True)
        Ho, Wo = H // stride, W // stride
        if stride == 2:
            y2 = self._sample(y2, B, H, W, True)
        if blk.downsample is not None:
            xs = self._sample(x, B, H, W, True) if stride == 2 else x
            idn = self._conv_bn(xs, blk.downsample[0], blk.downsample[1], Ho, Wo, False)
        else:
            idn = x
        return self._conv_bn(y2, blk.conv3, blk.bn3, Ho, Wo, True, aux=idn), Ho, Wo

    def _scratch(self, rows: int, Cc: int, dev):
        return torch.empty(lib().b200_bn_stats_blocks(rows) * 2 * Cc, device=dev, dtype=torch.float32)

    def _bn_consts(self, bn, x, H: int, W: int, count: int, training: bool):
        """[4, C] = scale, shift, mean, rstd"""
        rows, Cc = x.shape
        out = torch.empty(4, Cc, device=x.device, dtype=torch.float32)
        if training:
            track = bn.track_running_stats and bn.running_mean is not None
            if bn.momentum is None:           # nn.BatchNorm2d: cumulative moving average, factor 1 / (batches seen, this one included)
                seen = int(bn.num_batches_tracked) if (track and bn.num_batches_tracked is not None) else 0
                momentum = 1.0 / (seen + 1)
            else:
                momentum = float(bn.momentum)
            check(lib().b200_bn_stats(ptr(x), rows, Cc, H, W, float(count), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean) if track else None,
                                      ptr(bn.running_var) if track else None, momentum, float(bn.eps), ptr(out), ptr(self._scratch(rows, Cc, x.device)),
                                      stream_ptr()), 'bn_stats')
            self._eval_consts.pop(id(bn), None)                     # the kernel just rewrote the running statistics in place
            if track and bn.num_batches_tracked is not None:
                self._counters.append(bn.num_batches_tracked)       # bumped together at the end of the forward (one launch)
        else:
            if bn.running_mean is None or bn.running_var is None:
                raise B200Error('eval-mode BatchNorm without running statistics (track_running_stats=False) is not built')
            # eval mode: constants of the running statistics, rebuilt only when one of the four tensors changes
            key = (_plan._weight_epoch, bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, bn.weight.data_ptr())
            hit = self._eval_consts.get(id(bn))
            if hit is not None and hit[0] == key:
                return hit[1]
            rstd = torch.rsqrt(bn.running_var.float() + bn.eps)
            out[0] = bn.weight.detach() * rstd
            out[1] = bn.bias.detach() - bn.running_mean * out[0]
            out[2] = bn.running_mean
            out[3] = rstd
            self._eval_consts[id(bn)] = (key, out)
        return out

    def _bn_apply(self, x, consts, H: int, W: int, relu: bool, residual=None):
        rows, Cc = x.shape
        y = torch.empty_like(x)
        check(lib().b200_bn_apply(ptr(x), ptr(consts[0]), ptr(consts[1]), ptr(residual), int(relu), rows, Cc, H, W, ptr(y), stream_ptr()), 'bn_apply')
        return y

    def _bn_backward(self, dy, y, x, consts, gamma, H: int, W: int, count: int, want_dz: bool = False, relu_from_x: bool = False):
        rows, Cc = x.shape
        dx = torch.empty_like(x)
        dz = torch.empty_like(x) if want_dz else None
        sums = torch.empty(2, Cc, device=x.device, dtype=torch.float32)
        check(lib().b200_bn_backward(ptr(dy), ptr(y), int(relu_from_x), ptr(x), ptr(consts), ptr(gamma), rows, Cc, H, W, 0.0 if self._frozen else float(count), ptr(dx), ptr(dz), ptr(sums),
                                     ptr(self._scratch(rows, Cc, x.device)), stream_ptr()), 'bn_backward')
        return dx, dz, sums

    def _sample(self, x, B: int, H: int, W: int, down: bool):
        """down: (H, W) grid -> (H/2, W/2) grid; up: the adjoint, x on the (H/2, W/2) grid -> (H, W) grid"""
        Cc = x.shape[1]
        rows = B * ((H // 2 + 2) * (W // 2 + 2) if down else (H + 2) * (W + 2))
        out = torch.empty(rows, Cc, device=x.device, dtype=bf16)
        check(lib().b200_grid_sample2(ptr(x), B, H, W, Cc, int(down), ptr(out), stream_ptr()), 'grid_sample2')
        return out

    @staticmethod
    def _wgrad(dy, x):
        """dW [N, K] fp32 = dy[tokens, N]^T x[tokens, K]"""
        tokens, N = dy.shape
        K = x.shape[1]
        tiles = ((N + 127) // 128) * ((K + 255) // 256)
        splits = max(1, min(tokens // 256, (2 * 148 + tiles - 1) // tiles))
        return ops.splitk_reduce(ops.gemm_wgrad(dy, x, splits=splits))

    def _wgrad_taps(self, dy, x, W: int):
        """[N, 9, K] fp32: dW[n, t, k] = sum_p dy[p, n] x[p + shift_t, k], one launch for the nine taps (rows that fall off either
        end read as zero - they pair with ring rows of dy, which are zero anyway)"""
        tokens, N = dy.shape
        K = x.shape[1]
        arr = self._shift_array(W, 1)
        tiles = ((N + 127) // 128) * ((9 * K + 127) // 128)
        splits = lib().b200_gemm_splits(tokens, max(1, min(tokens // 256, (2 * 148 + tiles - 1) // tiles)))
        partial = torch.empty(splits, N, 9 * K, device=dy.device, dtype=torch.float32)
        check(lib().b200_gemm_wgrad_taps(ptr(dy), dy.stride(0), ptr(x), x.stride(0), tokens, N, K, 9, arr, ptr(partial), splits, stream_ptr()), 'gemm_wgrad_taps')
        return ops.splitk_reduce(partial).view(N, 9, K)

    # ------------------------------------------------------------------ forward
    def forward(self, img: torch.Tensor, save: bool) -> torch.Tensor:
        if not img.is_cuda:
            raise B200Error('ResNet.forward needs a CUDA tensor: the B200 path has no CPU fallback')
        abi.require_device()
        m = self.model
        if img.dim() != 4 or img.shape[1] != 3 or img.shape[2] % 32 or img.shape[3] % 32:
            raise B200Error(f'expected (B, 3, H, W) input with H, W multiples of 32, got {tuple(img.shape)}')
        is_u8 = img.dtype == torch.uint8
        img = img.contiguous() if is_u8 else img.contiguous().float()
        B, _, IH, IW = img.shape
        dev = img.device
        training = m.training
        wc = self.wcache
        ctx = {'B': B, 'blocks': [], 'batch_stats': training} if save else None

        # stem: conv1 7x7/2 (im2col of the 3-channel image + GEMM) -> bn1 + relu + maxpool fused
        OH, OW = IH // 2, IW // 2
        cols = torch.empty(B * OH * OW, 160, device=dev, dtype=bf16)
        check(lib().b200_stem_im2col(ptr(img), int(is_u8), B, IH, IW, ptr(cols), stream_ptr()), 'stem_im2col')
        a0 = ops.gemm_tn(cols, wc.get(m.conv1.weight, 'stem')[0])
        c0 = self._bn_consts(m.bn1, a0, 0, 0, a0.shape[0], training)
        H, W = OH // 2, OW // 2
        x = torch.empty(B * (H + 2) * (W + 2), 64, device=dev, dtype=bf16)
        tap = torch.empty(B * H * W, 64, device=dev, dtype=torch.uint8)
        check(lib().b200_stem_pool_fwd(ptr(a0), ptr(c0[0]), ptr(c0[1]), B, OH, OW, 64, ptr(x), ptr(tap), stream_ptr()), 'stem_pool_fwd')
        if save:
            ctx['stem'] = (cols, a0, c0, tap, OH, OW)
        else:
            del cols, a0

        fused_eval = not training and not save          # inference: BatchNorm folded into the convolution launches
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                if fused_eval:
                    x, H, W = self._block_forward_eval(blk, x, B, H, W)
                    continue
                x, H, W, rec = self._block_forward(blk, x, B, H, W, training, save)
                if save:
                    ctx['blocks'].append(rec)

        if self._counters:
            torch._foreach_add_(self._counters, 1)
            self._counters = []
        pooled = torch.empty(B, x.shape[1], device=dev, dtype=bf16)
        check(lib().b200_grid_avgpool(ptr(x), B, H, W, x.shape[1], 0, ptr(pooled), stream_ptr()), 'grid_avgpool')
        fc = m.fc
        emb = ops.gemm_tn(pooled, wc.get(fc.weight, 'linear')[0], bias=fc.bias.detach().float() if fc.bias is not None else None, out_fp32=True)
        if save:
            ctx['head'] = (pooled, H, W, x.shape[1])
            self._token += 1
            ctx['token'] = self._token
            self.saved = ctx
        return emb

    def _block_forward(self, blk, x, B: int, H: int, W: int, training: bool, save: bool):
        wc = self.wcache
        stride = blk.conv2.stride[0]
        if blk.conv1.stride[0] != 1 or blk.conv2.kernel_size != (3, 3) or blk.conv2.groups != 1 or blk.conv2.dilation[0] != 1:
            raise B200Error('ConvNetEngine supports the torchvision v1.5 Bottleneck (stride on the 3x3 convolution, no groups / dilation)')
        n_in = B * H * W
        a1 = ops.gemm_tn(x, wc.get(blk.conv1.weight, 'conv')[0])
        c1 = self._bn_consts(blk.bn1, a1, H, W, n_in, training)
        y1 = self._bn_apply(a1, c1, H, W, True)
        a2 = self._taps(y1, wc.get(blk.conv2.weight, 'conv')[0], W, 1)
        Ho, Wo = H // stride, W // stride
        if stride == 2:
            a2 = self._sample(a2, B, H, W, True)
        elif stride != 1:
            raise B200Error(f'unsupported stride {stride}')
        n_out = B * Ho * Wo
        c2 = self._bn_consts(blk.bn2, a2, Ho, Wo, n_out, training)
        y2 = self._bn_apply(a2, c2, Ho, Wo, True)
        a3 = ops.gemm_tn(y2, wc.get(blk.conv3.weight, 'conv')[0])
        c3 = self._bn_consts(blk.bn3, a3, Ho, Wo, n_out, training)
        xs = ad = cd = None
        if blk.downsample is not None:
            xs = self._sample(x, B, H, W, True) if stride == 2 else x
            ad = ops.gemm_tn(xs, wc.get(blk.downsample[0].weight, 'conv')[0])
            cd = self._bn_consts(blk.downsample[1], ad, Ho, Wo, n_out, training)
            idn = self._bn_apply(ad, cd, Ho, Wo, False)
        else:
            idn = x
        out = self._bn_apply(a3, c3, Ho, Wo, True, residual=idn)
        rec = (blk, x, a1, c1, y1, a2, c2, y2, a3, c3, xs, ad, cd, out, H, W, Ho, Wo) if save else None
        return out, Ho, Wo, rec

    # ------------------------------------------------------------------ backward
    def backward(self, demb: torch.Tensor, token: int) -> Dict[int, torch.Tensor]:
        """gradients of every parameter, keyed by id(parameter), in the parameter's own shape (fp32)"""
        ctx = self.saved
        if ctx is None or ctx['token'] != token:
            raise B200Error('backward through a ResNet forward whose saved activations were released by a later forward')
        self.saved = None
        m, wc, B = self.model, self.wcache, ctx['B']
        self._frozen = not ctx['batch_stats']        # eval-mode BatchNorm under autograd (fine-tuning with frozen statistics)
        grads: Dict[int, torch.Tensor] = {}
        pooled, H, W, Cl = ctx['head']
        fc = m.fc
        d = demb.contiguous().to(bf16)
        grads[id(fc.weight)] = self._wgrad(d, pooled)
        if fc.bias is not None:
            grads[id(fc.bias)] = demb.float().sum(0)
        dpooled = ops.gemm_tn(d, wc.get(fc.weight, 'linear')[1])
        dx = torch.empty(B * (H + 2) * (W + 2), Cl, device=d.device, dtype=bf16)
        check(lib().b200_grid_avgpool(ptr(dpooled), B, H, W, Cl, 1, ptr(dx), stream_ptr()), 'grid_avgpool_bwd')
        for rec in reversed(ctx['blocks']):
            dx = self._block_backward(rec, dx, B, grads)
        cols, a0, c0, tap, OH, OW = ctx['stem']
        dz0 = torch.empty_like(a0)
        check(lib().b200_stem_pool_bwd(ptr(dx), ptr(tap), ptr(a0), ptr(c0[0]), ptr(c0[1]), B, OH, OW, 64, ptr(dz0), stream_ptr()), 'stem_pool_bwd')
        da0, _, s0 = self._bn_backward(dz0, None, a0, c0, m.bn1.weight, 0, 0, a0.shape[0])
        grads[id(m.bn1.weight)], grads[id(m.bn1.bias)] = s0[1], s0[0]
        dw = self._wgrad(da0, cols)[:, :147]
        grads[id(m.conv1.weight)] = dw.reshape(64, 7, 7, 3).permute(0, 3, 1, 2).contiguous()
        return grads

    def _block_backward(self, rec, d_out, B: int, grads) -> torch.Tensor:
        blk, x, a1, c1, y1, a2, c2, y2, a3, c3, xs, ad, cd, out, H, W, Ho, Wo = rec
        wc = self.wcache
        stride = blk.conv2.stride[0]
        n_in, n_out = B * H * W, B * Ho * Wo

        def conv_grad(conv, g):
            grads[id(conv.weight)] = g.reshape(conv.weight.shape)

        def bn_grad(bn, sums):
            grads[id(bn.weight)], grads[id(bn.bias)] = sums[1], sums[0]

        da3, dz3, s3 = self._bn_backward(d_out, out, a3, c3, blk.bn3.weight, Ho, Wo, n_out, want_dz=True)
        bn_grad(blk.bn3, s3)
        conv_grad(blk.conv3, self._wgrad(da3, y2))
        dy2 = ops.gemm_tn(da3, wc.get(blk.conv3.weight, 'conv')[1])
        da2, _, s2 = self._bn_backward(dy2, None, a2, c2, blk.bn2.weight, Ho, Wo, n_out, relu_from_x=True)
        bn_grad(blk.bn2, s2)
        if stride == 2:
            # weight gradient on the coarse grid against gathered patches (a quarter of the contraction of the fine grid)
            cols = torch.empty(da2.shape[0], 9 * y1.shape[1], device=da2.device, dtype=bf16)
            check(lib().b200_grid_patches_s2(ptr(y1), B, H, W, y1.shape[1], ptr(cols), stream_ptr()), 'grid_patches_s2')
            g2 = self._wgrad(da2, cols).view(da2.shape[1], 9, y1.shape[1])
            del cols
            da2 = self._sample(da2, B, H, W, False)
        else:
            g2 = self._wgrad_taps(da2, y1, W)                               # [Co, 9, Ci]
        grads[id(blk.conv2.weight)] = g2.permute(0, 2, 1).reshape(blk.conv2.weight.shape).contiguous()
        dy1 = self._taps(da2, wc.get(blk.conv2.weight, 'conv')[1], W, -1)
        da1, _, s1 = self._bn_backward(dy1, None, a1, c1, blk.bn1.weight, H, W, n_in, relu_from_x=True)
        bn_grad(blk.bn1, s1)
        conv_grad(blk.conv1, self._wgrad(da1, x))
        if blk.downsample is not None:
            dad, _, sd = self._bn_backward(dz3, None, ad, cd, blk.downsample[1].weight, Ho, Wo, n_out)
            bn_grad(blk.downsample[1], sd)
            conv_grad(blk.downsample[0], self._wgrad(dad, xs))
            skip = ops.gemm_tn(dad, wc.get(blk.downsample[0].weight, 'conv')[1])
            if stride == 2:
                skip = self._sample(skip, B, H, W, False)
        else:
            skip = dz3
        # dx = da1 W1 + the gradient arriving over the skip connection (added in the GEMM epilogue)
        return ops.gemm_tn(da1, wc.get(blk.conv1.weight, 'conv')[1], mode=abi.EPI_RESID, aux=skip)


class ConvNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: ConvNetEngine, img, *params):
        emb = engine.forward(img, True)
        ctx.engine = engine
        ctx.token = engine._token
        ctx.param_ids = [id(p) for p in params]
        return emb

    @staticmethod
    def backward(ctx, demb):
        grads = ctx.engine.backward(demb, ctx.token)
        return (None, None) + tuple(grads.get(i) for i in ctx.param_ids)
