"""Tensor-level wrappers over the C ABI: one Python function per exported kernel.

These are what the parity tests drive, and what losses/large_margin.py and b200/gallery.py are built
from.  All tensors must be CUDA tensors; outputs are allocated with torch (caching allocator) and the
kernels run on torch's current stream.
"""
from __future__ import annotations

import torch

from . import abi
from .abi import check, lib, ptr, stream_ptr

bf16 = torch.bfloat16


def _align8(n: int) -> int:
    return (n + 7) // 8 * 8


def gemm_tn(a, b, *, mode=abi.EPI_STORE, bias=None, aux=None, out_fp32=False, want_grad=False, splits=1, block_n=0, out=None):
    """out[M,N] = a[M,K] @ b[N,K]^T with the requested epilogue.  a, b: bf16/fp16 2-D, row pitch = stride(0)."""
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1] and a.dtype == b.dtype
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    dev = a.device
    is_bf16 = 1 if a.dtype == bf16 else 0
    if mode == abi.EPI_PARTIAL:
        splits = lib().b200_gemm_splits(K, max(1, splits))
        out = torch.empty(splits, M, N, device=dev, dtype=torch.float32)
        check(lib().b200_gemm_tn(ptr(a), a.stride(0), ptr(b), b.stride(0), M, N, K, is_bf16, mode, ptr(out), N, 1, 0, 0,
                                 ptr(bias), 0, 0, splits, M * N, block_n, stream_ptr()), 'gemm_tn')
        return out
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_fp32 else bf16)
    # EPI_GELU_Q8: the derivative comes back as 8-bit codes (see include/b200_fe.h); EPI_DGELU_Q8 takes them as `aux`
    if mode == abi.EPI_GELU_Q8:
        dact = torch.empty(M, N, device=dev, dtype=torch.uint8)
    else:
        dact = torch.empty(M, N, device=dev, dtype=bf16) if (mode == abi.EPI_GELU and want_grad) else None
    check(lib().b200_gemm_tn(ptr(a), a.stride(0), ptr(b), b.stride(0), M, N, K, is_bf16, mode, ptr(out), out.stride(0),
                             1 if out.dtype == torch.float32 else 0, ptr(dact), N, ptr(bias), ptr(aux),
                             aux.stride(0) if aux is not None else 0, 1, 0, block_n, stream_ptr()), 'gemm_tn')
    return (out, dact) if dact is not None else out


def gelu_q8_decode(codes):
    """uint8 codes of B200_EPI_GELU_Q8 -> the GELU derivative they stand for (fp32)."""
    return codes.float() * (1.26 / 255.0) - 0.13


def gemm_wgrad(dy, x, splits=1, block_n=0):
    """fp32 partials [splits, N, K] of dy[tokens, N]^T @ x[tokens, K] (bf16 operands read in place)."""
    tokens, N = dy.shape
    K = x.shape[1]
    splits = lib().b200_gemm_splits(tokens, max(1, splits))
    out = torch.empty(splits, N, K, device=dy.device, dtype=torch.float32)
    check(lib().b200_gemm_wgrad(ptr(dy), dy.stride(0), ptr(x), x.stride(0), tokens, N, K, ptr(out), splits, block_n, stream_ptr()), 'gemm_wgrad')
    return out


def gemm_wgrad_bias(dy, x, splits=1, block_n=0):
    """gemm_wgrad that also returns the bias-gradient partial rows [splits, N] (column sums of dy from an all-ones MMA in the
    same kernel), or None for them when the tile shape cannot host the extra accumulator columns."""
    import ctypes
    tokens, N = dy.shape
    K = x.shape[1]
    splits = lib().b200_gemm_splits(tokens, max(1, splits))
    out = torch.empty(splits, N, K, device=dy.device, dtype=torch.float32)
    cs = torch.empty(splits, N, device=dy.device, dtype=torch.float32)
    fused = ctypes.c_int(0)
    check(lib().b200_gemm_wgrad_bias(ptr(dy), dy.stride(0), ptr(x), x.stride(0), tokens, N, K, ptr(out), ptr(cs), splits, block_n,
                                     ctypes.byref(fused), stream_ptr()), 'gemm_wgrad_bias')
    return out, (cs if fused.value else None)


def splitk_reduce(partial, out=None, accumulate=False):
    splits, n = partial.shape[0], partial[0].numel()
    if out is None:
        out = torch.empty(partial.shape[1:], device=partial.device, dtype=torch.float32)
    check(lib().b200_splitk_reduce(ptr(partial), ptr(out), n, splits, int(accumulate), stream_ptr()), 'splitk_reduce')
    return out


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    M, Cc = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    check(lib().b200_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), M, Cc, eps, stream_ptr()), 'layernorm_fwd')
    return y, mean, rstd


def layernorm_fwd_windows(x, gamma, beta, B, H, W, shifted, eps=1e-5):
    """LayerNorm whose output rows leave in window-major order (mean / rstd stay per raster row)."""
    M, Cc = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    check(lib().b200_layernorm_fwd_windows(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), B, H, W, Cc, int(shifted), eps,
                                           stream_ptr()), 'layernorm_fwd_windows')
    return y, mean, rstd


def layernorm_bwd_windows(dy, x, gamma, mean, rstd, B, H, W, shifted, dres=None):
    """LayerNorm backward whose dy rows arrive in window-major order."""
    M, Cc = x.shape
    dx = torch.empty_like(x)
    dgb = torch.empty(3, Cc, device=x.device, dtype=torch.float32)
    blocks = lib().b200_layernorm_bwd_blocks(M, Cc)
    partial = torch.empty(blocks, 3 * Cc, device=x.device, dtype=torch.float32)
    check(lib().b200_layernorm_bwd_windows(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dgb[0]), ptr(dgb[1]),
                                           0, ptr(partial), B, H, W, Cc, int(shifted), 0, stream_ptr()), 'layernorm_bwd_windows')
    return dx, dgb[0], dgb[1]


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, want_dres_colsum=False):
    M, Cc = x.shape
    dx = torch.empty_like(x)
    dgb = torch.empty(3, Cc, device=x.device, dtype=torch.float32)
    blocks = lib().b200_layernorm_bwd_blocks(M, Cc)
    partial = torch.empty(blocks, 3 * Cc, device=x.device, dtype=torch.float32)
    check(lib().b200_layernorm_bwd(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dgb[0]), ptr(dgb[1]),
                                   ptr(dgb[2]) if want_dres_colsum else 0, ptr(partial), M, Cc, 0, stream_ptr()), 'layernorm_bwd')
    return (dx, dgb[0], dgb[1], dgb[2]) if want_dres_colsum else (dx, dgb[0], dgb[1])


def patch_gather_image(img, df=4):
    B, Cin, H, W = img.shape
    cols = torch.empty(B * (H // df) * (W // df), Cin * df * df, device=img.device, dtype=bf16)
    check(lib().b200_patch_gather_image(ptr(img), ptr(cols), B, Cin, H, W, df, cols.stride(0), stream_ptr()), 'patch_gather_image')
    return cols


def patch_gather_nhwc(x, B, H, W, Cc):
    cols = torch.empty(B * (H // 2) * (W // 2), 4 * Cc, device=x.device, dtype=bf16)
    check(lib().b200_patch_gather_nhwc(ptr(x), ptr(cols), B, H, W, Cc, 0, stream_ptr()), 'patch_gather_nhwc')
    return cols


def patch_scatter_nhwc(dcols, B, H, W, Cc):
    dx = torch.empty(B * H * W, Cc, device=dcols.device, dtype=bf16)
    check(lib().b200_patch_gather_nhwc(ptr(dx), ptr(dcols), B, H, W, Cc, 1, stream_ptr()), 'patch_gather_nhwc(bwd)')
    return dx


def mean_pool(x, B, T, Cc):
    y = torch.empty(B, Cc, device=x.device, dtype=bf16)
    check(lib().b200_mean_pool(ptr(x), ptr(y), B, T, Cc, 0, stream_ptr()), 'mean_pool')
    return y


def mean_pool_bwd(dy, B, T, Cc):
    dx = torch.empty(B * T, Cc, device=dy.device, dtype=bf16)
    check(lib().b200_mean_pool(ptr(dy), ptr(dx), B, T, Cc, 1, stream_ptr()), 'mean_pool(bwd)')
    return dx


def transpose16(x, pad_to=8):
    R, Cc = x.shape
    ld = (R + pad_to - 1) // pad_to * pad_to
    out = torch.zeros(Cc, ld, device=x.device, dtype=x.dtype)
    check(lib().b200_transpose16(ptr(x), ptr(out), R, Cc, x.stride(0), ld, stream_ptr()), 'transpose16')
    return out[:, :R]


def cast_transpose(w, want=True, want_t=True):
    R, Cc = w.shape
    d = torch.empty(R, Cc, device=w.device, dtype=bf16) if want else None
    dt = torch.empty(Cc, R, device=w.device, dtype=bf16) if want_t else None
    check(lib().b200_cast_transpose(ptr(w), ptr(d), ptr(dt), R, Cc, stream_ptr()), 'cast_transpose')
    return d, dt


def colsum(x):
    M, N = x.shape
    out = torch.empty(N, device=x.device, dtype=torch.float32)
    partial = torch.empty(lib().b200_colsum_blocks(M), N, device=x.device, dtype=torch.float32)
    check(lib().b200_colsum(ptr(x), x.stride(0), M, N, ptr(out), ptr(partial), 0, stream_ptr()), 'colsum')
    return out


def window_rows(x, B, H, W, shifted, to_window=True):
    """Rows of a [B*H*W, n] tensor permuted raster -> window-major order (to_window) or back: the row order the attention
    kernels take their q / k / v in (include/b200_fe.h: b200_window_rows)."""
    x = x.contiguous()
    out = torch.empty_like(x)
    check(lib().b200_window_rows(ptr(x), ptr(out), B, H, W, x.shape[1] * x.element_size(), int(shifted), int(to_window), stream_ptr()),
          'window_rows')
    return out


def window_attn_fwd(qkv, pos, B, H, W, Cc, heads, shifted, want_lse=True, window_major=False):
    """qkv [B*H*W, 3C] in raster order (window_major=False: permuted here) or already window-major.  Returns the attention output
    in raster order and the row log-sum-exp in WINDOW-MAJOR order (what window_attn_bwd takes back)."""
    if not window_major:
        qkv = window_rows(qkv, B, H, W, shifted, True)
    out = torch.empty(B * H * W, Cc, device=qkv.device, dtype=bf16)
    lse = torch.empty(B * H * W, heads, device=qkv.device, dtype=torch.float32) if want_lse else None
    check(lib().b200_window_attn_fwd(ptr(qkv), ptr(pos), ptr(out), ptr(lse), B, H, W, Cc, heads, int(shifted), stream_ptr()), 'window_attn_fwd')
    return out, lse


def window_attn_bwd(qkv, pos, lse, dout, B, H, W, Cc, heads, shifted, window_major=False):
    """qkv / the returned dqkv in raster order (window_major=False) or window-major; lse as returned by window_attn_fwd."""
    if not window_major:
        qkv = window_rows(qkv, B, H, W, shifted, True)
    dqkv = torch.empty_like(qkv)
    dpos = torch.empty(169, device=qkv.device, dtype=torch.float32)
    blocks = lib().b200_window_attn_bwd_blocks(B, H, W, heads)
    partial = torch.empty(lib().b200_window_attn_bwd_scratch_floats(blocks), device=qkv.device, dtype=torch.float32)
    check(lib().b200_window_attn_bwd(ptr(qkv), ptr(pos), ptr(lse), ptr(dout), ptr(dqkv), ptr(dpos), ptr(partial), 0,
                                     B, H, W, Cc, heads, int(shifted), stream_ptr()), 'window_attn_bwd')
    if not window_major:
        dqkv = window_rows(dqkv, B, H, W, shifted, False)
    return dqkv, dpos.view(13, 13)


def unit_rows(x, as_f16=False, eps=1e-12, ld=None):
    R, E = x.shape
    ld = ld or E
    out = torch.zeros(R, ld, device=x.device, dtype=torch.float16 if as_f16 else bf16)
    inv = torch.empty(R, device=x.device, dtype=torch.float32)
    check(lib().b200_unit_rows(ptr(x), ptr(out), ptr(inv), R, E, ld, eps, int(as_f16), stream_ptr()), 'unit_rows')
    return out, inv


class MarginHeadFunction(torch.autograd.Function):
    """loss, logits = head(emb, weight, label): F.normalize both, cosine GEMM with the margin + scale
    epilogue, focal / cross-entropy loss - and, in the same pass, the gradient of the loss wrt the
    cosines, so backward is two GEMMs plus the normalisation backward."""

    @staticmethod
    def forward(ctx, emb, weight, label, s, m, kind, easy_margin, gamma):
        L = lib()
        st = stream_ptr()
        dev = emb.device
        B, E = emb.shape
        Cn = weight.shape[0]
        emb = emb.contiguous().float()
        w = weight.contiguous().float()
        label = label.contiguous().long()
        e16, e_inv = unit_rows(emb)
        w16, w_inv = unit_rows(w)
        logits = torch.empty(B, Cn, device=dev, dtype=torch.float32)
        cos_label = torch.empty(B, device=dev, dtype=torch.float32)
        check(L.b200_margin_logits(ptr(e16), ptr(w16), B, Cn, E, ptr(label), s, m, kind, int(easy_margin), ptr(logits), Cn,
                                   ptr(cos_label), st), 'margin_logits')
        need_grad = emb.requires_grad or weight.requires_grad or ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        loss_rows = torch.empty(B, device=dev, dtype=torch.float32)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        ldg = _align8(Cn)
        G = torch.zeros(B, ldg, device=dev, dtype=bf16) if need_grad else None
        rdot = torch.empty(B, device=dev, dtype=torch.float32) if need_grad else None
        cdot = torch.empty(Cn, device=dev, dtype=torch.float32) if need_grad else None
        check(L.b200_margin_ce(ptr(logits), Cn, ptr(label), ptr(cos_label), B, Cn, s, m, kind, int(easy_margin), gamma,
                               ptr(loss_rows), ptr(loss), ptr(G), ldg, ptr(rdot), ptr(cdot), st), 'margin_ce')
        if need_grad:
            ctx.save_for_backward(emb, w, e16, w16, e_inv, w_inv, G, rdot, cdot)
        ctx.dims = (B, Cn, E)
        ctx.mark_non_differentiable(logits)
        return loss.reshape(()), logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        emb, w, e16, w16, e_inv, w_inv, G, rdot, cdot = ctx.saved_tensors
        B, Cn, E = ctx.dims
        L = lib()
        st = stream_ptr()
        scale = dloss.reshape(1).float().contiguous()
        demb = dw = None
        if ctx.needs_input_grad[0]:
            w16t = transpose16(w16)                                            # [E, C] (pitch padded to 8)
            t1 = gemm_tn(G[:, :Cn], w16t, out_fp32=True)                       # G @ w^  -> [B, E]
            demb = torch.empty_like(emb)
            check(L.b200_unit_rows_bwd(ptr(t1), ptr(emb), ptr(e_inv), ptr(rdot), ptr(scale), ptr(demb), 0, B, E, 0, st), 'unit_rows_bwd')
        if ctx.needs_input_grad[1]:
            gt = transpose16(G[:, :Cn])                                        # [C, B]
            e16t = transpose16(e16)                                            # [E, B]
            t2 = gemm_tn(gt, e16t, out_fp32=True)                              # G^T @ e^ -> [C, E]
            dw = torch.empty_like(w)
            check(L.b200_unit_rows_bwd(ptr(t2), ptr(w), ptr(w_inv), ptr(cdot), ptr(scale), ptr(dw), 0, Cn, E, 0, st), 'unit_rows_bwd')
        return demb, dw, None, None, None, None, None, None


def margin_head(emb, weight, label, s=64.0, m=0.5, kind=0, easy_margin=False, gamma=0.0):
    return MarginHeadFunction.apply(emb, weight, label, float(s), float(m), int(kind), bool(easy_margin), float(gamma))


class FocalLossFunction(torch.autograd.Function):
    """Stand-alone focal / cross-entropy loss of given logits (csrc/arcface.cu: focal_rows_kernel): the criterion object of
    the reference (losses/losses.py:22-28) called on its own, outside the fused head."""

    @staticmethod
    def forward(ctx, logits, target, gamma):
        B, Cn = logits.shape
        x = logits.contiguous().float()
        t = target.contiguous().long()
        rows = torch.empty(B, device=x.device, dtype=torch.float32)
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        d = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        check(lib().b200_focal_loss(ptr(x), x.stride(0), ptr(t), B, Cn, float(gamma), ptr(rows), ptr(loss), ptr(d), Cn, stream_ptr()),
              'focal_loss')
        ctx.save_for_backward(d)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (d,) = ctx.saved_tensors
        return (d * dloss if d is not None else None), None, None


def focal_loss(logits, target, gamma=0.0):
    return FocalLossFunction.apply(logits, target, float(gamma))
