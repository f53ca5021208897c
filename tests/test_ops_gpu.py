"""Kernel-level parity (B200 only): every exported kernel, called through the C ABI, against the fp32
oracle / a plain torch fp32 statement of the same op on the same seeded inputs.

Tolerances: operands are bf16 (8 mantissa bits, eps = 2^-8 = 3.9e-3) with fp32 accumulation, so outputs are
compared after the same bf16 rounding of the inputs, at ~2 bf16 ulps of the output magnitude.
"""
import math

from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

DEV = 'cuda'
bf16 = torch.bfloat16


@pytest.fixture(scope='module', autouse=True)
def _need_gpu():
    from b200 import abi
    abi.require_device()


def rnd(*shape, seed=0, scale=1.0, dtype=bf16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def rel_err(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize('M,N,K', [(128, 96, 64), (256, 96, 48), (1000, 192, 96), (130, 288, 96), (4096, 384, 96),
                                   (392, 768, 3072), (257, 2304, 768), (64, 10000, 512), (2, 512, 768), (6272, 96, 384)])
def test_gemm_store(M, N, K):
    from b200 import ops
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2)
    bias = rnd(N, seed=3, dtype=torch.float32)
    ref = a.float() @ b.float().t() + bias
    out = ops.gemm_tn(a, b, bias=bias)
    assert out.dtype == bf16 and out.shape == (M, N)
    assert rel_err(out, ref) < 6e-3
    out32 = ops.gemm_tn(a, b, bias=bias, out_fp32=True)
    assert rel_err(out32, ref) < 1e-5 * math.sqrt(K) + 1e-6
    assert (out32 - ref).abs().max().item() < 1e-3 * math.sqrt(K)


def test_gemm_fp16_inputs():
    from b200 import ops
    a, b = rnd(300, 512, seed=4, dtype=torch.float16), rnd(704, 512, seed=5, dtype=torch.float16)
    out = ops.gemm_tn(a, b, out_fp32=True)
    assert rel_err(out, a.float() @ b.float().t()) < 1e-4


def test_gemm_epilogues():
    from b200 import abi, ops
    M, N, K = 1024, 384, 96
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.2)
    bias = rnd(N, seed=3, dtype=torch.float32)
    pre_ref = a.float() @ b.float().t() + bias
    act, dact = ops.gemm_tn(a, b, bias=bias, mode=abi.EPI_GELU, want_grad=True)     # second output: gelu'(pre-activation)
    pr = pre_ref.clone().requires_grad_(True)
    torch.nn.functional.gelu(pr).sum().backward()
    assert rel_err(dact, pr.grad) < 6e-3
    assert rel_err(act, torch.nn.functional.gelu(pre_ref)) < 6e-3
    assert rel_err(ops.gemm_tn(a, b, bias=bias, mode=abi.EPI_GELU), torch.nn.functional.gelu(pre_ref)) < 6e-3
    res = rnd(M, N, seed=7)
    out = ops.gemm_tn(a, b, bias=bias, mode=abi.EPI_RESID, aux=res)
    assert rel_err(out, pre_ref + res.float()) < 6e-3
    x = rnd(M, N, seed=8)
    out = ops.gemm_tn(a, b, mode=abi.EPI_DGELU, aux=x)          # aux = the derivative saved by the GELU epilogue
    assert rel_err(out, (a.float() @ b.float().t()) * x.float()) < 6e-3


@pytest.mark.parametrize('M,N,K', [(1100, 384, 384), (6272, 1536, 384), (2000, 384, 1536), (50176, 1152, 384), (257, 768, 3072), (1153, 2304, 768)])
def test_gemm_cta_pair_variant(M, N, K):
    """The cta_group::2 kernel (stages 3-4: K >= 384, whole tiles of >= 192 columns): every epilogue against fp32 matmul, odd and
    even numbers of row blocks (a phantom row block in the last pair), more tiles than clusters, and bit-equality with the
    single-CTA kernel (B200_CG2 is read once per process, so that comparison runs in a child process)."""
    import os, subprocess, sys
    from b200 import abi, ops
    a, b = rnd(M, K, seed=1, scale=0.5), rnd(N, K, seed=2, scale=0.1)
    bias = rnd(N, seed=3, dtype=torch.float32)
    pre = a.float() @ b.float().t() + bias
    out = ops.gemm_tn(a, b, bias=bias)
    assert rel_err(out, pre) < 6e-3
    act, dact = ops.gemm_tn(a, b, bias=bias, mode=abi.EPI_GELU, want_grad=True)
    pr = pre.clone().requires_grad_(True)
    torch.nn.functional.gelu(pr).sum().backward()
    assert rel_err(act, torch.nn.functional.gelu(pre)) < 6e-3 and rel_err(dact, pr.grad) < 6e-3
    res = rnd(M, N, seed=7)
    assert rel_err(ops.gemm_tn(a, b, bias=bias, mode=abi.EPI_RESID, aux=res), pre + res.float()) < 6e-3
    x = rnd(M, N, seed=8)
    assert rel_err(ops.gemm_tn(a, b, mode=abi.EPI_DGELU, aux=x), (a.float() @ b.float().t()) * x.float()) < 6e-3
    if (M, N, K) == (1100, 384, 384):
        code = ("import sys, torch; sys.path[:0] = %r; from b200 import ops; g = torch.Generator().manual_seed(1); "
                "a = (torch.randn(1100, 384, generator=g) * 0.5).cuda().bfloat16(); g2 = torch.Generator().manual_seed(2); "
                "b = (torch.randn(384, 384, generator=g2) * 0.1).cuda().bfloat16(); torch.save(ops.gemm_tn(a, b).cpu(), sys.argv[1])") % ([str(ROOT), str(ROOT / 'pets-face-recognition_b200')],)
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            subprocess.run([sys.executable, '-c', code, d + '/single.pt'], check=True, env=dict(os.environ, B200_CG2='0'), timeout=300)
            single = torch.load(d + '/single.pt')
        assert torch.equal(ops.gemm_tn(a, b).cpu(), single)


@pytest.mark.parametrize('M,N,K', [(1024, 384, 96), (777, 768, 192), (130, 1536, 384), (50, 3072, 768)])
def test_gemm_gelu_derivative_as_8bit_codes(M, N, K):
    """B200_EPI_GELU_Q8 / B200_EPI_DGELU_Q8: the saved GELU derivative travels as uint8 codes (step 1.26 / 255): decoded it is
    within half a step (+ the bf16 rounding of the pre-activation path) of torch's derivative, and the data-gradient epilogue
    that consumes the codes equals acc * decode(codes)."""
    from b200 import abi, ops
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.2)
    bias = rnd(N, seed=3, dtype=torch.float32)
    pre_ref = a.float() @ b.float().t() + bias
    act, codes = ops.gemm_tn(a, b, bias=bias, mode=abi.EPI_GELU_Q8)
    assert codes.dtype == torch.uint8 and codes.shape == (M, N)
    pr = pre_ref.clone().requires_grad_(True)
    torch.nn.functional.gelu(pr).sum().backward()
    assert rel_err(act, torch.nn.functional.gelu(pre_ref)) < 6e-3
    dec = ops.gelu_q8_decode(codes)
    assert (dec - pr.grad).abs().max().item() < 0.5 * 1.26 / 255 + 2e-3
    assert rel_err(dec, pr.grad) < 6e-3
    g, wt = rnd(M, K, seed=5), rnd(N, K, seed=6, scale=0.2)
    out = ops.gemm_tn(g, wt, mode=abi.EPI_DGELU_Q8, aux=codes)
    assert rel_err(out, (g.float() @ wt.float().t()) * dec) < 6e-3


@pytest.mark.parametrize('M,N,K,splits', [(384, 96, 50000, 37), (96, 48, 6272, 148), (768, 256, 1000, 3), (512, 768, 2, 4)])
def test_gemm_splitk_partial(M, N, K, splits):
    from b200 import abi, ops
    Kp = (K + 7) // 8 * 8     # row pitch must be a multiple of 16 B; K itself may be anything (TMA zero-fills)
    a, b = rnd(M, Kp, seed=1, scale=0.1)[:, :K], rnd(N, Kp, seed=2, scale=0.1)[:, :K]
    part = ops.gemm_tn(a, b, mode=abi.EPI_PARTIAL, splits=splits)
    out = ops.splitk_reduce(part)
    ref = a.float() @ b.float().t()
    assert rel_err(out, ref) < 2e-5 * math.sqrt(K) + 1e-6
    out2 = ops.splitk_reduce(part, out.clone(), accumulate=True)
    assert rel_err(out2, 2 * ref) < 2e-5 * math.sqrt(K) + 1e-6


@pytest.mark.parametrize('tokens,N,K,splits', [(6272, 384, 96, 49), (1000, 96, 48, 7), (130, 768, 3072, 2), (2, 512, 768, 1),
                                               (50000, 288, 96, 148), (392, 2304, 768, 3), (777, 96, 384, 5)])
def test_gemm_wgrad_mn_major(tokens, N, K, splits):
    """dW = dY^T X straight from the row-major activations (MN-major tcgen05 operands), vs fp32 matmul."""
    from b200 import ops
    dy, x = rnd(tokens, N, seed=1, scale=0.1), rnd(tokens, K, seed=2, scale=0.1)
    out = ops.splitk_reduce(ops.gemm_wgrad(dy, x, splits=splits))
    ref = dy.float().t() @ x.float()
    assert out.shape == (N, K)
    assert rel_err(out, ref) < 2e-5 * math.sqrt(tokens) + 1e-6


@pytest.mark.parametrize('tokens,N,K,splits', [(3136, 384, 96, 7), (1000, 200, 192, 3), (50, 1536, 384, 1), (777, 96, 48, 148), (640, 3072, 768, 4)])
def test_gemm_wgrad_fused_bias_gradient(tokens, N, K, splits):
    """The weight-gradient kernel's all-ones MMA: column sums of dY (= the bias gradient) from the same launch; shapes whose
    tile leaves no spare TMEM columns (K = 768 -> 256-wide tiles) must report `not fused` and leave the weight gradient intact."""
    from b200 import ops
    dy, x = rnd(tokens, N, seed=1, scale=0.1), rnd(tokens, K, seed=2, scale=0.1)
    part, cs = ops.gemm_wgrad_bias(dy, x, splits=splits)
    out = ops.splitk_reduce(part)
    assert rel_err(out, dy.float().t() @ x.float()) < 2e-5 * math.sqrt(tokens) + 1e-6
    if K == 768:
        assert cs is None
        return
    assert cs is not None and cs.shape == (part.shape[0], N)
    db = cs.sum(0)
    ref = dy.float().sum(0)
    assert (db - ref).abs().max().item() < 1e-5 * math.sqrt(tokens) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize('C', [96, 192, 384, 768])
@pytest.mark.parametrize('M', [1, 49, 1000])
def test_layernorm(C, M):
    from b200 import ops
    x = rnd(M, C, seed=1, scale=2.0) + 0.5
    g = 1 + 0.1 * rnd(C, seed=2, dtype=torch.float32)
    b = 0.1 * rnd(C, seed=3, dtype=torch.float32)
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5)
    assert (y.float() - yr).abs().max().item() < 4e-2 and rel_err(y, yr) < 5e-3
    np.testing.assert_allclose(mean.cpu().numpy(), x.float().mean(1).cpu().numpy(), rtol=1e-4, atol=1e-5)
    dy = rnd(M, C, seed=4)
    dres = rnd(M, C, seed=5)
    yr.backward(dy.float())
    dx, dgam, dbet = ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dres)
    assert rel_err(dx, xr.grad + dres.float()) < 6e-3
    assert rel_err(dgam, gr.grad) < 1e-3 and rel_err(dbet, br.grad) < 1e-3
    dx2, dgam2, dbet2, dcol = ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, want_dres_colsum=True)
    assert torch.equal(dx2, dx) and torch.equal(dgam2, dgam)
    assert rel_err(dcol, dres.float().sum(0)) < 1e-4          # fused bias gradient of the residual branch's Linear


@pytest.mark.parametrize('B,H,W,C,shifted', [(2, 14, 14, 96, 0), (3, 14, 21, 192, 1), (1, 7, 7, 768, 1), (5, 28, 28, 96, 1)])
def test_window_major_rows_and_layernorm_variants(B, H, W, C, shifted):
    """b200_window_rows == the reference's roll(-3) + window rearrange (models/swin.py:8-14, :112); the window-major LayerNorm
    forward / backward equal the raster kernels composed with that permutation, bit for bit."""
    from b200 import ops
    M = B * H * W
    x = rnd(M, C, seed=1, scale=2.0)
    wm = ops.window_rows(x, B, H, W, shifted, True)
    ref = x.view(B, H, W, C)
    if shifted:
        ref = torch.roll(ref, shifts=(-3, -3), dims=(1, 2))
    ref = ref.view(B, H // 7, 7, W // 7, 7, C).permute(0, 1, 3, 2, 4, 5).reshape(M, C)       # b (nw_h nw_w) (w_h w_w)
    assert torch.equal(wm, ref)
    assert torch.equal(ops.window_rows(wm, B, H, W, shifted, False), x)
    g = 1 + 0.1 * rnd(C, seed=2, dtype=torch.float32)
    b = 0.1 * rnd(C, seed=3, dtype=torch.float32)
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    yw, mean_w, rstd_w = ops.layernorm_fwd_windows(x, g, b, B, H, W, shifted)
    assert torch.equal(yw, ops.window_rows(y, B, H, W, shifted, True)) and torch.equal(mean, mean_w) and torch.equal(rstd, rstd_w)
    dy = rnd(M, C, seed=4)
    dres = rnd(M, C, seed=5)
    dx, dgam, dbet = ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dres)
    dxw, dgam_w, dbet_w = ops.layernorm_bwd_windows(ops.window_rows(dy, B, H, W, shifted, True), x, g, mean, rstd, B, H, W, shifted, dres=dres)
    assert torch.equal(dx, dxw) and torch.equal(dgam, dgam_w) and torch.equal(dbet, dbet_w)


def test_patch_gather_matches_unfold():
    from b200 import ops
    from oracle.swin_oracle import patch_merge
    img = torch.rand(3, 3, 56, 56, generator=torch.Generator().manual_seed(0)).to(DEV)
    cols = ops.patch_gather_image(img)
    eye = torch.eye(48, device=DEV)
    ref = patch_merge(img.permute(0, 2, 3, 1), eye, torch.zeros(48, device=DEV), 4).reshape(-1, 48)
    assert torch.equal(cols, ref.to(bf16))
    # same order as nn.Unfold itself (models/swin.py:159,165)
    unf = torch.nn.Unfold(4, stride=4)(img).view(3, 48, 14, 14).permute(0, 2, 3, 1).reshape(-1, 48)
    assert torch.equal(cols, unf.to(bf16))
    B, H, W, C = 2, 14, 28, 64
    x = rnd(B * H * W, C, seed=3)
    cols = ops.patch_gather_nhwc(x, B, H, W, C)
    unf = torch.nn.Unfold(2, stride=2)(x.float().view(B, H, W, C).permute(0, 3, 1, 2)).view(B, 4 * C, H // 2, W // 2)
    assert torch.equal(cols, unf.permute(0, 2, 3, 1).reshape(-1, 4 * C).to(bf16))
    back = ops.patch_scatter_nhwc(cols, B, H, W, C)      # scatter is the exact inverse
    assert torch.equal(back, x)


def test_mean_pool_transpose_colsum_cast():
    from b200 import ops
    B, T, C = 5, 49, 768
    x = rnd(B * T, C, seed=1)
    y = ops.mean_pool(x, B, T, C)
    assert rel_err(y, x.float().view(B, T, C).mean(1)) < 4e-3
    dy = rnd(B, C, seed=2)
    dx = ops.mean_pool_bwd(dy, B, T, C)
    assert rel_err(dx, (dy.float() / T)[:, None, :].expand(B, T, C).reshape(-1, C)) < 4e-3
    for R, Cc in [(100, 96), (3136, 384), (7, 48), (129, 130)]:
        m = rnd(R, Cc, seed=R)
        assert torch.equal(ops.transpose16(m), m.t())
    w = rnd(288, 96, seed=9, dtype=torch.float32)
    d, dt = ops.cast_transpose(w)
    assert torch.equal(d, w.to(bf16)) and torch.equal(dt, w.to(bf16).t())
    m = rnd(5000, 384, seed=11)
    assert rel_err(ops.colsum(m), m.float().sum(0)) < 1e-5


@pytest.mark.parametrize('heads,H,W,shifted,B', [(3, 14, 14, 0, 2), (3, 14, 14, 1, 2), (6, 7, 7, 1, 2), (2, 21, 14, 1, 2), (24, 7, 7, 0, 2),
                                                  (1, 28, 28, 1, 2), (3, 7, 7, 1, 1), (4, 14, 7, 1, 3), (3, 56, 56, 1, 4), (6, 28, 28, 0, 7)])
def test_window_attention_fwd_bwd(heads, H, W, shifted, B):
    """tcgen05 window attention vs the oracle: odd / even head counts (a lone head in the last head pair), odd window counts
    (a lone window in the last window pair), more work units than SMs (persistent loops, barrier phase flips)."""
    from b200 import ops
    from oracle.swin_oracle import attention_core
    C = heads * 32
    qkv = rnd(B * H * W, 3 * C, seed=1)
    pos = rnd(13, 13, seed=2, dtype=torch.float32)
    out, lse = ops.window_attn_fwd(qkv, pos, B, H, W, C, heads, shifted)
    q_ref = qkv.float().cpu().view(B, H, W, 3 * C).requires_grad_(True)
    p_ref = pos.cpu().clone().requires_grad_(True)
    ref = attention_core(q_ref, p_ref, heads, 32, 7, bool(shifted))
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.cpu(), ref.reshape(-1, C)) < 8e-3
    dout = rnd(B * H * W, C, seed=3)
    ref.backward(dout.float().cpu().view(B, H, W, C))
    dqkv, dpos = ops.window_attn_bwd(qkv, pos, lse, dout, B, H, W, C, heads, shifted)
    assert torch.isfinite(dqkv.float()).all()
    assert rel_err(dqkv.cpu(), q_ref.grad.reshape(-1, 3 * C)) < 1.5e-2
    assert rel_err(dpos.cpu(), p_ref.grad) < 1e-2


def test_margin_head_against_oracle_and_golden(golden_dir):
    from b200 import ops
    from oracle import head_oracle
    g = np.load(golden_dir / 'heads_small.npz')
    e = torch.tensor(g['emb']).to(DEV).requires_grad_(True)
    w = torch.tensor(g['weight']).to(DEV).requires_grad_(True)
    lab = torch.tensor(g['label']).to(DEV)
    loss, logits = ops.margin_head(e, w, lab, 64.0, 0.5, 0, False, 0.0)
    # bf16 unit vectors: |d cos| <~ 2^-9 -> |d logit| <~ 64 * 4e-3
    np.testing.assert_allclose(logits.cpu().numpy(), g['arcface'], atol=0.35)
    assert abs(loss.item() - float(g['focal_g0'])) < 0.05 * float(g['focal_g0']) + 0.05
    loss.backward()
    assert rel_err(e.grad.cpu(), torch.tensor(g['demb'])) < 3e-2
    assert rel_err(w.grad.cpu(), torch.tensor(g['dweight'])) < 3e-2
    _, lc = ops.margin_head(e.detach(), w.detach(), lab, 64.0, 0.5, 1, False, 0.0)
    np.testing.assert_allclose(lc.cpu().numpy(), g['cosface'], atol=0.35)
    # larger random case, gamma = 2, vs the oracle
    B, Cn = 64, 1000
    e2 = rnd(B, 512, seed=1, dtype=torch.float32).requires_grad_(True)
    w2 = rnd(Cn, 512, seed=2, dtype=torch.float32, scale=0.05).requires_grad_(True)
    lab2 = torch.randint(0, Cn, (B,), generator=torch.Generator().manual_seed(3)).to(DEV)
    for gamma in (0.0, 2.0):
        e2.grad = w2.grad = None
        loss, logits = ops.margin_head(e2, w2, lab2, 64.0, 0.5, 0, False, gamma)
        ec, wc = e2.detach().cpu().requires_grad_(True), w2.detach().cpu().requires_grad_(True)
        lref = head_oracle.arcface_logits(ec, wc, lab2.cpu(), clamp_sine=True)
        ref = head_oracle.focal_loss(lref, lab2.cpu(), gamma)
        ref.backward()
        assert (logits.cpu() - lref).abs().max().item() < 0.35
        assert abs(loss.item() - ref.item()) < 2e-2 * abs(ref.item())
        loss.backward()
        assert rel_err(e2.grad.cpu(), ec.grad) < 3e-2
        assert rel_err(w2.grad.cpu(), wc.grad) < 3e-2


@pytest.mark.parametrize('gamma', [0.0, 2.0])
def test_standalone_focal_loss_and_bad_labels(gamma):
    """losses.FocalLoss.forward(logits, target) on its own (losses/losses.py:22-28): value and gradient vs the oracle; a label
    outside [0, C) must poison the loss (NaN), in the stand-alone kernel and in the fused head alike."""
    from b200 import ops
    from losses import FocalLoss
    from oracle import head_oracle
    B, Cn = 37, 1000
    x = rnd(B, Cn, seed=1, dtype=torch.float32, scale=3.0).requires_grad_(True)
    lab = torch.randint(0, Cn, (B,), generator=torch.Generator().manual_seed(3)).to(DEV)
    crit = FocalLoss(num_class=Cn, gamma=gamma)
    loss = crit(x, lab)
    xc = x.detach().cpu().requires_grad_(True)
    ref = head_oracle.focal_loss(xc, lab.cpu(), gamma)
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    (2 * loss).backward()
    (2 * ref).backward()
    assert rel_err(x.grad.cpu(), xc.grad) < 1e-4
    bad = lab.clone(); bad[5] = Cn
    assert math.isnan(crit(x.detach(), bad).item())
    e = rnd(B, 512, seed=4, dtype=torch.float32)
    w = rnd(Cn, 512, seed=5, dtype=torch.float32, scale=0.05)
    loss_bad, _ = ops.margin_head(e, w, bad, 64.0, 0.5, 0, False, gamma)
    assert math.isnan(loss_bad.item())
    bad[5] = -1
    loss_bad, _ = ops.margin_head(e, w, bad, 64.0, 0.5, 0, False, gamma)
    assert math.isnan(loss_bad.item())
    loss_ok, _ = ops.margin_head(e, w, lab, 64.0, 0.5, 0, False, gamma)
    assert math.isfinite(loss_ok.item())


def test_cpu_tensors_are_rejected():
    from b200 import abi, ops
    with pytest.raises(abi.B200Error):
        ops.layernorm_fwd(torch.zeros(4, 96, dtype=bf16), torch.ones(96), torch.zeros(96))


def test_batched_reduce_matches_fp64_and_is_deterministic():
    """b200_reduce_defer_begin / b200_reduce_flush: several recorded reductions (aligned and not, 1..3 splits to 300) are
    folded by one launch; results match an fp64 sum and repeat bit for bit."""
    from b200 import abi, ops
    L = abi.lib()
    g = torch.Generator(device='cuda').manual_seed(5)
    shapes = [(7, 1000), (300, 4096), (1, 128), (148, 169), (33, 12)]
    parts = [torch.randn(s, n, device='cuda', generator=g) for s, n in shapes]

    def run():
        outs = [torch.full((n,), float('nan'), device='cuda') for _, n in shapes]
        abi.check(L.b200_reduce_defer_begin(), 'defer_begin')
        for p, o in zip(parts, outs):
            ops.splitk_reduce(p, out=o)
        assert L.b200_reduce_pending() == len(shapes)
        abi.check(L.b200_reduce_flush(abi.stream_ptr(), 0), 'flush')
        assert L.b200_reduce_pending() == 0
        torch.cuda.synchronize()
        return outs
    a, b = run(), run()
    for p, x, y in zip(parts, a, b):
        ref = p.double().sum(0)
        assert torch.equal(x, y)
        assert (x.double() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()) * p.shape[0] ** 0.5
    # outside a batch the call is immediate again
    o = ops.splitk_reduce(parts[0])
    torch.cuda.synchronize()
    assert torch.isfinite(o).all()
