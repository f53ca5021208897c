"""Drive single kernels at the bench's stage shapes (batch 256) for ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/<name> python tools/prof_kernels.py <what>

what: attn_bwd | attn_fwd | gemm_fc1 | gemm_qkv | gemm_fc2 | gemm_outproj | gemm_wgrad | gemm_wgrad_small | ln | gallery | gallery_bench | gallery_big
Each op runs 3 times (the first launches warm the caches / instruction memory); capture with -s to skip warm-ups.
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / 'pets-face-recognition_b200')]

import torch  # noqa: E402

from b200 import abi, ops  # noqa: E402

bf16 = torch.bfloat16


def main(what, stage=0, B=256):
    abi.require_device()
    C = 96 << stage
    H = 56 >> stage
    M = B * H * H
    heads = C // 32
    g = torch.Generator(device='cuda').manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, device='cuda', generator=g) * 0.5).to(bf16)
    if what in ('attn_fwd', 'attn_bwd'):
        qkv, pos = rnd(M, 3 * C), torch.randn(13, 13, device='cuda')
        for _ in range(3):
            out, lse = ops.window_attn_fwd(qkv, pos, B, H, H, C, heads, 1)
        if what == 'attn_bwd':
            dout = rnd(M, C)
            for _ in range(3):
                ops.window_attn_bwd(qkv, pos, lse, dout, B, H, H, C, heads, 1)
    elif what == 'gemm_fc1':
        x, w, b = rnd(M, C), rnd(4 * C, C), torch.randn(4 * C, device='cuda')
        for _ in range(3):
            ops.gemm_tn(x, w, bias=b, mode=abi.EPI_GELU, want_grad=True)
    elif what == 'gemm_qkv':
        x, w = rnd(M, C), rnd(3 * C, C)
        for _ in range(3):
            ops.gemm_tn(x, w)
    elif what == 'gemm_dqkv':          # data gradient of the qkv projection: K = 3C, N = C
        x, w = rnd(M, 3 * C), rnd(C, 3 * C)
        for _ in range(3):
            ops.gemm_tn(x, w)
    elif what == 'gemm_fc2':
        x, w, b, r = rnd(M, 4 * C), rnd(C, 4 * C), torch.randn(C, device='cuda'), rnd(M, C)
        for _ in range(3):
            ops.gemm_tn(x, w, bias=b, mode=abi.EPI_RESID, aux=r)
    elif what == 'gemm_outproj':       # K = N = C with the residual aux: the smallest-K layer shape
        x, w, b, r = rnd(M, C), rnd(C, C), torch.randn(C, device='cuda'), rnd(M, C)
        for _ in range(3):
            ops.gemm_tn(x, w, bias=b, mode=abi.EPI_RESID, aux=r)
    elif what == 'gemm_wgrad_small':   # out-projection weight gradient: one 128 x C output tile, split over the tokens
        dy, x = rnd(M, C), rnd(M, C)
        for _ in range(3):
            ops.splitk_reduce(ops.gemm_wgrad(dy, x, splits=147))
    elif what == 'gemm_wgrad':
        dy, x = rnd(M, 4 * C), rnd(M, C)
        for _ in range(3):
            ops.splitk_reduce(ops.gemm_wgrad(dy, x, splits=49))
    elif what == 'ln':
        x, gm, bt = rnd(M, C), torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
        for _ in range(3):
            y, mean, rstd = ops.layernorm_fwd(x, gm, bt)
            ops.layernorm_bwd(y, x, gm, mean, rstd, dres=y)
    elif what == 'time_ln':          # event-timed LayerNorm backward (with the residual-gradient add) at the four stage shapes
        for st in range(4):
            Cs, Ms = 96 << st, B * (56 >> st) ** 2
            x, dy, dr = rnd(Ms, Cs), rnd(Ms, Cs), rnd(Ms, Cs)
            gm, bt = torch.ones(Cs, device='cuda'), torch.zeros(Cs, device='cuda')
            y, mean, rstd = ops.layernorm_fwd(x, gm, bt)
            fn = lambda: ops.layernorm_bwd(dy, x, gm, mean, rstd, dres=dr, want_dres_colsum=True)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print(f'LN bwd C={Cs:4d} M={Ms:7d}: {us:7.1f} us  {4 * Ms * Cs * 2 / us / 1e3:7.0f} GB/s')
    elif what == 'gallery':
        from b200 import gallery
        q = torch.randn(8192, 512, device='cuda', generator=g)
        gal = torch.randn(262144, 512, device='cuda', generator=g)
        for _ in range(2):
            gallery.cosine_topk(q, gal, 100)
    elif what == 'gallery_bench':    # the bench leg's shape per GPU through the certified path (witness pass, filter, re-rank)
        from b200 import gallery
        gal = torch.randn(125000, 512, device='cuda', generator=g)
        q = gal[:50000].contiguous()
        gp = gallery.Prepared(gal, as_gallery=True)
        qp = gallery.Prepared(q, as_gallery=False, frame_of=gp)
        for _ in range(2):
            gallery.cosine_topk(q, gal, 100, exclude_self_offset=0, q_prepared=qp, g_prepared=gp)
    elif what == 'gallery_big':      # the bench leg's size per GPU, timed with events (not for use under ncu)
        from b200 import gallery
        q = torch.randn(50000, 512, device='cuda', generator=g)
        gal = torch.randn(125000, 512, device='cuda', generator=g)
        for _ in range(2):
            gallery.cosine_topk(q, gal, 100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gallery.cosine_topk(q, gal, 100)
        e1.record()
        torch.cuda.synchronize()
        print('gallery_big ms/call', e0.elapsed_time(e1) / 5)
    elif what == 'time_gemms':       # event-timed GEMM shapes of stages 1-2 (for A/B runs with B200_BOX_COLS / B200_WAIT_NS)
        def timed(name, fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f'{name:28s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us')
        for st in (0, 1):
            Cs, Ms = 96 << st, B * (56 >> st) ** 2
            x, x4, r = rnd(Ms, Cs), rnd(Ms, 4 * Cs), rnd(Ms, Cs)
            wq, wo, w1, w2 = rnd(3 * Cs, Cs), rnd(Cs, Cs), rnd(4 * Cs, Cs), rnd(Cs, 4 * Cs)
            b1, bo = torch.randn(4 * Cs, device='cuda'), torch.randn(Cs, device='cuda')
            dq = rnd(Ms, 3 * Cs)
            wqt = rnd(Cs, 3 * Cs)
            timed(f's{st + 1} qkv fwd   N={3 * Cs} K={Cs}', lambda: ops.gemm_tn(x, wq))
            timed(f's{st + 1} out-proj  N={Cs} K={Cs}', lambda: ops.gemm_tn(x, wo, bias=bo, mode=abi.EPI_RESID, aux=r))
            timed(f's{st + 1} out dgrad N={Cs} K={Cs}', lambda: ops.gemm_tn(x, wo))
            timed(f's{st + 1} fc1 dual  N={4 * Cs} K={Cs}', lambda: ops.gemm_tn(x, w1, bias=b1, mode=abi.EPI_GELU, want_grad=True))
            timed(f's{st + 1} fc2       N={Cs} K={4 * Cs}', lambda: ops.gemm_tn(x4, w2, bias=bo, mode=abi.EPI_RESID, aux=r))
            timed(f's{st + 1} qkv dgrad N={Cs} K={3 * Cs}', lambda: ops.gemm_tn(dq, wqt))
            del x, x4, r, dq
    elif what == 'time_bn':          # tile-width sweep for the small-K layers (more, narrower tiles = deeper A ring)
        def timed(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps * 1e3
        for st in ((0, 1) if stage == 0 else (stage,)):      # `time_bn 2` / `time_bn 3`: stage 3 / 4 shapes only
            Cs, Ms = 96 << st, B * (56 >> st) ** 2
            x, x4, r = rnd(Ms, Cs), rnd(Ms, 4 * Cs), rnd(Ms, Cs)
            wq, wo, w1, w2 = rnd(3 * Cs, Cs), rnd(Cs, Cs), rnd(4 * Cs, Cs), rnd(Cs, 4 * Cs)
            b1, bo = torch.randn(4 * Cs, device='cuda'), torch.randn(Cs, device='cuda')
            dq, wqt, w1t = rnd(Ms, 3 * Cs), rnd(Cs, 3 * Cs), rnd(Cs, 4 * Cs)
            for bn in (0, 48, 64, 96, 128, 192, 256):
                row = [f's{st + 1} bn={bn:3d}']
                for name, fn in (('qkv', lambda: ops.gemm_tn(x, wq, block_n=bn)),
                                 ('fc1', lambda: ops.gemm_tn(x, w1, bias=b1, mode=abi.EPI_GELU, want_grad=True, block_n=bn)),
                                 ('outp', lambda: ops.gemm_tn(x, wo, bias=bo, mode=abi.EPI_RESID, aux=r, block_n=bn)),
                                 ('fc2', lambda: ops.gemm_tn(x4, w2, bias=bo, mode=abi.EPI_RESID, aux=r, block_n=bn)),
                                 ('fc1dg', lambda: ops.gemm_tn(x4, w1t, block_n=bn)),
                                 ('qkvdg', lambda: ops.gemm_tn(dq, wqt, block_n=bn))):
                    try:
                        row.append(f'{name} {timed(fn):7.1f}')
                    except Exception as e:           # tile width not valid for this shape
                        row.append(f'{name}     n/a')
                print('  '.join(row), flush=True)
            del x, x4, r, dq
    torch.cuda.synchronize()
    print('done', what)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
