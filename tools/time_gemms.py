"""CUDA-event timing of every distinct tcgen05 GEMM launch of one training step (batch 256): time, algorithmic GB/s and TFLOP/s,
the launch's share of the step and its distance from max(HBM bound, tensor bound).   python tools/time_gemms.py [iters]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / 'pets-face-recognition_b200')]

import torch  # noqa: E402

from b200 import abi, ops  # noqa: E402

bf16 = torch.bfloat16


def main(iters=5, B=256):
    abi.require_device()
    pk = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()) if (ROOT / 'MEASURED_PEAKS.json').exists() else {'hbm_gbs': 6650.0, 'bf16_tflops_sustained': 1400.0}
    hbm, tf = pk['hbm_gbs'] * 1e9, pk['bf16_tflops_sustained'] * 1e12
    g = torch.Generator(device='cuda').manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, device='cuda', generator=g) * 0.3).to(bf16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    rows = []

    def t(name, count, fn, M, N, K, extra_bytes=0.0, out_bytes=2):
        fn()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot / iters * 1e3
        by = 2.0 * (M + N) * K + 1.0 * M * N * out_bytes + extra_bytes
        fl = 2.0 * M * N * K
        bound = max(by / hbm, fl / tf) * 1e6
        rows.append((name, count, us, by / us / 1e3, fl / us / 1e6, bound, count * (us - bound)))

    for s, blocks in enumerate((2, 2, 6, 2)):
        C, H = 96 << s, 56 >> s
        M = B * H * H
        x, x4 = rnd(M, C), rnd(M, 4 * C)
        wq, wo, w1, w2 = rnd(3 * C, C), rnd(C, C), rnd(4 * C, C), rnd(C, 4 * C)
        bias1, biasc = torch.randn(4 * C, device='cuda'), torch.randn(C, device='cuda')
        x3 = rnd(M, 3 * C)
        st = f's{s + 1} '
        t(st + 'qkv            [M,C]x[3C,C]', blocks, lambda: ops.gemm_tn(x, wq), M, 3 * C, C)
        t(st + 'out-proj+res   [M,C]x[C,C]', blocks, lambda: ops.gemm_tn(x, wo, bias=biasc, mode=abi.EPI_RESID, aux=x), M, C, C, extra_bytes=2.0 * M * C)
        t(st + 'fc1+gelu+gelu\' [M,C]x[4C,C]', blocks, lambda: ops.gemm_tn(x, w1, bias=bias1, mode=abi.EPI_GELU, want_grad=True), M, 4 * C, C, extra_bytes=2.0 * M * 4 * C)
        t(st + 'fc2+res        [M,4C]x[C,4C]', blocks, lambda: ops.gemm_tn(x4, w2, bias=biasc, mode=abi.EPI_RESID, aux=x), M, C, 4 * C, extra_bytes=2.0 * M * C)
        t(st + 'd fc2 * gelu\'  [M,C]x[4C,C]', blocks, lambda: ops.gemm_tn(x, w1, mode=abi.EPI_DGELU, aux=x4), M, 4 * C, C, extra_bytes=2.0 * M * 4 * C)
        t(st + 'd fc1          [M,4C]x[C,4C]', blocks, lambda: ops.gemm_tn(x4, w2), M, C, 4 * C)
        t(st + 'd out-proj     [M,C]x[C,C]', blocks, lambda: ops.gemm_tn(x, wo), M, C, C)
        t(st + 'd qkv          [M,3C]x[C,3C]', blocks, lambda: ops.gemm_tn(x3, rnd(C, 3 * C)), M, C, 3 * C)
        for nm, dy, xx in (('wgrad fc2 ', x, x4), ('wgrad fc1 ', x4, x), ('wgrad out ', x, x), ('wgrad qkv ', x3, x)):
            N_, K_ = dy.shape[1], xx.shape[1]
            bn = (K_ + 15) // 16 * 16 if K_ <= 256 else 0
            tiles = ((N_ + 127) // 128) * ((K_ + (bn or 256) - 1) // (bn or 256))
            splits = max(1, 148 // tiles)
            t(st + nm + f'    [{N_},{K_}] over M', blocks, lambda dy=dy, xx=xx, splits=splits, bn=bn: ops.gemm_wgrad(dy, xx, splits=splits, block_n=bn), N_, K_, M,
              out_bytes=4 * splits)
    tot_gap = sum(r[6] for r in rows)
    tot = sum(r[1] * r[2] for r in rows)
    print(f'{"launch":44s} {"x":>2s} {"us":>8s} {"GB/s":>7s} {"TF/s":>7s} {"bound us":>9s} {"gap us/step":>11s}')
    for r in sorted(rows, key=lambda r: -r[6]):
        print(f'{r[0]:44s} {r[1]:2d} {r[2]:8.1f} {r[3]:7.0f} {r[4]:7.0f} {r[5]:9.1f} {r[6]:11.1f}')
    print(f'total {tot / 1e3:.2f} ms per step in these launches, {tot_gap / 1e3:.2f} ms above their bounds')


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
