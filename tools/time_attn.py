"""CUDA-event timing of the window-attention kernels at the bench's shapes (batch 256, all four stages):
   python tools/time_attn.py [iters]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / 'pets-face-recognition_b200')]

import torch  # noqa: E402

from b200 import abi, ops  # noqa: E402


def main(iters=10, B=256):
    abi.require_device()
    g = torch.Generator(device='cuda').manual_seed(0)
    tot_f = tot_b = 0.0
    for stage, blocks in enumerate((2, 2, 6, 2)):
        C, H = 96 << stage, 56 >> stage
        M, heads = B * H * H, (96 << stage) // 32
        qkv = (torch.randn(M, 3 * C, device='cuda', generator=g) * 0.5).to(torch.bfloat16)
        dout = (torch.randn(M, C, device='cuda', generator=g) * 0.5).to(torch.bfloat16)
        pos = torch.randn(13, 13, device='cuda')
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        out, lse = ops.window_attn_fwd(qkv, pos, B, H, H, C, heads, 1, window_major=True)
        ops.window_attn_bwd(qkv, pos, lse, dout, B, H, H, C, heads, 1, window_major=True)
        tf = tb = 0.0
        for _ in range(iters):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out, lse = ops.window_attn_fwd(qkv, pos, B, H, H, C, heads, 1, window_major=True)
            e[1].record()
            ops.window_attn_bwd(qkv, pos, lse, dout, B, H, H, C, heads, 1, window_major=True)
            e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        tf, tb = tf / iters * 1e3, tb / iters * 1e3
        bytes_f, bytes_b = 4 * M * C * 2, 9 * M * C * 2
        print(f'stage {stage + 1}: fwd {tf:7.1f} us ({bytes_f / tf / 1e3:6.0f} GB/s)  bwd {tb:7.1f} us ({bytes_b / tb / 1e3:6.0f} GB/s)   x{blocks} blocks')
        tot_f += tf * blocks; tot_b += tb * blocks
    print(f'per train step: fwd {tot_f / 1e3:.3f} ms, bwd {tot_b / 1e3:.3f} ms')


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
