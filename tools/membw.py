"""Fill / copy bandwidth of this box with plain torch kernels: the write-stream ceiling the store-heavy GEMMs are compared with."""
import torch

n = 1 << 30   # 1 GiB
a = torch.empty(n, dtype=torch.uint8, device='cuda')
b = torch.empty(n, dtype=torch.uint8, device='cuda')
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.view(torch.float32).fill_(1.0)); print(f'fill  1 GiB: {ms:.3f} ms  {n / ms / 1e6:.0f} GB/s written')
ms = t(lambda: a.zero_()); print(f'zero  1 GiB: {ms:.3f} ms  {n / ms / 1e6:.0f} GB/s written')
ms = t(lambda: b.copy_(a)); print(f'copy  1 GiB: {ms:.3f} ms  {2 * n / ms / 1e6:.0f} GB/s read+write')
ms = t(lambda: a.view(torch.float32).sum()); print(f'read  1 GiB: {ms:.3f} ms  {n / ms / 1e6:.0f} GB/s read')
