import sys
sys.path[:0] = ['/root/repo', '/root/repo/pets-face-recognition_b200']
import torch
from bench import build_model
wrap = build_model(1000, torch.device('cuda')).eval()
g = torch.Generator(device='cuda').manual_seed(0)
embs = []
with torch.no_grad():
    for i in range(4):
        x = torch.rand(256, 3, 224, 224, device='cuda', generator=g)
        embs.append(wrap(x))
    # structured: identity pattern (low-frequency) + noise
    base = torch.nn.functional.interpolate(torch.rand(64, 3, 7, 7, device='cuda', generator=g), size=224, mode='bilinear')
    x = (base.repeat_interleave(4, 0) * 0.8 + 0.2 * torch.rand(256, 3, 224, 224, device='cuda', generator=g)).clamp(0, 1)
    es = wrap(x)
e = torch.nn.functional.normalize(torch.cat(embs))
c = e @ e.t()
off = c[~torch.eye(len(e), dtype=torch.bool, device='cuda')]
print('uniform-noise images: cos mean %.6f min %.6f max %.6f std %.2e' % (off.mean(), off.min(), off.max(), off.std()))
e = torch.nn.functional.normalize(es)
c = e @ e.t()
same = torch.arange(256, device='cuda') // 4
m_same = (same[:, None] == same[None, :]) & ~torch.eye(256, dtype=torch.bool, device='cuda')
print('structured: same-id cos mean %.6f; diff-id cos mean %.6f min %.6f std %.2e' % (c[m_same].mean(), c[~m_same & ~torch.eye(256, dtype=torch.bool, device="cuda")].mean(), c.min(), c[~m_same].std()))
