import sys
sys.path[:0] = ['/root/repo', '/root/repo/pets-face-recognition_b200', '/root/repo/tests']
import torch
from b200 import gallery
from bench import build_model
dev = torch.device('cuda')
wrap = build_model(1000, dev).eval()
g = torch.Generator(device=dev).manual_seed(500)
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 12800
imgs = torch.empty(n_img, 3, 224, 224, device=dev, dtype=torch.uint8)
for lo in range(0, n_img // 2, 128):
    base = torch.nn.functional.interpolate(torch.rand(128, 3, 7, 7, device=dev, generator=g), size=224, mode='bilinear')
    both = (base.repeat_interleave(2, 0) * 0.8 + 0.2 * torch.rand(256, 3, 224, 224, device=dev, generator=g)).clamp_(0, 1)
    imgs[2 * lo:2 * (lo + 128)] = (both * 255).to(torch.uint8)
with torch.no_grad():
    emb = torch.cat([wrap(imgs[lo:lo + 256]) for lo in range(0, n_img, 256)])
q = emb[:2000].contiguous()
gp = gallery.Prepared(emb, as_gallery=True)
qp = gallery.Prepared(q, as_gallery=False, frame_of=gp)
idx, score, unc = gallery.cosine_topk(q, emb, 100, exclude_self_offset=0, q_prepared=qp, g_prepared=gp, return_uncertified=True)
print('uncertified', int(unc), 'of 2000; stats [G1, Gr, dG1, dGr]', gp.stats.tolist(), 'frame |mu|', gp.frame[-1].item())
print('q err mean [dQ1, dQr, Q1, Qr]', qp.err.mean(0).tolist(), 'max', qp.err.max(0).values.tolist())
approx = (qp.rows.float() @ gp.rows.float().t()) / gp.scale          # [2000, n]
approx[torch.arange(2000), torch.arange(2000)] = -1e9
exact = torch.nn.functional.normalize(q.double()) @ torch.nn.functional.normalize(emb.double()).t()
qmu = torch.nn.functional.normalize(q.double()) @ gp.frame[:512].double()
t = exact - qmu[:, None]
t[torch.arange(2000), torch.arange(2000)] = -1e9
print('max |approx - t|', (approx.double() - t)[t > -1e8].abs().max().item())
a_sorted = approx.sort(1, descending=True).values
t_sorted = t.sort(1, descending=True).values
gap = t_sorted[:, 99] - a_sorted[:, 128].double()
G1, Gr, dG1, dGr = gp.stats.double().tolist()
E = qp.err[:, 0].double() * G1 + qp.err[:, 1].double() * Gr + qp.err[:, 2].double() * dG1 + qp.err[:, 3].double() * dGr
print('gap(exact 100th - approx 129th): median %.3e min %.3e; E median %.3e max %.3e; would certify: %d' % (gap.median(), gap.min(), E.median(), E.max(), int((gap > E).sum())))
for cut in (128, 256, 512, 1024, 2048):
    gap_c = t_sorted[:, 99] - a_sorted[:, cut].double()
    print('cut at approx rank %d: would certify %d of 2000' % (cut + 1, int((gap_c > E).sum())))
print('score std among gallery for a query', t[0][t[0] > -1e8].std().item())
