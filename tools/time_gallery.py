"""Time the certified cosine top-k (filter + rerank) alone:  python tools/time_gallery.py [nq] [ng] [k]
B200_GALLERY_PAIR=0 selects the single-CTA filter for A/B runs."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'pets-face-recognition_b200')]
import torch
from b200 import gallery

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 125000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(1)
emb = torch.randn(ng, 512, device=dev, generator=g)
q = emb[:nq].contiguous() if nq <= ng else torch.randn(nq, 512, device=dev, generator=g)
gp = gallery.Prepared(emb, as_gallery=True)
qp = gallery.Prepared(q, as_gallery=False, frame_of=gp)
for _ in range(3):
    idx, score, unc = gallery.cosine_topk(q, emb, k, exclude_self_offset=0, q_prepared=qp, g_prepared=gp, return_uncertified=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    idx, score, unc = gallery.cosine_topk(q, emb, k, exclude_self_offset=0, q_prepared=qp, g_prepared=gp, return_uncertified=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print('pair=%s nq=%d ng=%d k=%d: %.3f ms, %.3f M q/s, %.1f TFLOP/s, uncertified %d, checksum %d' % (
    os.environ.get('B200_GALLERY_PAIR', '1'), nq, ng, k, ms, nq / ms / 1e3, 2.0 * nq * ng * 512 / ms / 1e9, int(unc), int(idx.sum())))
