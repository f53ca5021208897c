import sys
sys.path[:0] = ['/root/repo', '/root/repo/pets-face-recognition_b200', '/root/repo/tests']
import torch
from b200 import gallery
from test_gallery_gpu import _spec_fp64_gpu
for spread in (0.02, 0.002, 0.0005):
    gen = torch.Generator().manual_seed(11)
    common = torch.nn.functional.normalize(torch.randn(1, 512, generator=gen))
    g = torch.nn.functional.normalize(common + spread * torch.randn(25600, 512, generator=gen) / 512 ** 0.5).cuda()
    q = g[:2000].contiguous()
    idx, score, unc = gallery.cosine_topk(q, g, 100, exclude_self_offset=0, return_uncertified=True)
    n_unc = int(unc)
    ref_i, ref_s = _spec_fp64_gpu(q, g, 100, 0)
    bad = (idx.long() != ref_i).any(1)
    # which queries were re-done
    qp = None
    print(f'spread {spread}: uncertified {n_unc}; mismatching queries {int(bad.sum())}; mean cos {(g[:500] @ g[500:1000].t()).mean().item():.8f}')
    if bad.any():
        b = bad.nonzero().flatten()[0].item()
        d = (idx[b].long() != ref_i[b]).nonzero().flatten()
        print('  query', b, 'first diff at rank', d[0].item(), 'kernel', idx[b, d[0]].item(), score[b, d[0]].item(), 'ref', ref_i[b, d[0]].item(), ref_s[b, d[0]].item(),
              'n diffs', len(d), 'set equal', set(idx[b].tolist()) == set(ref_i[b].tolist()))
        # recompute both candidates' scores in plain fp64 per-row
        for cand in (idx[b, d[0]].item(), ref_i[b, d[0]].item()):
            qq, gg = q[b].double(), g[cand].double()
            print('   cand', cand, 'fp64 cos (torch dot)', (qq @ gg / (qq.norm() * gg.norm())).item())
