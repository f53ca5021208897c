import sys
sys.path[:0] = ['/root/repo', '/root/repo/pets-face-recognition_b200', '/root/repo/tests']
import copy
import torch
from test_resnet_gpu import _pair, _rel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
layers = tuple(int(v) for v in sys.argv[2].split(',')) if len(sys.argv) > 2 else (3, 4, 6, 3)
ref, ours, img = _pair(B, seed=3, layers=layers)
if len(sys.argv) > 3:      # bf16-representable weights and inputs: removes the deterministic rounding both bf16 pipelines share
    with torch.no_grad():
        for q in ref.parameters():
            if q.dim() > 1:
                q.copy_(q.to(torch.bfloat16).float())
    ours.load_state_dict(ref.state_dict())
    img = (img.float() / 255).to(torch.bfloat16).float()
ref.train(); ours.train()
imgf = img.float() / 255 if img.dtype == torch.uint8 else img
ref2 = copy.deepcopy(ref)
dev = img.device
tgt = torch.randn(B, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(9))
want = ref(imgf)
(want * tgt).sum().backward()
with torch.autocast('cuda', dtype=torch.bfloat16):
    w2 = ref2(imgf)
(w2.float() * tgt).sum().backward()
from oracle.resnet_oracle import forward_emulated
ref3 = copy.deepcopy(ref2)
for q in ref3.parameters():
    q.grad = None
w3 = forward_emulated(ref3, imgf)
(w3 * tgt).sum().backward()
got = ours(img)
(got * tgt).sum().backward()
print('emb rel: ours', _rel(got, want), 'autocast', _rel(w2.float(), want))
names = [n for n, _ in ref.named_parameters()]
ro = [_rel(p.grad, q.grad) for p, q in zip(ours.parameters(), ref.parameters())]
ra = [_rel(p.grad, q.grad) for p, q in zip(ref2.parameters(), ref.parameters())]
for i in list(range(0, len(names), 12)) + list(range(len(names) - 14, len(names))):
    print(f'{names[i]:40s} ours {ro[i]:.4f}  autocast {ra[i]:.4f}')
rb = [_rel(p.grad, q.grad) for p, q in zip(ours.parameters(), ref2.parameters())]
print('ours vs autocast: emb', _rel(got, w2.float()), 'grads median', sorted(rb)[len(rb) // 2], 'max', max(rb))
rc = [_rel(p.grad, q.grad) for p, q in zip(ours.parameters(), ref3.parameters())]
print('ours vs emulated: emb', _rel(got, w3), 'grads median', sorted(rc)[len(rc) // 2], 'max', max(rc), names[rc.index(max(rc))])
print('median ours', sorted(ro)[len(ro) // 2], 'autocast', sorted(ra)[len(ra) // 2])
