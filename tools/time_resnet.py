"""ResNet-50 FE on the B200 kernels: train step and eval forward at a batch, next to PyTorch eager (cuDNN, bf16 autocast,
channels_last) on the same GPU.   python tools/time_resnet.py [batch] [--no-eager]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'pets-face-recognition_b200')]
import torch

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
dev = torch.device('cuda')


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from b200 import abi
    from b200.optim import FusedStep
    from losses import SoftmaxBasedMetricLearning
    from models import resnet50
    model = resnet50()
    model.fc = torch.nn.Linear(2048, 512)
    wrap = SoftmaxBasedMetricLearning(model, num_class=10000, embedding_size=512, is_focal=True, arc_margin=True).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.randint(0, 256, (B, 3, 224, 224), device=dev, dtype=torch.uint8, generator=g)
    label = torch.randint(0, 10000, (B,), device=dev, generator=g)
    params1 = [p for i, p in wrap.module.named_parameters() if 'fc' not in i]
    params2 = [p for i, p in wrap.module.named_parameters() if 'fc' in i]
    opt = torch.optim.SGD([{'lr': 0.005, 'params': params1}, {'lr': 0.01, 'params': params2},
                           {'lr': 0.01, 'params': wrap.add_margin.parameters(), 'weight_decay': 1e-4}], 0.01, momentum=0.9)
    fused = FusedStep(opt)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = wrap(img, label)['loss']
        loss.backward()
        fused.step()
        return loss

    wrap.train()
    L = abi.lib()
    if '--quick' in sys.argv:          # one warm-up and one step: for an ncu launch list
        timed(step, reps=1, warm=1)
        return
    ms = timed(step)
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    t_cpu = (time.perf_counter() - t0) / 3
    torch.cuda.synchronize()
    print(f'host time to ISSUE one step (no synchronisation): {t_cpu * 1e3:.2f} ms')
    if '--cprofile' in sys.argv:
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(3):
            step()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
    abi.check(L.b200_prof_begin(0), 'prof_begin')
    step()
    import ctypes as C
    n = C.c_longlong()
    abi.check(L.b200_prof_end(None, None, None, C.byref(n)), 'prof_end')
    print(f'b200 resnet50 train step  B={B}: {ms:8.2f} ms  {B / ms * 1e3:9.0f} img/s  {B * 24.5e9 / ms / 1e9:7.1f} TFLOP/s  ({n.value} library launches, loss {float(step()):.3f})')
    wrap.eval()
    with torch.no_grad():
        ms_e = timed(lambda: wrap(img))
    print(f'b200 resnet50 eval fwd    B={B}: {ms_e:8.2f} ms  {B / ms_e * 1e3:9.0f} img/s  {B * 8.2e9 / ms_e / 1e9:7.1f} TFLOP/s')
    if '--no-eager' in sys.argv:
        return
    import torchvision
    ref = torchvision.models.resnet50(weights=None)
    ref.fc = torch.nn.Linear(2048, 512)
    ref = ref.to(dev).to(memory_format=torch.channels_last)
    head = torch.nn.Linear(512, 10000, bias=False).to(dev)
    ropt = torch.optim.SGD(list(ref.parameters()) + list(head.parameters()), 0.01, momentum=0.9)
    x = (img.float() / 255).contiguous(memory_format=torch.channels_last)

    def eager_step():
        ropt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = head(ref(x))
        loss = torch.nn.functional.cross_entropy(out.float(), label)
        loss.backward()
        ropt.step()

    ref.train()
    ms_r = timed(eager_step)
    print(f'torch eager (cuDNN, bf16 autocast, channels_last) train step: {ms_r:8.2f} ms  {B / ms_r * 1e3:9.0f} img/s')
    ref.eval()
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        ms_re = timed(lambda: ref(x))
    print(f'torch eager eval fwd: {ms_re:8.2f} ms  {B / ms_re * 1e3:9.0f} img/s')


main()
