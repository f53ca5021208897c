"""bench.py - the headline metric of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): dog-head FE training step - Swin-T (224x224, 512-d) + ArcFace head
(C = 10,000 classes, s = 64, m = 0.5) + mean CE + SGD(momentum 0.9, the config's three parameter groups),
per-GPU batch 256, bf16 storage / fp32 accumulate, synthetic images and seeded random-init weights.
N > 1 (torchrun, one rank per GPU): pure data parallel, weak scaling, NCCL all-reduce of the gradients.

A "step" = forward + backward + (all-reduce) + optimizer step over one batch.
  value  : whole-job images/s with the batch already resident in HBM (CUDA events, max over ranks).
  e2e    : the same step through the public API (engine.Trainer.train_batches on HOST batches: pinned staging,
           H2D of the images every step, D2H read of the loss every step).
  roofline: the tcgen05 GEMM kernel (all Linear fwd/dgrad/wgrad + the ArcFace cosine GEMM): algorithmic bytes and FLOPs
           of its launches / their CUDA-event durations, against MEASURED_PEAKS.json's HBM and sustained bf16 figures; the
           larger fraction names the bound.
  cpu_baseline: the oracle port (oracle/, plain PyTorch fp32) of the same step timed on this box's host cores.

--impl reference times that CPU port alone (the reference itself needs pytorch-lightning and its datasets, neither
installable offline; its arithmetic for this path is restated in oracle/ and pinned to it by tests/golden/).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / 'pets-face-recognition_b200'
for _p in (str(ROOT), str(PKG)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = 'FE train images/sec (Swin-T + ArcFace, 224x224, bf16)'
UNIT = 'images/s'
NUM_CLASS = 10000
FWD_GFLOP = 8.980            # Swin-T forward per image


def train_flops_per_image(num_class=NUM_CLASS):
    return 3 * FWD_GFLOP * 1e9 + 6 * 512 * num_class      # 26.97 GFLOP at C = 10k


def peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return d.get('bf16_tflops_sustained', d.get('bf16_tflops')), d.get('hbm_gbs'), 'measured (MEASURED_PEAKS.json, sustained)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained, 6.65 TB/s)'


def ncu_traffic():
    """DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture (profiles/*_ncu_full.json,
    written by tools/ncu_suite.sh + tools/ncu_read.py on the GPU box): the largest single GEMM launch of the step (stage-1 fc1)."""
    files = sorted((ROOT / 'profiles').glob('*_ncu_full.json'))
    if not files:
        return None
    try:
        k = json.loads(files[-1].read_text())['kernels']['gemm_fc1']
        return {'source': f'profiles/{files[-1].name}', 'launch': k['what'], 'dram_bytes': k['dram_bytes'],
                'algorithmic_bytes': k['algorithmic_bytes'], 'duration_us': k['duration_us']}
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------ CPU side
def oracle_train_setup(batch, num_class=NUM_CLASS):
    from b200 import synth
    from oracle.swin_oracle import SwinSpec, param_shapes
    spec = SwinSpec()
    sd = synth.synth_state_dict(param_shapes(spec), seed=123)
    sd = {k: v.requires_grad_(not k.endswith('_mask')) for k, v in sd.items()}
    w = synth.synth_tensor('add_margin.weight', (num_class, 512), seed=123).requires_grad_(True)
    img = synth.synth_images(batch, seed=123)
    label = synth.synth_labels(batch, num_class, seed=123)
    return spec, sd, w, img, label


def oracle_train_step(spec, sd, w, img, label, bufs):
    """One reference-semantics step on the CPU: losses/__init__.py:37-46 forward, autograd backward, SGD(momentum) with the
    groups of configs/dog_fe/fe_dogs_config.py:123-133."""
    from oracle import head_oracle
    from oracle.swin_oracle import swin_forward
    params = [p for p in sd.values() if p.requires_grad] + [w]
    for p in params:
        p.grad = None
    out = head_oracle.metric_learning_forward(lambda x: swin_forward(sd, x, spec), w, img, label)
    out['loss'].backward()
    with torch.no_grad():
        n = len(params) - 1
        new = head_oracle.sgd_momentum_step([p for p in params[:n]], [p.grad for p in params[:n]], bufs[:n], 5e-3, 0.9, 0.0)
        new += head_oracle.sgd_momentum_step([params[n]], [params[n].grad], bufs[n:], 1e-2, 0.9, 1e-4)
    return out['loss'].item(), new


def time_oracle(batch, warmup, steps):
    torch.set_num_threads(os.cpu_count() or 1)
    spec, sd, w, img, label = oracle_train_setup(batch)
    bufs = [None] * (sum(1 for p in sd.values() if p.requires_grad) + 1)
    for _ in range(warmup):
        _, bufs = oracle_train_step(spec, sd, w, img, label, bufs)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, bufs = oracle_train_step(spec, sd, w, img, label, bufs)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def time_oracle_extract(batch=32, steps=2):
    """the oracle port's eval forward (embedding extraction) on the host cores, bounded sample."""
    from oracle.swin_oracle import swin_forward
    torch.set_num_threads(os.cpu_count() or 1)
    spec, sd, _, img, _ = oracle_train_setup(batch)
    with torch.no_grad():
        swin_forward(sd, img, spec)
        t0 = time.perf_counter()
        for _ in range(steps):
            swin_forward(sd, img, spec)
        dt = time.perf_counter() - t0
    return {'value': batch * steps / dt, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{steps} eval forwards of batch {batch} after 1 warm-up, oracle/ port, fp32'}


def time_oracle_gpu_eager(device, batch=256, steps=3):
    """SURVEY.md 8(d) 'GPU-side bar': the same oracle port (plain PyTorch ops: cuBLAS / ATen kernels) on the B200 under
    torch.autocast(bfloat16) - what the reference's own modules would do on this GPU.  Part of the baseline leg; never on
    the product path."""
    spec, sd, w, img, label = oracle_train_setup(batch)
    sd = {k: v.detach().to(device).requires_grad_(v.requires_grad) for k, v in sd.items()}
    w = w.detach().to(device).requires_grad_(True)
    img, label = img.to(device), label.to(device)
    bufs = [None] * (sum(1 for p in sd.values() if p.requires_grad) + 1)

    def step(b):
        with torch.autocast('cuda', dtype=torch.bfloat16):
            return oracle_train_step(spec, sd, w, img, label, b)[1]
    bufs = step(bufs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        bufs = step(bufs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {'value': batch * steps / dt, 'unit': UNIT, 'sample': f'{steps} steps of batch {batch}, oracle/ port on the GPU (PyTorch eager, autocast bf16)'}


def time_reference_modules(batch, warmup, steps):
    """The UNMODIFIED reference modules (baseline/_ref: models/swin.py, losses/* copied from the reference by
    __graft_entry__.install_reference) through the reference's own public API on the host cores: swin_t ->
    SoftmaxBasedMetricLearning(arc_margin=True) forward, autograd backward, torch.optim.SGD with the groups of
    configs/dog_fe/fe_dogs_config.py:123-133.  Raises ImportError if the copy is not there."""
    ref = ROOT / 'baseline' / '_ref'
    if not (ref / 'models' / 'swin.py').exists():
        raise ImportError('baseline/_ref not installed (run __graft_entry__.build() where /root/reference exists)')
    for name in [m for m in sys.modules if m == 'models' or m.startswith('models.') or m == 'losses' or m.startswith('losses.')]:
        del sys.modules[name]
    sys.path.insert(0, str(ref))
    try:
        import losses as ref_losses
        import models as ref_models
        assert str(ref) in ref_models.__file__ and str(ref) in ref_losses.__file__, 'reference modules shadowed'
    finally:
        sys.path.remove(str(ref))
    from b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    model = ref_models.swin_t(num_classes=512)
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=123)
    model.load_state_dict(sd, strict=True)
    wrap = ref_losses.SoftmaxBasedMetricLearning(model, num_class=NUM_CLASS, embedding_size=512, is_focal=True, arc_margin=True)
    wrap.add_margin.weight.data.copy_(synth.synth_tensor('add_margin.weight', (NUM_CLASS, 512), seed=123))
    wrap.train()
    params1 = [p for i, p in wrap.module.named_parameters() if 'fc' not in i]
    params2 = [p for i, p in wrap.module.named_parameters() if 'fc' in i]
    optim = torch.optim.SGD([{'lr': 10 ** -2 / 2, 'params': params1}, {'lr': 10 ** -2, 'params': params2},
                             {'lr': 10 ** -2, 'params': wrap.add_margin.parameters(), 'weight_decay': 1 * (10 ** -4)}], 0.01, momentum=0.9)
    img, label = synth.synth_images(batch, seed=123), synth.synth_labels(batch, NUM_CLASS, seed=123)

    def step():
        optim.zero_grad()
        loss = wrap(img, label)['loss']
        loss.backward()
        optim.step()
        return loss.item()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        last = step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, last


def reference_arm(args, rank):
    if rank != 0:
        return
    batch = 8
    try:
        ips, sec, _ = time_reference_modules(batch, args.warmup, args.steps)
        kind, what = 'reference', 'the reference\'s own models/swin.py + losses/* (unmodified copy in baseline/_ref) + torch.optim.SGD'
    except Exception as e:          # no copy of the reference on this box: the oracle port (pinned to it by tests/golden) stands in
        ips, sec = time_oracle(batch, args.warmup, args.steps)
        kind, what = 'port', f'oracle/ port (baseline/_ref unavailable: {type(e).__name__}: {str(e)[:80]})'
    cores = torch.get_num_threads()
    line = {'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'dog-head FE train step, Swin-T + ArcFace(C={NUM_CLASS}) + SGD, CPU fp32, bounded sample: batch {batch} per step'},
            'cpu_baseline': {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': kind,
                             'sample': f'{args.steps} steps of batch {batch} after {args.warmup} warm-up, {what}, {cores} threads'},
            'e2e': {'value': ips, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------------ GPU side
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith('active') for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'power_w_max': max(pw) if pw else None, 'samples': len(self.rows)}


def build_model(num_class, device):
    from b200 import synth
    from losses import SoftmaxBasedMetricLearning
    from models import swin_t
    model = swin_t(num_classes=512)
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=123)
    model.load_state_dict(sd, strict=True)
    wrap = SoftmaxBasedMetricLearning(model, num_class=num_class, embedding_size=512, is_focal=True, arc_margin=True)
    wrap.add_margin.weight.data.copy_(synth.synth_tensor('add_margin.weight', (num_class, 512), seed=123))
    return wrap.to(device)


def make_optimizer(wrap):
    # configs/dog_fe/fe_dogs_config.py:123-133
    params1 = [p for i, p in wrap.module.named_parameters() if 'fc' not in i]
    params2 = [p for i, p in wrap.module.named_parameters() if 'fc' in i]
    d = [{'lr': 10 ** -2 / 2, 'params': params1}, {'lr': 10 ** -2, 'params': params2},
         {'lr': 10 ** -2, 'params': wrap.add_margin.parameters(), 'weight_decay': 1 * (10 ** -4)}]
    return torch.optim.SGD(d, 0.01, momentum=0.9)


class _Module(torch.nn.Module):
    """Minimal Controller-shaped module (engine/controller.py:27-29) around the loss wrapper for the e2e leg."""

    def __init__(self, wrap):
        super().__init__()
        self.model_loss = wrap

    def training_step(self, batch, batch_idx):
        return self.model_loss(batch['x'], batch['label'])['loss']


def gpu_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    from b200 import abi, synth
    from engine.trainer import Trainer
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    abi.require_device()
    L = abi.lib()
    B = args.batch

    wrap = build_model(NUM_CLASS, device)
    module = _Module(wrap)
    opt = make_optimizer(wrap)
    trainer = Trainer(gpus=[local_rank], strategy='ddp' if world > 1 else None, max_epochs=1)
    trainer._allreduce_hooks(module)

    # two distinct device-resident batches (154 MB of images each: larger than the 126 MB L2), rank-offset data seed
    dev_batches = [{'x': synth.synth_images(B, seed=1000 * rank + i).to(device), 'label': synth.synth_labels(B, NUM_CLASS, seed=1000 * rank + i).to(device)}
                   for i in range(2)]

    def step(i):
        return trainer.run_training_batch(module, dev_batches[i % 2], [opt])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    check = abi.check
    # timed region A: K steps, nothing but the step itself on the stream -> `value`
    check(L.b200_prof_begin(0), 'prof_begin')
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    gemm_ms, gemm_fl, gemm_n, total_n = C.c_double(), C.c_double(), C.c_longlong(), C.c_longlong()
    check(L.b200_prof_end(None, None, None, C.byref(total_n)), 'prof_end')
    # timed region B: the same K steps again with a CUDA-event pair around every tcgen05 GEMM launch -> `roofline`.
    # (Kept apart from region A because an event record between two kernels suspends their programmatic dependent launch
    # overlap: it would tax the number it is meant to explain.)
    check(L.b200_prof_begin(1), 'prof_begin')
    evb0, evb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evb0.record()
    for i in range(args.steps):
        step(i)
    evb1.record()
    barrier()
    ms_b = evb0.elapsed_time(evb1)
    check(L.b200_prof_end(C.byref(gemm_ms), C.byref(gemm_fl), C.byref(gemm_n), None), 'prof_end')
    gemm_by = C.c_double()
    check(L.b200_prof_gemm_bytes(C.byref(gemm_by)), 'prof_gemm_bytes')
    # the same timed region, by kind of kernel (include/b200_fe.h: B200_PROF_*)
    NK = 6
    k_ms, k_fl, k_by, k_n = (C.c_double * NK)(), (C.c_double * NK)(), (C.c_double * NK)(), (C.c_longlong * NK)()
    check(L.b200_prof_kernels(NK, k_ms, k_fl, k_by, k_n), 'prof_kernels')
    kind_names = ['gemm_tn_kernel (tcgen05)', 'window_attn_fwd_kernel (tcgen05 + TMA)', 'window_attn_bwd_kernel (tcgen05 + TMA)',
                  'layernorm_fwd_kernel', 'layernorm_bwd_kernel', 'reduce_batch_kernel']
    kinds = [{'kernel': kind_names[i], 'launches_per_step': k_n[i] / args.steps, 'ms_per_step': k_ms[i] / args.steps,
              'algorithmic_gb_per_step': k_by[i] / args.steps / 1e9, 'algorithmic_tflop_per_step': k_fl[i] / args.steps / 1e12,
              'achieved_gbs': (k_by[i] / (k_ms[i] * 1e-3) / 1e9) if k_ms[i] > 0 else None,
              'achieved_tflops': (k_fl[i] / (k_ms[i] * 1e-3) / 1e12) if k_ms[i] > 0 else None} for i in range(NK)]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * args.steps / (ms * 1e-3)
    final_loss = float(loss.item())

    # ---- end to end through the public API: host batches, pinned staging + H2D every step, loss D2H every step
    # the host batch is what a decoder hands over: uint8 pixels (38.5 MB per 256 images instead of 154 MB of floats);
    # ToTensor's / 255 is fused into the first gather kernel (models/swin.py takes either dtype)
    host_batches = [{'x': (synth.synth_images(B, seed=2000 * rank + i) * 255).to(torch.uint8), 'label': synth.synth_labels(B, NUM_CLASS, seed=2000 * rank + i)}
                    for i in range(2)]
    host_batches = [{k: v.pin_memory() for k, v in b.items()} for b in host_batches]
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())

    hb = host_batches[0]['x']
    dst = torch.empty_like(hb, device=device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        dst.copy_(hb, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbps = 3 * hb.numel() * hb.element_size() / (time.perf_counter() - t0) / 1e9
    del dst

    def host_iter(n):
        for i in range(n):
            yield host_batches[i % 2]
    trainer.train_batches(module, host_iter(max(3, args.warmup)), [opt])
    barrier()
    t0 = time.perf_counter()
    losses = trainer.train_batches(module, host_iter(args.steps), [opt], read_loss_every=1)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / t.item()

    gallery = gallery_leg(args, rank, world, device) if not args.no_gallery else None
    extract = extract_leg(args, wrap, dev_batches, world, device) if not args.no_gallery else None
    rows_f = rows_f_leg(device) if (world == 1 and not args.no_gallery) else None
    pipeline = pipeline_leg(args, wrap, rank, world, device) if not args.no_gallery else None
    resnet = resnet_leg(args, device, local_rank) if (world == 1 and not args.no_gallery) else None

    ddp_mode = trainer.ddp_mode
    trainer.close()            # collective: before the other ranks leave
    if rank != 0:
        return
    peak_tf, peak_hbm, peak_src = peaks()
    achieved_tf = gemm_fl.value / (gemm_ms.value * 1e-3) / 1e12 if gemm_ms.value > 0 else 0.0
    achieved_gbs = gemm_by.value / (gemm_ms.value * 1e-3) / 1e9 if gemm_ms.value > 0 else 0.0
    # the GEMM launches of this network are mostly small-K (96..768): in aggregate their algorithmic HBM time exceeds their
    # tensor-core time; both fractions are reported
    grad_exchange = None
    if world > 1:
        grad_exchange = ('copy-engine pushes into every rank\'s peer arena (CUDA IPC over NVLink) under the backward + slot sum inside the '
                         'optimizer kernel (b200/peer.py)') if ddp_mode == 'p2p' else 'per-stage NCCL all-reduce on a side stream'
    # the bounding roofline of the training step is the TENSOR one (the unfused HBM reading - every launch's
    # operands and outputs counted as algorithmic bytes - is the kinder of the two and is kept as `roofline.hbm`)
    hbm_bound = False
    traffic = ncu_traffic()
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic',
        'config': {'workload': f'configs[1]: dog-head FE training step, Swin-T 224x224 + ArcFace(C={NUM_CLASS}, s=64, m=0.5) + mean CE + '
                               f'SGD(momentum 0.9, 3 param groups), per-GPU batch {B}, global batch {B * world}',
                   'parallelism': f'dp{world}', 'grad_exchange': grad_exchange, 'l2': 'inputs larger than L2 (154 MB of images per batch, two batches alternated; '
                                                      'each step streams > 10 GB of activations)',
                   'flops_per_image': train_flops_per_image(), 'final_loss': final_loss},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                'api': 'engine.Trainer.train_batches(module, host_batches, optimizers)', 'h2d_pinned_gbps_this_box': h2d_gbps,
                'host_batch': 'uint8 pixels [B, 3, 224, 224] + int64 labels in pinned memory (ToTensor / 255 fused into the first kernel)'},
        'gpu_launches': int(total_n.value),
        'roofline': {'bound': 'hbm' if hbm_bound else 'tensor',
                     'kernel': 'gemm::gemm_tn_kernel (tcgen05, all Linear fwd/dgrad/wgrad + ArcFace cosine)',
                     'achieved': achieved_gbs if hbm_bound else achieved_tf, 'peak': peak_hbm if hbm_bound else peak_tf,
                     'unit': 'GB/s' if hbm_bound else 'TFLOP/s',
                     'frac': (achieved_gbs / peak_hbm if hbm_bound else achieved_tf / peak_tf) if peak_tf and peak_hbm else None,
                     'traffic': traffic.get('dram_bytes') if traffic else None, 'traffic_launch': traffic,
                     'hbm': {'achieved_gbs': achieved_gbs, 'peak_gbs': peak_hbm, 'frac': achieved_gbs / peak_hbm if peak_hbm else None,
                             'algorithmic_bytes_per_step': gemm_by.value / args.steps},
                     'tensor': {'achieved_tflops': achieved_tf, 'peak_tflops': peak_tf, 'frac': achieved_tf / peak_tf if peak_tf else None,
                                'algorithmic_flops_per_step': gemm_fl.value / args.steps},
                     'peak_source': peak_src, 'launches_timed': int(gemm_n.value),
                     'kernel_ms_per_step': gemm_ms.value / args.steps, 'step_share': gemm_ms.value / ms_b if ms_b else None,
                     'timed_region': f'{args.steps} further steps with per-launch events, everything on one stream - the weight-gradient side stream of the backward is off while launches are timed one by one ({ms_b / args.steps:.2f} ms/step)',
                     'whole_step_frac': value / world * train_flops_per_image() / (peak_tf * 1e12) if peak_tf else None},
        'clocks': clocks,
    }
    # every instrumented kernel of the step with its own live fractions of the measured peaks: the dominant-kernel choice
    # above can be re-derived from this table
    for kd in kinds:
        kd['hbm_frac'] = kd['achieved_gbs'] / peak_hbm if (kd['achieved_gbs'] and peak_hbm) else None
        kd['tensor_frac'] = kd['achieved_tflops'] / peak_tf if (kd['achieved_tflops'] and peak_tf) else None
        kd['step_share'] = kd['ms_per_step'] * args.steps / ms_b if ms_b else None
    line['roofline']['kernels'] = kinds
    line['roofline']['uninstrumented_ms_per_step'] = ms_b / args.steps - sum(k['ms_per_step'] for k in kinds)
    if extract is not None:
        extract['frac_of_peak'] = extract['tflops'] / peak_tf if peak_tf else None
        line['extract'] = extract
    if rows_f is not None:
        line['tsv_scoring'], line['pair_scoring'] = rows_f
    if gallery is not None:
        gallery['frac_of_peak'] = gallery['tflops'] / peak_tf if peak_tf else None
        line['gallery'] = gallery
    if pipeline is not None:
        line['pipeline'] = pipeline
    if resnet is not None:
        line['resnet50'] = resnet
    if world == 1 and not args.no_cpu_baseline:
        cb, steps_cb = 32, 2
        ips, _ = time_oracle(cb, 1, steps_cb)
        line['cpu_baseline'] = {'value': ips, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                                'sample': f'{steps_cb} steps of batch {cb} after 1 warm-up: same step (Swin-T + ArcFace(C={NUM_CLASS}) + SGD), '
                                          f'oracle/ port, fp32, {torch.get_num_threads()} threads'}
        if rows_f is not None:
            line['cpu_baseline'].update(rows_f_cpu())
        if extract is not None:
            try:
                line['cpu_baseline']['extract'] = time_oracle_extract()
            except Exception as e:      # a baseline, not the product
                line['cpu_baseline']['extract'] = {'unavailable': f'{type(e).__name__}: {str(e)[:160]}'}
        if gallery is not None:
            try:
                line['cpu_baseline']['gallery'] = gallery_cpu(rows=args.gallery_rows)
            except Exception as e:      # a baseline, not the product
                line['cpu_baseline']['gallery'] = {'unavailable': f'{type(e).__name__}: {str(e)[:160]}'}
        try:
            line['cpu_baseline']['gpu_eager'] = time_oracle_gpu_eager(torch.device('cuda', 0))
        except Exception as e:          # a baseline, not the product: report why it is missing and go on
            line['cpu_baseline']['gpu_eager'] = {'unavailable': f'{type(e).__name__}: {str(e)[:160]}'}
    emit(line)


def extract_leg(args, wrap, dev_batches, world, device):
    """Embedding extraction (BASELINE.json configs[4], first half): forward only, eval mode, the same per-GPU batch, images
    resident in HBM; value = whole-job images/s (every rank extracts its own shard, no collective)."""
    import torch.distributed as dist
    was_training = wrap.training
    wrap.eval()
    B = dev_batches[0]['x'].shape[0]
    reps = max(5, args.steps // 2)
    with torch.no_grad():
        for i in range(3):
            wrap(dev_batches[i % 2]['x'])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(reps):
            emb = wrap(dev_batches[i % 2]['x'])
        ev1.record()
        torch.cuda.synchronize()
    wrap.train(was_training)
    t = torch.tensor([ev0.elapsed_time(ev1) / reps], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    return {'metric': 'FE extract images/sec (Swin-T forward, eval, bf16)', 'value': world * B / (ms * 1e-3), 'unit': 'images/s',
            'ms_per_batch': ms, 'batch_per_gpu': B, 'tflops': B * FWD_GFLOP * 1e9 / (ms * 1e-3) / 1e12, 'finite': bool(torch.isfinite(emb).all().item())}


RESNET_FWD_GFLOP = 8.18          # torchvision resnet50 at 224 x 224: 4.09 GMAC per image (fc -> 512 included)


def resnet_leg(args, device, local_rank):
    """the FE backbone the reference's configs ship - torchvision's ResNet-50 with a 512-d fc
    (configs/dog_fe/fe_dogs_config.py:96-109) - through the same loss wrapper, SGD groups and Trainer step as the main arm, at
    the same per-GPU batch, uint8 images resident in HBM; then the eval-mode extraction.  N = 1 only."""
    from engine.trainer import Trainer
    from losses import SoftmaxBasedMetricLearning
    from models import resnet50
    torch.manual_seed(123)
    model = resnet50()
    model.fc = torch.nn.Linear(2048, 512)
    wrap = SoftmaxBasedMetricLearning(model, num_class=NUM_CLASS, embedding_size=512, is_focal=True, arc_margin=True).to(device)
    module = _Module(wrap)
    opt = make_optimizer(wrap)
    trainer = Trainer(gpus=[local_rank], strategy=None, max_epochs=1)
    B = args.batch
    g = torch.Generator(device=device).manual_seed(7)
    batches = [{'x': torch.randint(0, 256, (B, 3, 224, 224), device=device, dtype=torch.uint8, generator=g),
                'label': torch.randint(0, NUM_CLASS, (B,), device=device, generator=g)} for _ in range(2)]
    steps = max(5, args.steps // 2)
    for i in range(3):
        loss = trainer.run_training_batch(module, batches[i % 2], [opt])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        loss = trainer.run_training_batch(module, batches[i % 2], [opt])
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    wrap.eval()
    with torch.no_grad():
        for i in range(2):
            emb = wrap(batches[i % 2]['x'])
        torch.cuda.synchronize()
        ev0.record()
        for i in range(steps):
            emb = wrap(batches[i % 2]['x'])
        ev1.record()
        torch.cuda.synchronize()
    ms_e = ev0.elapsed_time(ev1) / steps
    out = {'metric': 'ResNet-50 FE (torchvision resnet50, fc -> 512) + ArcFace train step and eval extraction, bf16, synthetic uint8 images',
           'batch_per_gpu': B, 'train_images_per_s': B / (ms * 1e-3), 'train_ms_per_step': ms, 'train_tflops': 3 * RESNET_FWD_GFLOP * B / ms,
           'extract_images_per_s': B / (ms_e * 1e-3), 'extract_ms_per_batch': ms_e, 'extract_tflops': RESNET_FWD_GFLOP * B / ms_e,
           'final_loss': float(loss), 'finite': bool(torch.isfinite(emb).all().item())}
    del wrap, module, opt, trainer, batches, model
    torch.cuda.empty_cache()
    return out


def _topk_spec_fp64_gpu(q, g, k, self_rows=None, chunk=32768):
    """CHECKER (not product): oracle/rank_oracle.py:topk_spec evaluated in fp64 with torch on the GPU, gallery in chunks -
    score = <q, g> / (max(|q|, 1e-8) max(|g|, 1e-8)), order = (score desc, index asc); self_rows[i] is excluded for query i."""
    qd = q.double()
    qn = qd.norm(dim=1).clamp_min(1e-8)
    best_s = torch.zeros((q.shape[0], 0), dtype=torch.float64, device=q.device)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64, device=q.device)
    for lo in range(0, g.shape[0], chunk):
        gd = g[lo:lo + chunk].double()
        sc = (qd @ gd.t()) / (qn[:, None] * gd.norm(dim=1).clamp_min(1e-8)[None, :])
        idx = torch.arange(lo, lo + gd.shape[0], device=q.device).expand(q.shape[0], -1)
        if self_rows is not None:
            sc = sc.masked_fill(idx == self_rows[:, None], float('-inf'))
        sc, idx = torch.cat([best_s, sc], 1), torch.cat([best_i, idx], 1)
        sc, order = torch.sort(sc, dim=1, descending=True, stable=True)
        best_s, best_i = sc[:, :k].contiguous(), torch.gather(idx, 1, order[:, :k]).contiguous()
    return best_i


def pipeline_leg(args, wrap, rank, world, device):
    """BASELINE.json configs[4]: the eval_fe pipeline end to end on all GPUs - Swin-T embedding extraction of every rank's
    image shard, NCCL all-gather of the embeddings, leave-one-out candR@10 / candR@100 (engine/controller.py:77-91) with every
    rank ranking its own queries against the whole set, counts all-reduced - plus parity: the top-100 lists of a 2,000-query
    subsample bit-exact against an fp64 evaluation of the specification, and candR on that subsample equal.
    Synthetic identities: two images per identity = one low-frequency pattern + pixel noise (uint8, resident in HBM)."""
    import torch.distributed as dist
    from b200 import gallery
    n_img, B = args.pipeline_images, 256
    n_img -= n_img % B
    # seeded random-init weights, as BASELINE.json's configs ask (the `wrap` of the training legs has taken SGD steps on random
    # labels by now, which collapses its embeddings onto one point: cos 0.9999995 between any two images)
    wrap = build_model(NUM_CLASS, device)
    g = torch.Generator(device=device).manual_seed(500 + rank)
    n_id = n_img // 2
    imgs = torch.empty(n_img, 3, 224, 224, device=device, dtype=torch.uint8)
    for lo in range(0, n_id, 128):
        m = min(128, n_id - lo)
        base = torch.nn.functional.interpolate(torch.rand(m, 3, 7, 7, device=device, generator=g), size=224, mode='bilinear')
        both = (base.repeat_interleave(2, 0) * 0.8 + 0.2 * torch.rand(2 * m, 3, 224, 224, device=device, generator=g)).clamp_(0, 1)
        imgs[2 * lo:2 * (lo + m)] = (both * 255).to(torch.uint8)
    classes_local = (torch.arange(n_img, device=device) // 2) + rank * n_id
    was_training = wrap.training
    wrap.eval()

    def run():
        emb = torch.empty(n_img, 512, device=device, dtype=torch.float32)
        with torch.no_grad():
            for lo in range(0, n_img, B):
                emb[lo:lo + B] = wrap(imgs[lo:lo + B])
        ev_x.record()
        if world > 1:
            emb_all, sizes = gallery._all_gather_rows(emb)
            cls_all, _ = gallery._all_gather_rows(classes_local)
            start = sum(sizes[:rank])
        else:
            emb_all, cls_all, start = emb, classes_local, 0
        rows = torch.arange(start, start + n_img, device=device)
        recall = gallery.recall_at_k_rows(emb_all, cls_all, rows, (10, 100))
        return emb_all, cls_all, rows, recall
    ev0, ev_x, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    run()                                                    # warm-up (plans, workspaces)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0.record()
    emb_all, cls_all, rows, recall = run()
    ev1.record()
    torch.cuda.synchronize()
    wrap.train(was_training)
    t = torch.tensor([ev0.elapsed_time(ev_x), ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_extract, ms_total = t.tolist()
    # parity on a subsample of this rank's queries (checker: fp64 torch on the GPU; top-100 = integer work -> bit-exact)
    sub = rows[torch.randperm(n_img, device=device, generator=g)[:2000]]
    mask = torch.ones(emb_all.shape[0], dtype=torch.bool, device=device)
    # same leave-one-out as the pipeline: the query's own row is excluded (here by score masking in the checker and by
    # dropping it from the kernel's 101-candidate answer would change k, so the kernel is asked for the offset form)
    order = torch.cat([sub, (mask.index_fill(0, sub, False)).nonzero().flatten()])
    g_re = emb_all[order].contiguous()
    idx_k, _, unc = gallery.cosine_topk(g_re[:sub.numel()], g_re, 100, exclude_self_offset=0, return_uncertified=True)
    ref = _topk_spec_fp64_gpu(g_re[:sub.numel()], g_re, 100, self_rows=torch.arange(sub.numel(), device=device))
    exact = bool(torch.equal(idx_k.long(), ref))
    # where the lists differ: is it more than a tie at fp64 rounding level (two rows whose cosines agree to 1e-13 - their order
    # depends on the summation order of the fp64 dot product, which the specification does not fix)?
    diff = (idx_k.long() != ref)
    beyond = 0
    if diff.any():
        qi, pos = diff.nonzero(as_tuple=True)
        qd = g_re[qi].double()

        def cos64(rows):
            gd = g_re[rows].double()
            return (qd * gd).sum(1) / (qd.norm(dim=1) * gd.norm(dim=1))
        beyond = int(((cos64(idx_k.long()[qi, pos]) - cos64(ref[qi, pos])).abs() > 1e-13).sum().item())
    c_re = cls_all[order]
    hit_k = [(c_re[idx_k.long()[:, :k]] == c_re[:sub.numel(), None]).any(1).float().mean().item() for k in (10, 100)]
    hit_r = [(c_re[ref[:, :k]] == c_re[:sub.numel(), None]).any(1).float().mean().item() for k in (10, 100)]
    n_all = emb_all.shape[0]
    cosm = (torch.nn.functional.normalize(emb_all[:2000]) @ torch.nn.functional.normalize(emb_all[2000:4000]).t()).mean().item() if n_all >= 4000 else None
    return {'metric': 'eval_fe pipeline: Swin-T extract + all-gather + leave-one-out candR@10/100', 'images': n_all, 'images_per_gpu': n_img,
            'identities': n_all // 2, 'seconds': ms_total * 1e-3, 'extract_images_per_s': n_all / (ms_extract * 1e-3),
            'match_queries_per_s': n_all / ((ms_total - ms_extract) * 1e-3), 'images_per_s_end_to_end': n_all / (ms_total * 1e-3),
            'candR@10': recall['Recall@K=10'], 'candR@100': recall['Recall@K=100'], 'mean_cosine_between_images': cosm,
            'parity': {'checker': 'fp64 evaluation of oracle/rank_oracle.py:topk_spec on the GPU, 2,000 sampled queries x all rows',
                       'top100_indices_bit_exact': exact, 'differences_beyond_fp64_rounding_ties': beyond, 'candR@10_kernel_vs_checker': [hit_k[0], hit_r[0]],
                       'candR@100_kernel_vs_checker': [hit_k[1], hit_r[1]], 'queries_re_done_by_exact_scan': int(unc)}}


def _folder_db(n_sets, seed, n_ids, prefix, dim=512):
    """Synthetic folder database in the reference's layout (generate_tsv_to_reproduce2.py:31-52): 1..4 (1, dim) head vectors."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(n_ids, dim, generator=torch.Generator().manual_seed(4321))
    db = {}
    for s in range(n_sets):
        ident = s % n_ids
        nvec = int(torch.randint(1, 5, (1,), generator=g).item())
        db[f'{prefix}{s:06d}'] = {'head_vectors': [(centres[ident] + 0.6 * torch.randn(dim, generator=g)).reshape(1, dim) for _ in range(nvec)],
                                  'type': 1 + ident % 2}
    return db


def rows_f_leg(device):
    """SURVEY.md 8f-1 / 8f-2 through their public calls: the submission table for 2,000 enroll x 20,000 verify folders
    (b200.multivector.calc_scores, host dicts in, rows out) and the 20,000 verification-pair scores + AUROC of one
    Controller evaluation (b200.gallery.pair_similarity + engine.metrics on the device)."""
    from b200 import gallery, multivector
    from engine import metrics as M
    n_q, n_g = 2000, 20000
    dq, dg = _folder_db(n_q, 1, 5000, 'q'), _folder_db(n_g, 2, 5000, 'g')
    multivector.calc_scores(dq, dg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = multivector.calc_scores(dq, dg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tsv = {'metric': 'folder pairs scored/sec (multi-vector mean strategy + top-100 + TSV rows, host dicts in)', 'value': n_q * (n_g / 2) / dt,
           'unit': 'folder pairs/s', 'enroll': n_q, 'verify': n_g, 'seconds': dt, 'rows': len(rows)}
    g = torch.Generator().manual_seed(5)
    n, n_pairs = 20000, 20000
    emb = torch.randn(n, 512, generator=g).to(device)
    i1, i2 = torch.randint(0, n, (n_pairs,), generator=g), torch.randint(0, n, (n_pairs,), generator=g)
    labels = torch.randint(0, 2, (n_pairs,), generator=g).to(device)
    M.auroc(gallery.pair_similarity(emb, i1, i2), labels)          # warm-up (first use of the sort / scan kernels)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sc = gallery.pair_similarity(emb, i1, i2)
    auc = M.auroc(sc, labels)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    pair = {'metric': 'verification pairs/sec (scores + AUROC, device)', 'value': n_pairs / dt, 'unit': 'pairs/s', 'pairs': n_pairs,
            'seconds': dt, 'auroc': auc}
    return tsv, pair


def rows_f_cpu():
    """The same two pieces as the reference runs them, on a bounded sample: its per-folder-pair Python loop
    (oracle.tsv_oracle.calc_scores: 4 enroll x 2,000 verify folders) and similarity_f over a list of 20,000 tensor pairs."""
    from oracle import rank_oracle, tsv_oracle
    dq, dg = _folder_db(4, 1, 5000, 'q'), _folder_db(2000, 2, 5000, 'g')
    t0 = time.perf_counter()
    tsv_oracle.calc_scores(dq, dg)
    dt = time.perf_counter() - t0
    g = torch.Generator().manual_seed(5)
    emb = torch.randn(20000, 512, generator=g)
    i1, i2 = torch.randint(0, 20000, (20000,), generator=g), torch.randint(0, 20000, (20000,), generator=g)
    t1 = time.perf_counter()
    rank_oracle.similarity_f([(emb[a], emb[b]) for a, b in zip(i1.tolist(), i2.tolist())])
    dt2 = time.perf_counter() - t1
    return {'tsv_scoring': {'value': 4 * 1000 / dt, 'unit': 'folder pairs/s', 'sample': '4 enroll x 2,000 verify folders, reference loop (oracle port)'},
            'pair_scoring': {'value': 20000 / dt2, 'unit': 'pairs/s', 'sample': 'similarity_f over 20,000 gathered tensor pairs (oracle port)'}}


def gallery_cpu(rows=125000, queries=2000, loop_n=400):
    """the reference's gallery ranking on the host cores, on a bounded sample - its leave-one-out loop
    statement for statement (oracle.rank_oracle.recall_at_k_loop = engine/controller.py:77-91) on `loop_n` embeddings, and the
    vectorised fp32 restatement (normalize -> chunked matmul -> topk(100)) against a shard of the size the GPU leg uses."""
    from oracle import rank_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(11)
    emb = torch.nn.functional.normalize(torch.randn(loop_n, 512, generator=g))
    classes = torch.arange(loop_n) // 2
    t0 = time.perf_counter()
    rank_oracle.recall_at_k_loop(emb, classes, (10, 100))
    dt_loop = time.perf_counter() - t0
    gal = torch.randn(rows, 512, generator=g)
    q = torch.randn(queries, 512, generator=g)
    rank_oracle.gallery_match_vectorised(q[:64], gal, 100, chunk=64)                  # warm-up
    t1 = time.perf_counter()
    rank_oracle.gallery_match_vectorised(q, gal, 100, chunk=1000)
    dt_vec = time.perf_counter() - t1
    return {'value': queries / dt_vec, 'unit': 'queries/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{queries} queries x {rows} gallery rows x 512, vectorised fp32 restatement (F.normalize -> chunked matmul -> '
                      f'topk(100)), oracle/rank_oracle.py:gallery_match_vectorised',
            'reference_loop': {'value': loop_n / dt_loop, 'unit': 'queries/s', 'us_per_gallery_row': dt_loop / (loop_n * (loop_n - 1)) * 1e6,
                               'sample': f'leave-one-out loop of engine/controller.py:77-91, statement for statement, on {loop_n} embeddings '
                                         f'(similarity_f over a Python list of tensor pairs + full argsort per query)'}}


def gallery_leg(args, rank, world, device):
    """Second half of BASELINE.json's metric: gallery queries/s (configs[3] shape, bounded): every rank holds a gallery shard of
    1M / 8 = 125k x 512 rows (the per-GPU share of config 4) and matches all `--gallery-queries` (50k) queries against it with the
    fused cosine + top-100 kernel; value = queries/s against the FULL gallery of world * 125k rows (all ranks work in parallel on
    their shards; the merge of the partial lists is included for world > 1)."""
    import torch.distributed as dist
    from b200 import gallery
    g = torch.Generator(device=device).manual_seed(100 + rank)
    nq, ng = args.gallery_queries, args.gallery_rows
    gal = torch.nn.functional.normalize(torch.randn(ng, 512, device=device, generator=g))
    q = torch.nn.functional.normalize(torch.randn(nq, 512, device=device, generator=torch.Generator(device=device).manual_seed(7)))
    gprep = gallery.Prepared(gal, as_gallery=True)      # gallery-side preparation (centred fp16 rows) is done once per gallery

    def once():
        idx, score, unc = gallery.cosine_topk(q, gal, 100, g_index_base=rank * ng, g_prepared=gprep, return_uncertified=True)
        once.uncertified = unc
        if world > 1:
            idx_all = [torch.empty_like(idx) for _ in range(world)]
            sc_all = [torch.empty_like(score) for _ in range(world)]
            dist.all_gather(idx_all, idx)
            dist.all_gather(sc_all, score)
            lo, hi = rank * nq // world, (rank + 1) * nq // world
            idx, score = gallery.topk_merge(torch.stack([s[lo:hi] for s in sc_all]), torch.stack([i[lo:hi] for i in idx_all]), 100)
        return idx
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    reps = 2
    for _ in range(reps):
        once()
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / reps], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    flops = 2.0 * 512 * nq * ng            # per GPU
    return {'metric': 'gallery queries/sec (cosine + top-100, 512-d, fp16 tensor-core pass + exact fp64 re-rank)', 'value': nq / (ms * 1e-3),
            'unit': 'queries/s', 'queries': nq, 'gallery_rows_total': ng * world, 'gallery_rows_per_gpu': ng, 'ms': ms,
            'tflops': flops / (ms * 1e-3) / 1e12,
            'exactness': 'certificate on (centred fp16 rows, rigorous error bound vs pruning margin); queries re-done by the exact fp64 scan: '
                         f'{int(once.uncertified)} of {nq}'}


_RESULT_FD = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints (NCCL's version banner
    arrives on fd 1) has been diverted to stderr by main()."""
    data = (json.dumps(line) + '\n').encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gallery', action='store_true')
    ap.add_argument('--gallery-queries', type=int, default=50000)
    ap.add_argument('--gallery-rows', type=int, default=125000)
    ap.add_argument('--pipeline-images', type=int, default=25600, help='images per GPU of the config-5 leg (extract + candR@10/100)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        reference_arm(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        gpu_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
