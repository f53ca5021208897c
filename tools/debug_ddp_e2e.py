"""Where does a data-parallel step spend its HOST time?  torchrun --nproc-per-node N tools/debug_ddp_e2e.py
Runs the bench model under the Trainer with B200_TRACE_STEP stamps, through three loops: device-resident batches (no host sync),
train_batches on host batches with a loss read per step, and the same without loss reads."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / 'pets-face-recognition_b200')]
os.environ['B200_TRACE_STEP'] = '1'

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import bench
    from b200 import synth
    from engine.trainer import Trainer
    dev = torch.device('cuda', local)
    B = 256
    wrap = bench.build_model(bench.NUM_CLASS, dev)
    module = bench._Module(wrap)
    opt = bench.make_optimizer(wrap)
    tr = Trainer(gpus=[local], strategy='ddp', max_epochs=1)
    tr._allreduce_hooks(module)
    devb = [{'x': synth.synth_images(B, seed=10 * rank + i).to(dev), 'label': synth.synth_labels(B, bench.NUM_CLASS, seed=10 * rank + i).to(dev)} for i in range(2)]
    hostb = [{'x': (synth.synth_images(B, seed=20 * rank + i) * 255).to(torch.uint8).pin_memory(), 'label': synth.synth_labels(B, bench.NUM_CLASS, seed=20 * rank + i).pin_memory()}
             for i in range(2)]

    def report(name, t0, n):
        torch.cuda.synchronize()
        dist.barrier()
        dt = (time.perf_counter() - t0) / n * 1e3
        rows = tr._trace[-n:]
        if rank == 0:
            print(f'== {name} ({tr.ddp_mode}): {dt:.2f} ms/step wall')
            for r in rows:
                d = [(b - a) * 1e3 for a, b in zip(r, r[1:])]
                print('   host ms: fwd %.2f  bwd %.2f  exchange %.2f  optimizer %.2f' % tuple(d))
        sys.stdout.flush()

    for i in range(4):
        tr.run_training_batch(module, devb[i % 2], [opt])
    torch.cuda.synchronize(); dist.barrier()
    n = 6
    t0 = time.perf_counter()
    for i in range(n):
        tr.run_training_batch(module, devb[i % 2], [opt])
    report('device batches, no host sync', t0, n)
    t0 = time.perf_counter()
    for i in range(n):
        tr.run_training_batch(module, devb[i % 2], [opt]).item()
    report('device batches, loss.item() every step', t0, n)
    tr.train_batches(module, (hostb[i % 2] for i in range(3)), [opt])
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    tr.train_batches(module, (hostb[i % 2] for i in range(n)), [opt], read_loss_every=1)
    report('train_batches on host batches, loss read per step', t0, n)
    t0 = time.perf_counter()
    tr.train_batches(module, (hostb[i % 2] for i in range(n)), [opt], read_loss_every=0)
    report('train_batches on host batches, no loss reads', t0, n)
    t0 = time.perf_counter()
    for i in range(n):
        b = {k: v.to(dev, non_blocking=True) for k, v in hostb[i % 2].items()}
        tr.run_training_batch(module, b, [opt])
    report('plain H2D on the compute stream, no prefetcher', t0, n)
    tr.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
