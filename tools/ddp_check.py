"""N-GPU == 1-GPU check of the data-parallel training step (run under torchrun on the GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py

Every rank takes its slice of one global batch, runs forward/backward through engine.Trainer.run_training_batch's
all-reduce path, and rank 0 compares the averaged gradients and the updated parameters with a single-GPU run over the whole
batch (rows are independent and the loss is a mean, so they must agree to fp32 summation-order noise)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / 'pets-face-recognition_b200')]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build(device):
    from b200 import synth
    from losses import SoftmaxBasedMetricLearning
    from models import swin_t
    model = swin_t(num_classes=512)
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=123)
    model.load_state_dict(sd)
    wrap = SoftmaxBasedMetricLearning(model, num_class=1000, embedding_size=512, is_focal=True, arc_margin=True)
    wrap.add_margin.weight.data.copy_(synth.synth_tensor('add_margin.weight', (1000, 512), seed=123))
    return wrap.to(device)


class Mod(torch.nn.Module):
    def __init__(self, wrap):
        super().__init__()
        self.model_loss = wrap

    def training_step(self, batch, idx):
        return self.model_loss(batch['x'], batch['label'])['loss']


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl')
    from b200 import synth
    from engine.trainer import Trainer
    dev = torch.device('cuda', local)
    per = 4
    img = synth.synth_images(per * world, seed=5).to(dev)
    lab = synth.synth_labels(per * world, 1000, seed=5).to(dev)

    mod = Mod(build(dev))
    opt = torch.optim.SGD([p for p in mod.parameters() if p.requires_grad], 5e-3, momentum=0.9)
    tr = Trainer(gpus=[local], strategy='ddp', max_epochs=1)
    tr._allreduce_hooks(mod)
    sl = slice(rank * per, (rank + 1) * per)
    loss = tr.run_training_batch(mod, {'x': img[sl], 'label': lab[sl]}, [opt])
    torch.cuda.synchronize()
    after = torch.cat([p.detach().flatten() for p in mod.parameters() if p.requires_grad])
    grads = torch.cat([p.grad.flatten() / world for p in mod.parameters() if p.grad is not None])

    # every rank must hold identical parameters after the step
    ref_after = after.clone()
    dist.broadcast(ref_after, 0)
    same = torch.equal(ref_after, after)
    flags = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)

    if rank == 0:
        single = Mod(build(dev))
        opt1 = torch.optim.SGD([p for p in single.parameters() if p.requires_grad], 5e-3, momentum=0.9)
        tr1 = Trainer(gpus=[local], max_epochs=1)
        loss1 = tr1.run_training_batch(single, {'x': img, 'label': lab}, [opt1])
        g1 = torch.cat([p.grad.flatten() for p in single.parameters() if p.grad is not None])
        a1 = torch.cat([p.detach().flatten() for p in single.parameters() if p.requires_grad])
        rel_g = ((grads - g1).norm() / g1.norm()).item()
        rel_p = ((after - a1).norm() / a1.norm()).item()
        print(f'world {world}: ranks identical after step: {bool(flags.item())}; grad rel-L2 vs 1 GPU {rel_g:.3e}; params rel-L2 {rel_p:.3e}; '
              f'loss(rank0 slice) {loss.item():.4f} loss(full) {loss1.item():.4f}')
        assert flags.item() == 1 and rel_g < 1e-3 and rel_p < 1e-6, (rel_g, rel_p)
        print('DDP CHECK OK')
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
