"""Worker of tests/test_multigpu_gpu.py - one process per GPU under torchrun (NCCL).  Real kernels on every rank:

  1. data-parallel training steps: averaged gradients / updated parameters of N ranks == one GPU over the whole batch, and the
     peer-arena gradient exchange (copy-engine pushes + slot sum in the optimizer kernel) == the NCCL all-reduce path bit for bit;
  2. Trainer.fit under strategy='ddp': rank-sharded loader, rank 0 writes the checkpoint;
  3. gallery-sharded cosine top-k (BASELINE config 4 layout) == single-GPU pass, bit for bit;
  4. leave-one-out Recall@K with row-sharded embeddings == single-GPU value;
  5. the ResNet-50 FE through the same DDP trainer step == one GPU.
Rank 0 prints 'MGPU CHECK OK' when everything holds."""
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / 'pets-face-recognition_b200')]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build(device, num_class=1000):
    from b200 import synth
    from losses import SoftmaxBasedMetricLearning
    from models import swin_t
    model = swin_t(num_classes=512)
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=123)
    model.load_state_dict(sd)
    wrap = SoftmaxBasedMetricLearning(model, num_class=num_class, embedding_size=512, is_focal=True, arc_margin=True)
    wrap.add_margin.weight.data.copy_(synth.synth_tensor('add_margin.weight', (num_class, 512), seed=123))
    return wrap.to(device)


class Mod(torch.nn.Module):
    """The slice of engine.Controller the Trainer drives (training_step / dataloaders / optimizers)."""

    def __init__(self, wrap, dataset=None):
        super().__init__()
        self.model_loss = wrap
        self.dataset = dataset

    def training_step(self, batch, idx):
        return self.model_loss(batch['x'], batch['label'])['loss']

    def train_dataloader(self):
        torch.manual_seed(123)                   # as the shipped configs: the same seed on every rank
        return torch.utils.data.DataLoader(self.dataset, batch_size=4, shuffle=True, drop_last=True)

    def val_dataloader(self):
        return []

    def validation_step(self, *a):
        return None

    def validation_epoch_end(self, outputs):
        pass

    def configure_optimizers(self):
        return [torch.optim.SGD([p for p in self.parameters() if p.requires_grad], 5e-3, momentum=0.9)], []


class Items(torch.utils.data.Dataset):
    def __init__(self, n):
        from b200 import synth
        self.x, self.y = synth.synth_images(n, seed=9), synth.synth_labels(n, 1000, seed=9)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return {'x': self.x[i], 'label': self.y[i], 'index': i}


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl')
    from b200 import gallery, synth
    from engine.trainer import Trainer
    dev = torch.device('cuda', local)

    # ---- 1. DDP steps == single GPU; the peer-arena exchange (default) == the NCCL all-reduce path, bit for bit
    per, n_steps = 16, 3
    img = synth.synth_images(per * world * n_steps, seed=5).to(dev)
    lab = synth.synth_labels(per * world * n_steps, 1000, seed=5).to(dev)

    def ddp_run(mode, steps):
        os.environ['B200_DDP'] = mode
        mod = Mod(build(dev))
        opt = torch.optim.SGD([p for p in mod.parameters() if p.requires_grad], 5e-3, momentum=0.9)
        tr = Trainer(gpus=[local], strategy='ddp', max_epochs=1)
        tr._allreduce_hooks(mod)
        assert tr.ddp_mode == mode, (tr.ddp_mode, mode)
        for s in range(steps):
            lo = (s * world + rank) * per
            tr.run_training_batch(mod, {'x': img[lo:lo + per], 'label': lab[lo:lo + per]}, [opt])
        torch.cuda.synchronize()
        after = torch.cat([p.detach().flatten() for p in mod.parameters() if p.requires_grad])
        mom = torch.cat([opt.state[p]['momentum_buffer'].flatten() for p in mod.parameters() if p.requires_grad and p.grad is not None])
        tr.close()
        return after, mom

    def same_on_all_ranks(t):
        ref = t.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([int(torch.equal(ref, t))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    after, grads = ddp_run('p2p', 1)               # first step, no weight decay: the momentum buffer IS the averaged gradient
    after_nccl, grads_nccl = ddp_run('nccl', 1)
    identical = same_on_all_ranks(after)
    assert torch.equal(after, after_nccl) and torch.equal(grads, grads_nccl), 'peer-arena step differs from the NCCL all-reduce step'
    after3, _ = ddp_run('p2p', n_steps)            # three steps: both arena buffers, momentum, the bf16 weight cache refresh
    after3_nccl, _ = ddp_run('nccl', n_steps)
    identical3 = same_on_all_ranks(after3)
    if world == 2:                                 # two addends commute: the slot-order sum equals NCCL's sum bit for bit
        assert torch.equal(after3, after3_nccl), 'peer-arena trajectory differs from the NCCL one'
    os.environ.pop('B200_DDP', None)
    if rank == 0:
        single = Mod(build(dev))
        opt1 = torch.optim.SGD([p for p in single.parameters() if p.requires_grad], 5e-3, momentum=0.9)
        tr1 = Trainer(gpus=[local], max_epochs=1)
        tr1.run_training_batch(single, {'x': img[:per * world], 'label': lab[:per * world]}, [opt1])
        g1 = torch.cat([p.grad.flatten() for p in single.parameters() if p.grad is not None])
        a1 = torch.cat([p.detach().flatten() for p in single.parameters() if p.requires_grad])
        rel_g = ((grads - g1).norm() / g1.norm()).item()
        rel_p = ((after - a1).norm() / a1.norm()).item()
        for s in range(1, n_steps):
            lo = s * world * per
            tr1.run_training_batch(single, {'x': img[lo:lo + per * world], 'label': lab[lo:lo + per * world]}, [opt1])
        a3 = torch.cat([p.detach().flatten() for p in single.parameters() if p.requires_grad])
        rel_p3 = ((after3 - a3).norm() / a3.norm()).item()
        print(f'world {world}: ranks identical {identical} / {identical3}; grad rel-L2 vs 1 GPU {rel_g:.3e}; params rel-L2 {rel_p:.3e} '
              f'(after {n_steps} steps {rel_p3:.3e}); peer arena == NCCL bit for bit')
        assert identical and identical3 and rel_g < 1e-3 and rel_p < 1e-6 and rel_p3 < 1e-4, (rel_g, rel_p, rel_p3)
        del single

    # ---- 2. fit(): sharded loader + rank-0 checkpoint
    root = Path(os.environ.get('MGPU_TMP', tempfile.gettempdir())) / 'mgpu_ckpt'
    fit_mod = Mod(build(dev), Items(8 * world))
    tr2 = Trainer(gpus=[local], strategy='ddp', max_epochs=1, default_root_dir=str(root), enable_checkpointing=True,
                  log_every_n_steps=0)
    seen = [int(i) for b in tr2._shard_loader(fit_mod.train_dataloader(), 0) for i in b['index']]
    all_seen = [None] * world
    dist.all_gather_object(all_seen, seen)
    flat = [i for s in all_seen for i in s]
    assert len(flat) == len(set(flat)) == 8 * world, all_seen            # disjoint shards covering the set
    tr2.fit(fit_mod)
    dist.barrier()
    if rank == 0:
        assert list(root.glob('epoch=0-step=*.ckpt')) and (root / 'trainer_state.pt').exists(), list(root.iterdir())

    # ---- 3. gallery-sharded top-k == single pass; 4. sharded leave-one-out Recall@K == single value
    emb, classes = synth.synth_embeddings(4000, 5, sigma=1.0, seed=31)      # 20,000 rows
    emb, classes = emb.to(dev), classes.to(dev)
    q, _ = synth.synth_embeddings(700, 3, sigma=1.0, seed=32)               # 2,100 queries
    q = q.to(dev)
    gb = [emb.shape[0] * r // world for r in range(world + 1)]
    gb[1] += 37 if world > 1 else 0                                          # ragged shards
    qb = [q.shape[0] * r // world for r in range(world + 1)]
    idx, score = gallery.cosine_topk_gallery_sharded(q[qb[rank]:qb[rank + 1]], emb[gb[rank]:gb[rank + 1]], 100)
    ref_idx, ref_score = gallery.cosine_topk(q, emb, 100)
    assert torch.equal(idx, ref_idx[qb[rank]:qb[rank + 1]]) and torch.equal(score, ref_score[qb[rank]:qb[rank + 1]])
    r_sh = gallery.recall_at_k_sharded(emb[gb[rank]:gb[rank + 1]], classes[gb[rank]:gb[rank + 1]], (10, 100))
    r_one = gallery.recall_at_k(emb, classes, (10, 100))
    assert r_sh == r_one, (r_sh, r_one)
    # ---- 5. ResNet-50 FE under DDP (generic path: every parameter's gradient all-reduced from a post-accumulate hook).  BatchNorm
    # statistics are per rank (as torch DDP without SyncBN), so every rank is fed the SAME batch: then the averaged gradients and
    # the updated parameters must equal the single-GPU step.
    from losses import SoftmaxBasedMetricLearning
    from models import resnet50

    def build_resnet():
        torch.manual_seed(77)
        net = resnet50()
        net.fc = torch.nn.Linear(2048, 512)
        return Mod(SoftmaxBasedMetricLearning(net, num_class=1000, embedding_size=512, is_focal=True, arc_margin=True).to(dev))

    rimg, rlab = synth.synth_images(8, seed=6).to(dev), synth.synth_labels(8, 1000, seed=6).to(dev)
    rmod = build_resnet()
    ropt = torch.optim.SGD([p for p in rmod.parameters() if p.requires_grad], 5e-3, momentum=0.9)
    tr5 = Trainer(gpus=[local], strategy='ddp', max_epochs=1)
    tr5._allreduce_hooks(rmod)
    tr5.run_training_batch(rmod, {'x': rimg, 'label': rlab}, [ropt])
    torch.cuda.synchronize()
    r_after = torch.cat([p.detach().flatten() for p in rmod.parameters()])
    if rank == 0:
        one = build_resnet()
        oopt = torch.optim.SGD([p for p in one.parameters() if p.requires_grad], 5e-3, momentum=0.9)
        Trainer(gpus=[local], max_epochs=1).run_training_batch(one, {'x': rimg, 'label': rlab}, [oopt])
        o_after = torch.cat([p.detach().flatten() for p in one.parameters()])
        rel = ((r_after - o_after).norm() / o_after.norm()).item()
        print(f'ResNet-50 DDP step on {world} ranks vs 1 GPU: params rel-L2 {rel:.3e}')
        assert rel < 1e-6, rel
    ok = torch.ones(1, device=dev)
    dist.all_reduce(ok)
    if rank == 0:
        print(f'sharded gallery == single pass on {world} ranks; Recall@K {r_sh}')
        print('MGPU CHECK OK')
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
