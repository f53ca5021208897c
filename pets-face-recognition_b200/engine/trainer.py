"""Lightning-free Trainer with the reference's surface (engine/trainer.py:64-129 takes pl.Trainer's keyword list;
the ones the reference actually passes - utils/__init__.py:122-134 - are honoured, the rest are accepted and ignored).

Behaviours re-created from PL 1.5.9 as the reference relies on them:
  * fit: configure_optimizers() -> ([optim], [sched]); per batch H2D -> training_step -> zero_grad -> backward ->
    optimizer step; schedulers stepped once per epoch; validation every epoch in eval()/no_grad;
  * *_epoch_end(outputs) always gets outputs[dataloader_idx][batch_idx] (engine/loops/eval_loop.py:30-37);
  * num_sanity_val_steps = 0, enable_checkpointing via plain state_dict files (the format the reference's released
    checkpoints use, eval_fe_dog_head_sgd.py:18-21).
B200 specifics: the optimizer arithmetic is one fused multi-tensor kernel (b200/optim.py) driven by the config's own
torch.optim object; host batches are staged through pinned memory on a copy stream one step ahead; with
strategy='ddp' (one process per GPU, torchrun-style env) every rank pushes its gradients, per stage bucket and as soon
as each stage's backward has been enqueued, into the other ranks' peer arenas with the copy engines (NVLink / NVSwitch,
b200/peer.py), and the optimizer kernel sums the ranks' copies and folds in the 1/world factor (B200_DDP=nccl: the same
buckets through NCCL all-reduces on a side stream).
"""
from __future__ import annotations

import os
import time
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional

import torch

from b200.optim import FusedStep


def _to_device(batch, device, non_blocking=True):
    if torch.is_tensor(batch):
        return batch.to(device, non_blocking=non_blocking)
    if isinstance(batch, dict):
        return {k: _to_device(v, device, non_blocking) for k, v in batch.items()}
    if isinstance(batch, (list, tuple)):
        return type(batch)(_to_device(v, device, non_blocking) for v in batch)
    return batch


def _pin(batch):
    if torch.is_tensor(batch):
        return batch if (batch.is_cuda or batch.is_pinned()) else batch.pin_memory()
    if isinstance(batch, dict):
        return {k: _pin(v) for k, v in batch.items()}
    if isinstance(batch, (list, tuple)):
        return type(batch)(_pin(v) for v in batch)
    return batch


def _outside_grad_ready(param) -> None:
    """post-accumulate hook of a parameter outside the native engines (DDP only): hand the finished gradient to the Trainer."""
    owner = getattr(param, '_b200_ddp_owner', None)
    if owner is not None and param.grad is not None:
        owner._on_outside_grad(param)


class _Prefetcher:
    """Yields device batches; batch i+1 is copied host->device on a side stream while step i computes.

    reuse=True (training loops, which do not keep a batch after its step): the device copies live in three persistent
    slots guarded by events instead of fresh allocations per batch - no caching-allocator traffic (a 154 MB block per step
    that cannot be recycled until the compute stream has passed it) and no cudaMalloc inside a steady-state step."""

    def __init__(self, batches: Iterable, device, reuse: bool = False):
        self.it = iter(batches)
        self.device = device
        self.stream = torch.cuda.Stream(device=device) if device.type == 'cuda' else None
        self.reuse = reuse and self.stream is not None
        self._slots = [None, None, None]
        self._slot_free = [None, None, None]
        self._count = 0
        self._prev_slot = None
        self._next = None
        self._preload()

    def _into_slot(self, slot, host):
        """Copy `host` into slot's persistent device tensors (re-created when the batch structure changes)."""
        def alloc(h):
            if torch.is_tensor(h):
                return torch.empty(h.shape, dtype=h.dtype, device=self.device)
            if isinstance(h, dict):
                return {k: alloc(v) for k, v in h.items()}
            if isinstance(h, (list, tuple)):
                return type(h)(alloc(v) for v in h)
            return h

        def same(d, h):
            if torch.is_tensor(h):
                return torch.is_tensor(d) and d.shape == h.shape and d.dtype == h.dtype
            if isinstance(h, dict):
                return isinstance(d, dict) and d.keys() == h.keys() and all(same(d[k], h[k]) for k in h)
            if isinstance(h, (list, tuple)):
                return isinstance(d, type(h)) and len(d) == len(h) and all(same(a, b) for a, b in zip(d, h))
            return True

        def copy(d, h):
            if torch.is_tensor(h):
                d.copy_(h, non_blocking=True)
                return d
            if isinstance(h, dict):
                return {k: copy(d[k], h[k]) for k in h}
            if isinstance(h, (list, tuple)):
                return type(h)(copy(a, b) for a, b in zip(d, h))
            return h
        if self._slots[slot] is None or not same(self._slots[slot], host):
            self._slots[slot] = alloc(host)
        return copy(self._slots[slot], host)

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        if self.stream is None:
            self._next = _to_device(host, self.device, False)
            return
        host = _pin(host)
        if not self.reuse:
            with torch.cuda.stream(self.stream):
                self._next = (_to_device(host, self.device, True), host, None)
            return
        slot = self._count % 3
        self._count += 1
        with torch.cuda.stream(self.stream):
            if self._slot_free[slot] is not None:
                self.stream.wait_event(self._slot_free[slot])     # the step that last read this slot has finished
            self._next = (self._into_slot(slot, host), host, slot)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        if self.stream is None:
            out = self._next
        else:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(self.stream)
            out, _host, slot = self._next
            if self.reuse:
                if self._prev_slot is not None:                   # everything that reads the previous batch is enqueued by now
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    self._slot_free[self._prev_slot] = ev
                self._prev_slot = slot
            else:
                for t in (out.values() if isinstance(out, dict) else [out]):
                    if torch.is_tensor(t):
                        t.record_stream(cur)
        self._preload()
        return out


class Trainer:
    def __init__(self, gpus=None, default_root_dir=None, strategy=None, max_epochs=None, logger=False,
                 enable_checkpointing=False, callbacks=None, precision='bf16', num_sanity_val_steps=0,
                 check_val_every_n_epoch=1, limit_train_batches=None, limit_val_batches=None, benchmark=False,
                 log_every_n_steps=50, **ignored):
        self.gpus = gpus
        self.default_root_dir = default_root_dir
        self.strategy = strategy
        self.max_epochs = max_epochs if max_epochs is not None else 1
        self.logger = logger if logger not in (False, None) else None
        self.enable_checkpointing = enable_checkpointing
        self.callbacks = callbacks or []
        self.precision = precision            # compute is always bf16 storage / fp32 accumulate on this path
        self.check_val_every_n_epoch = check_val_every_n_epoch
        self.limit_train_batches, self.limit_val_batches = limit_train_batches, limit_val_batches
        self.log_every_n_steps = log_every_n_steps
        self.ignored_kwargs = dict(ignored)
        self.current_epoch = 0
        self.global_step = 0
        self.world_size, self.rank = 1, 0
        self._comm_stream = None
        self.ddp_mode, self._peer_state = 'nccl', None
        # B200_TRACE_STEP=1: host time stamps per step [start, forward enqueued, backward enqueued, exchange closed, optimizer enqueued]
        self._trace = [] if os.environ.get('B200_TRACE_STEP') else None
        self._early_reduced = set()
        self._fused: Dict[int, FusedStep] = {}
        self.device = self._pick_device()
        if strategy is not None:
            self._init_distributed()

    # ------------------------------------------------------------------ setup
    def _pick_device(self) -> torch.device:
        if not self.gpus:
            return torch.device('cpu')
        if 'LOCAL_RANK' in os.environ and self.strategy is not None:
            return torch.device('cuda', int(os.environ['LOCAL_RANK']))
        g = self.gpus
        idx = g[0] if isinstance(g, (list, tuple)) else 0
        return torch.device('cuda', int(idx))

    def _init_distributed(self):
        import torch.distributed as dist
        if not dist.is_initialized():
            if 'RANK' not in os.environ:
                raise RuntimeError("strategy='ddp' expects one process per GPU launched torchrun-style "
                                   '(RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment)')
            if self.device.type == 'cuda':
                torch.cuda.set_device(self.device)
            dist.init_process_group('nccl' if self.device.type == 'cuda' else 'gloo')
        self.world_size, self.rank = dist.get_world_size(), dist.get_rank()

    # ------------------------------------------------------------------ the hot step
    def _allreduce_hooks(self, module):
        """Install the data-parallel gradient exchange on `module` (idempotent per Trainer).

        Native Swin engines (default, B200_DDP=p2p): per-stage buckets of the flat gradient - and the parameters outside the
        engines, e.g. the ArcFace weight, from post-accumulate hooks: their gradient is the FIRST thing backward produces -
        are pushed into every rank's peer arena by the copy engines as soon as they are final (b200/peer.py), and the
        optimizer kernel sums the ranks' slots.  B200_DDP=nccl, CPU / gloo runs and models without a native engine (the
        ResNet path: hundreds of autograd-owned gradient tensors) all-reduce with torch.distributed instead: per stage bucket
        / per parameter on a side stream, launched as early."""
        if self.world_size == 1:
            return []
        import torch.distributed as dist
        engines = [m.engine for m in module.modules() if hasattr(m, '_engine') and hasattr(m, 'engine')]
        if self._comm_stream is None and self.device.type == 'cuda':
            self._comm_stream = torch.cuda.Stream(device=self.device)
        covered = set()
        for eng in engines:
            covered.update(id(p) for p in eng.params)
        outside = [p for p in module.parameters() if p.requires_grad and id(p) not in covered]
        use_p2p = bool(engines) and self.device.type == 'cuda' and os.environ.get('B200_DDP', 'p2p') != 'nccl'
        self.ddp_mode = 'p2p' if use_p2p else 'nccl'
        self._peer_state = {'engines': engines, 'outside': outside, 'source': None} if use_p2p else None

        for k, eng in enumerate(engines):
            def hook(stage, flat_grad, eng=eng, k=k):
                begin, end = eng.stage_param_range(stage)
                if self.ddp_mode == 'p2p':
                    src = self._peer_source()
                    src.exchange.push(flat_grad[begin:end], src.offsets[('engine', k)] + begin)
                else:
                    self._reduce_async(flat_grad[begin:end])
            eng.grad_hook = hook
        self._early_reduced = set()
        for prm in outside:
            if getattr(prm, '_b200_ddp_owner', None) is None:
                prm.register_post_accumulate_grad_hook(_outside_grad_ready)
            prm._b200_ddp_owner = self           # the hook serves whichever Trainer installed itself last
        return engines

    def _reduce_async(self, t) -> None:
        import torch.distributed as dist
        if self._comm_stream is None:
            dist.all_reduce(t)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            dist.all_reduce(t)

    def _on_outside_grad(self, param) -> None:
        if self.ddp_mode == 'p2p':
            src = self._peer_source()
            src.exchange.push(param.grad, src.offsets[id(param)])
        else:
            self._reduce_async(param.grad)
        self._early_reduced.add(id(param))

    def _peer_source(self):
        """The peer arena of this trainer's module, built at the first gradient push (the flat layouts of the native engines
        exist once a forward has run).  Collective: every rank gets here at the same point of its first backward."""
        st = self._peer_state
        if st is None:
            return None
        if st['source'] is None:
            from b200 import peer
            segments = [(('engine', k), eng.total) for k, eng in enumerate(st['engines'])] + [(id(p), p.numel()) for p in st['outside']]
            offsets, total = peer.build_layout(segments)
            for k, eng in enumerate(st['engines']):
                base = offsets[('engine', k)]
                for prm, o in zip(eng.params, eng.offsets):
                    offsets[id(prm)] = base + o
            try:
                st['source'] = peer.ArenaGradSource(peer.PeerGradExchange(total, self.device), offsets)
            except peer.B200Error as e:
                # decided collectively inside PeerGradExchange: every rank lands here together.  NCCL is a transport of equal
                # standing (the round-1 path), not a degraded compute path - but say so.
                import sys
                if self.rank == 0:
                    sys.stderr.write(f'[b200] DDP: {e}; gradients go through NCCL all-reduce instead\n')
                self.ddp_mode, self._peer_state = 'nccl', None
                return None
        return st['source']

    def close(self) -> None:
        """Release the peer arena (collective under DDP).  Optional: process exit frees it as well."""
        st = getattr(self, '_peer_state', None)
        if st is not None and st['source'] is not None:
            st['source'].exchange.close()
            st['source'] = None

    def run_training_batch(self, module, batch, optimizers) -> torch.Tensor:
        """One optimizer step == PL's optimizer.step(closure): training_step -> zero_grad -> backward -> step."""
        trace = self._trace
        if trace is not None:
            trace.append([time.perf_counter()])
        for opt in optimizers:
            opt.zero_grad(set_to_none=True)
        loss = module.training_step(batch, self.global_step)
        if self.ddp_mode == 'p2p':
            self._peer_source()               # first step: build the arena here, on the calling thread, not inside backward
        if trace is not None:
            trace[-1].append(time.perf_counter())
        loss.backward()
        if trace is not None:
            trace[-1].append(time.perf_counter())
        grad_src = None
        if self.world_size > 1:
            import torch.distributed as dist
            covered = set()
            for m in module.modules():
                if hasattr(m, '_engine') and m._engine is not None and m._engine.last_flat_grad is not None:
                    covered.update(id(p) for p in m._engine.params)
            early = getattr(self, '_early_reduced', set())
            rest = [p for p in module.parameters() if p.grad is not None and id(p) not in covered and id(p) not in early]
            self._early_reduced = set()
            if self.ddp_mode == 'p2p':
                grad_src = self._peer_source()
                for p in rest:
                    grad_src.exchange.push(p.grad, grad_src.offsets[id(p)])
                grad_src.shift = grad_src.exchange.finish()
            elif self._comm_stream is not None:
                ev = torch.cuda.Event(); ev.record()
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(ev)
                    for p in rest:
                        dist.all_reduce(p.grad)
                torch.cuda.current_stream().wait_stream(self._comm_stream)
            else:
                for p in rest:
                    dist.all_reduce(p.grad)
        if trace is not None:
            trace[-1].append(time.perf_counter())
        for opt in optimizers:
            fused = self._fused.get(id(opt))
            if fused is None:
                fused = self._fused[id(opt)] = FusedStep(opt)
            fused.step(grad_scale=1.0 / self.world_size, grad_src=grad_src)
        if trace is not None:
            trace[-1].append(time.perf_counter())
        self.global_step += 1
        return loss.detach()

    def train_batches(self, module, host_batches: Iterable, optimizers, read_loss_every: int = 1) -> List[float]:
        """Run optimizer steps over an iterable of HOST batches (dicts of CPU tensors): pinned staging + async H2D one
        step ahead, and a device->host read of the loss every `read_loss_every` steps."""
        losses = []
        if self.device.type != 'cuda':
            for i, batch in enumerate(_Prefetcher(host_batches, self.device)):
                loss = self.run_training_batch(module, batch, optimizers)
                if read_loss_every and (i + 1) % read_loss_every == 0:
                    losses.append(float(loss.item()))
            return losses
        # The loss of step i crosses to the host through a pinned buffer and is read after step i + 1 has been enqueued,
        # so the read never leaves the GPU without queued work (a blocking .item() per step would idle it for the whole
        # launch time of the next step).
        bufs = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        pending = None
        n_read = 0
        for i, batch in enumerate(_Prefetcher(host_batches, self.device, reuse=True)):
            loss = self.run_training_batch(module, batch, optimizers)
            if read_loss_every and (i + 1) % read_loss_every == 0:
                buf = bufs[n_read % 2]
                n_read += 1
                buf.copy_(loss.detach().reshape(1).float(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[0].synchronize()
                    losses.append(float(pending[1][0]))
                pending = (ev, buf)
        if pending is not None:
            pending[0].synchronize()
            losses.append(float(pending[1][0]))
        return losses

    # ------------------------------------------------------------------ loops
    def _unpack_optimizers(self, module):
        cfg = module.configure_optimizers()
        if isinstance(cfg, (list, tuple)) and len(cfg) == 2 and isinstance(cfg[0], (list, tuple)):
            return list(cfg[0]), list(cfg[1])
        if isinstance(cfg, (list, tuple)):
            return list(cfg), []
        return [cfg], []

    def _shard_eval_loader(self, loader):
        """Evaluation under DDP: rank r takes samples r, r + world, ... of the loader's dataset (no padding, no duplicates -
        a repeated sample would count twice in Recall@K); Controller.*_epoch_end gathers the shards again."""
        if self.world_size == 1:
            return loader
        from torch.utils.data import DataLoader, Sampler

        class _Strided(Sampler):
            def __init__(self, n, rank, world):
                self.idx = list(range(rank, n, world))

            def __iter__(self):
                return iter(self.idx)

            def __len__(self):
                return len(self.idx)
        if not isinstance(loader, DataLoader) or loader.batch_size is None:
            return loader
        kw = dict(batch_size=loader.batch_size, sampler=_Strided(len(loader.dataset), self.rank, self.world_size),
                  num_workers=loader.num_workers, collate_fn=loader.collate_fn, pin_memory=loader.pin_memory, drop_last=False,
                  timeout=loader.timeout, worker_init_fn=loader.worker_init_fn)
        if loader.num_workers > 0:
            kw.update(prefetch_factor=loader.prefetch_factor, persistent_workers=loader.persistent_workers)
        return DataLoader(loader.dataset, **kw)

    def _run_eval(self, module, dataloaders, step_name: str, limit=None):
        if not isinstance(dataloaders, (list, tuple)):
            dataloaders = [dataloaders]
        dataloaders = [self._shard_eval_loader(dl) for dl in dataloaders]
        was_training = module.training
        module.eval()
        outputs = []
        with torch.no_grad():
            for dl_idx, dl in enumerate(dataloaders):
                dl_out = []
                for b_idx, batch in enumerate(_Prefetcher(dl, self.device)):
                    if limit is not None and b_idx >= limit:
                        break
                    dl_out.append(getattr(module, step_name)(batch, b_idx, dl_idx))
                outputs.append(dl_out)               # ALWAYS [dataloader][batch]
        module.train(was_training)
        return outputs

    def _shard_loader(self, loader, epoch: int):
        """PL's replace_sampler_ddp: under DDP every rank must see its own shard of the training set.
        A DataLoader without a distributed sampler is rebuilt around DistributedSampler(shuffle = the loader's own shuffle);
        set_epoch(epoch) reseeds the shuffle every epoch.  Anything that is not a DataLoader is used as it is."""
        if self.world_size == 1:
            return loader
        from torch.utils.data import DataLoader, RandomSampler
        from torch.utils.data.distributed import DistributedSampler
        if not isinstance(loader, DataLoader):
            return loader
        sampler = getattr(loader, 'sampler', None)
        if isinstance(sampler, DistributedSampler):
            sampler.set_epoch(epoch)
            return loader
        if loader.batch_sampler is not None and loader.batch_size is None:
            raise RuntimeError("strategy='ddp': a train DataLoader with a custom batch_sampler must shard itself by rank")
        dist_sampler = DistributedSampler(loader.dataset, num_replicas=self.world_size, rank=self.rank,
                                          shuffle=isinstance(sampler, RandomSampler), drop_last=loader.drop_last)
        dist_sampler.set_epoch(epoch)
        kw = dict(batch_size=loader.batch_size, sampler=dist_sampler, num_workers=loader.num_workers, collate_fn=loader.collate_fn,
                  pin_memory=loader.pin_memory, drop_last=loader.drop_last, timeout=loader.timeout,
                  worker_init_fn=loader.worker_init_fn)
        if loader.num_workers > 0:
            kw.update(prefetch_factor=loader.prefetch_factor, persistent_workers=loader.persistent_workers)
        return DataLoader(loader.dataset, **kw)

    def fit(self, module) -> None:
        module.to(self.device)
        module.logger = self.logger
        optimizers, schedulers = self._unpack_optimizers(module)
        self._allreduce_hooks(module)
        module.train()
        for epoch in range(self.current_epoch, self.max_epochs):
            self.current_epoch = module.current_epoch = epoch
            t0, n_img, last = time.time(), 0, None
            for b_idx, batch in enumerate(_Prefetcher(self._shard_loader(module.train_dataloader(), epoch), self.device, reuse=True)):
                if self.limit_train_batches is not None and b_idx >= self.limit_train_batches:
                    break
                last = self.run_training_batch(module, batch, optimizers)
                n_img += int(batch['x'].shape[0]) if isinstance(batch, dict) and 'x' in batch else 0
                if self.log_every_n_steps and (b_idx + 1) % self.log_every_n_steps == 0 and self.rank == 0:
                    print(f'epoch {epoch} step {b_idx + 1} loss {last.item():.4f}')
            for s in schedulers:
                s.step()
            if self.rank == 0 and last is not None:
                dt = time.time() - t0
                print(f'epoch {epoch}: loss {last.item():.4f}  {n_img / max(dt, 1e-9):.1f} img/s')
                if self.logger is not None and hasattr(self.logger, 'log_metrics'):
                    self.logger.log_metrics({'train_loss': last.item()}, epoch)
            if (epoch + 1) % self.check_val_every_n_epoch == 0:
                # every rank extracts its shard of the validation set; the Controller gathers and ranks (rank 0 prints)
                outputs = self._run_eval(module, module.val_dataloader(), 'validation_step', self.limit_val_batches)
                module.validation_epoch_end(outputs)
            if self.world_size > 1:
                import torch.distributed as dist
                dist.barrier()                        # engine/loops/train_loop.py:16-17
            if self.enable_checkpointing and self.default_root_dir is not None and self.rank == 0:
                ck = Path(self.default_root_dir)
                ck.mkdir(parents=True, exist_ok=True)
                torch.save(module.state_dict(), ck / f'epoch={epoch}-step={self.global_step}.ckpt')
                torch.save({'optimizers': [o.state_dict() for o in optimizers], 'schedulers': [s.state_dict() for s in schedulers],
                            'epoch': epoch + 1, 'global_step': self.global_step}, ck / 'trainer_state.pt')

    def resume(self, module, optimizers, schedulers, path) -> None:
        st = torch.load(Path(path) / 'trainer_state.pt')
        for o, s in zip(optimizers, st['optimizers']):
            o.load_state_dict(s)
        for o, s in zip(schedulers, st['schedulers']):
            o.load_state_dict(s)
        self.current_epoch, self.global_step = st['epoch'], st['global_step']

    def validate(self, module):
        module.to(self.device)
        outputs = self._run_eval(module, module.val_dataloader(), 'validation_step', self.limit_val_batches)
        module.validation_epoch_end(outputs)
        return getattr(module, 'last_metrics', None)

    def test(self, module):
        module.to(self.device)
        outputs = self._run_eval(module, module.test_dataloader(), 'test_step')
        module.test_epoch_end(outputs)
        return getattr(module, 'last_metrics', None)

    def predict(self, module):
        module.to(self.device)
        return self._run_eval(module, module.predict_dataloader(), 'test_step')
