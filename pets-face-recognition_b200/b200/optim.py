"""Fused multi-tensor optimizer step (csrc/elementwise.cu: sgd_kernel / adamw_kernel).

The reference's configs return stock ``torch.optim.SGD`` / ``torch.optim.AdamW`` objects from
``config.optimizer(model_loss)`` (configs/dog_fe/fe_dogs_config.py:123-133, body_dog_fe.py:123-131).  To stay a
drop-in, the trainer keeps those objects - their param_groups (lr, momentum, weight_decay; mutated by
MultiStepLR) and their ``state`` (so ``optimizer.state_dict()`` checkpoints keep the torch layout) - and
replaces only the arithmetic of ``optimizer.step()`` by ONE kernel launch over all parameters.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import abi, plan
from .abi import B200Error, check, lib, stream_ptr


class FusedStep:
    def __init__(self, optimizer: torch.optim.Optimizer):
        if isinstance(optimizer, torch.optim.AdamW):
            self.kind = abi.OPT_ADAMW
        elif isinstance(optimizer, torch.optim.SGD):
            self.kind = abi.OPT_SGD
        else:
            raise B200Error(f'no fused step for {type(optimizer).__name__} (SGD and AdamW are built)')
        for g in optimizer.param_groups:
            if self.kind == abi.OPT_SGD and (g.get('nesterov') or g.get('dampening', 0) != 0 or g.get('maximize')):
                raise B200Error('fused SGD implements momentum / weight_decay only (as the reference configs use it)')
            if self.kind == abi.OPT_ADAMW and (g.get('amsgrad') or g.get('maximize')):
                raise B200Error('fused AdamW: amsgrad / maximize are not built')
        self.opt = optimizer
        self._key = None
        self._tensors_dev = self._chunks_dev = None
        self._n_chunks = 0
        self._since_build = 0
        self._base_steps = []
        self._chunk = lib().b200_opt_chunk_elems()

    def _entries(self, grad_src=None):
        out = []
        for g in self.opt.param_groups:
            for p in g['params']:
                if p.grad is None:           # torch skips these too (e.g. the requires_grad=False shift masks)
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise B200Error('fused optimizer step needs contiguous fp32 CUDA parameters and gradients')
                st = self.opt.state[p]
                # data parallel: the gradient is read from the peer arena (the sum of every rank's slot), not from p.grad
                gptr = p.grad.data_ptr() if grad_src is None else grad_src.ptr_of(p)
                if self.kind == abi.OPT_SGD:
                    mom = float(g.get('momentum', 0.0))
                    first = 'momentum_buffer' not in st or st['momentum_buffer'] is None
                    if first:
                        st['momentum_buffer'] = torch.zeros_like(p)
                    out.append((p, gptr, st['momentum_buffer'], None, float(g['lr']), float(g.get('weight_decay', 0.0)), mom, 0.0,
                                0.0, 0 if first else 1))
                else:
                    if 'step' not in st:
                        st['step'] = torch.tensor(0.0)
                        st['exp_avg'] = torch.zeros_like(p)
                        st['exp_avg_sq'] = torch.zeros_like(p)
                    b1, b2 = g['betas']
                    out.append((p, gptr, st['exp_avg'], st['exp_avg_sq'], float(g['lr']), float(g.get('weight_decay', 0.0)),
                                float(b1), float(b2), float(g['eps']), int(st['step'].item())))
        return out

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, grad_src=None) -> None:
        """grad_src (b200/peer.py: ArenaGradSource): take every gradient as the sum of `n_src` arena slots instead of p.grad."""
        ents = self._entries(grad_src)
        if not ents:
            return
        dev = ents[0][0].device
        # The device table is rebuilt only when a pointer or a hyper-parameter changes.  The step counters are NOT part of the
        # key: the kernel adds `offset` (steps since the table was written) to the stored ones, so AdamW's per-step counter
        # and SGD's first-step flag do not force a rebuild + pageable H2D copy every step.
        key = tuple((e[0].data_ptr(), e[1], e[2].data_ptr(), e[4], e[5], e[6], e[7], e[8]) for e in ents)
        offset = self._since_build
        if key == self._key:
            if self.kind == abi.OPT_ADAMW:
                uniform = all(e[9] == b + offset for e, b in zip(ents, self._base_steps))
            else:
                uniform = all((e[9] == 0) == (b + offset == 0) for e, b in zip(ents, self._base_steps))
        if key != self._key or not uniform:
            offset = self._since_build = 0
            self._base_steps = [e[9] for e in ents]
            arr = (abi.OptTensor * len(ents))()
            chunks = []
            for i, (p, g, s1, s2, lr, wd, b1, b2, eps, step) in enumerate(ents):
                t = arr[i]
                t.param, t.grad, t.state1 = p.data_ptr(), g, s1.data_ptr()
                t.state2 = s2.data_ptr() if s2 is not None else 0
                t.param_bf16 = 0
                t.numel, t.lr, t.weight_decay, t.beta1, t.beta2, t.eps, t.step = p.numel(), lr, wd, b1, b2, eps, step
                chunks.extend((i, o) for o in range(0, p.numel(), self._chunk))
            # pinned staging + asynchronous copies: a model whose gradient tensors are fresh allocations every step (the ResNet
            # path: autograd owns them) rebuilds this table every step, and a pageable copy would stall the host until the whole
            # backward has drained
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory()
            chunk_t = torch.tensor(chunks, dtype=torch.int32).pin_memory()
            self._tensors_dev = raw.to(dev, non_blocking=True)
            self._chunks_dev = chunk_t.to(dev, non_blocking=True)
            self._n_chunks = len(chunks)
            self._key = key
        n_src, stride, shift = (1, 0, 0) if grad_src is None else (grad_src.n_src, grad_src.stride, grad_src.shift)
        check(lib().b200_optimizer_step_sum(self.kind, self._tensors_dev.data_ptr(), self._chunks_dev.data_ptr(), self._n_chunks,
                                            float(grad_scale), int(offset), int(n_src), int(stride), int(shift), stream_ptr()),
              'optimizer_step')
        self._since_build += 1
        self.opt._opt_called = True          # what optimizer.step() would set: lr_scheduler.step() checks it (order warning)
        if self.kind == abi.OPT_ADAMW:
            for g in self.opt.param_groups:
                for p in g['params']:
                    if p.grad is not None:
                        self.opt.state[p]['step'] += 1
        plan.bump_weight_epoch()
