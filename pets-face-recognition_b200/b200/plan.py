"""Host side of the native Swin plan (csrc/swin_plan.cu): flat parameter storage, workspace ownership
and the autograd bridge.  PyTorch is plumbing here - device memory, streams, autograd bookkeeping;
every FLOP of the backbone runs inside b200_swin_forward / b200_swin_backward.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import abi
from .abi import B200Error, check, lib, ptr, stream_ptr

# bumped by the fused optimizer (which updates parameters through raw pointers, invisible to
# tensor._version) so that plans refresh their bf16 weight caches
_weight_epoch = 0


def bump_weight_epoch() -> None:
    global _weight_epoch
    _weight_epoch += 1


class _Plan:
    """One native plan = (batch, training) specialisation: buffer layout only, no memory."""

    def __init__(self, spec: dict, batch: int, training: bool):
        L = lib()
        arr = lambda v: (C.c_int * 4)(*v)
        self.handle = L.b200_swin_create(batch, spec['img'], spec['channels'], spec['hidden_dim'], arr(spec['layers']),
                                         arr(spec['heads']), arr(spec['downscaling_factors']), spec['num_classes'],
                                         spec['head_dim'], spec['window_size'], int(training))
        if not self.handle:
            raise B200Error(f'swin_create failed: {abi.last_error()}')
        self.batch, self.training = batch, training
        self.workspace_bytes = L.b200_swin_workspace_bytes(self.handle)
        self.workspace = None

    def layout(self) -> Tuple[List[int], List[int], int]:
        L = lib()
        n = L.b200_swin_param_count(self.handle)
        off = (C.c_longlong * n)()
        num = (C.c_longlong * n)()
        check(L.b200_swin_param_offsets(self.handle, off, num, n), 'swin_param_offsets')
        return list(off), list(num), L.b200_swin_param_elems(self.handle)

    def __del__(self):
        try:
            if self.handle:
                lib().b200_swin_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class SwinEngine:
    """Owns the flat fp32 parameters, the bf16 weight cache and the per-batch-size plans of one model."""

    def __init__(self, spec: dict, params: List[torch.nn.Parameter]):
        self.spec = dict(spec)
        self.params = params                  # trainable parameters in state-dict order (masks excluded)
        self.flat = None
        self.wcache = None
        self.offsets = self.numels = None
        self.total = 0
        self.plans: Dict[Tuple[int, bool], _Plan] = {}
        self._cache_key = None
        self._pending = None                  # (plan, token) of the last training forward
        self._token = 0
        self.last_flat_grad = None

    # ------------------------------------------------------------------ parameters
    def _ensure_flat(self, device) -> None:
        if self.offsets is None:
            probe = _Plan(self.spec, 1, False)
            self.offsets, self.numels, self.total = probe.layout()
            self.wcache_bytes = lib().b200_swin_wcache_bytes(probe.handle)
            if len(self.offsets) != len(self.params) or any(n != p.numel() for n, p in zip(self.numels, self.params)):
                raise B200Error('parameter list does not match the native plan layout')
        ok = self.flat is not None and self.flat.device == device
        if ok:
            base = self.flat.data_ptr()
            ok = all(p.data_ptr() == base + 4 * o for p, o in zip(self.params, self.offsets))
        if ok:
            return
        flat = torch.zeros(self.total, device=device, dtype=torch.float32)
        for p, o, n in zip(self.params, self.offsets, self.numels):
            view = flat[o:o + n].view(p.shape)
            view.copy_(p.data.to(device=device, dtype=torch.float32))
            p.data = view
        self.flat = flat
        self.wcache = torch.empty(self.wcache_bytes, device=device, dtype=torch.uint8)
        self._cache_key = None

    def _sync_weights(self) -> None:
        key = (_weight_epoch, sum(p._version for p in self.params), self.flat.data_ptr())
        if key != self._cache_key:
            plan = next(iter(self.plans.values()))
            check(lib().b200_swin_sync_weights(plan.handle, ptr(self.flat), ptr(self.wcache), stream_ptr()), 'swin_sync_weights')
            self._cache_key = key

    def _plan(self, batch: int, training: bool, device) -> _Plan:
        key = (batch, training)
        plan = self.plans.get(key)
        if plan is None:
            plan = self.plans[key] = _Plan(self.spec, batch, training)
        if plan.workspace is None or plan.workspace.device != device:
            plan.workspace = torch.empty(plan.workspace_bytes, device=device, dtype=torch.uint8)
        return plan

    # ------------------------------------------------------------------ compute
    def forward(self, img: torch.Tensor, training: bool) -> torch.Tensor:
        if not img.is_cuda:
            raise B200Error('SwinTransformer.forward needs a CUDA tensor: the B200 path has no CPU fallback')
        abi.require_device()
        if img.dim() != 4 or img.shape[1] != self.spec['channels'] or img.shape[2] != self.spec['img'] or img.shape[3] != self.spec['img']:
            raise B200Error(f'expected (B, {self.spec["channels"]}, {self.spec["img"]}, {self.spec["img"]}) input, got {tuple(img.shape)}')
        is_u8 = img.dtype == torch.uint8          # raw pixels: ToTensor's /255 is fused into the first gather kernel
        img = img.contiguous() if is_u8 else img.contiguous().float()
        self._ensure_flat(img.device)
        plan = self._plan(img.shape[0], training, img.device)
        self._sync_weights()
        emb = torch.empty(img.shape[0], self.spec['num_classes'], device=img.device, dtype=torch.float32)
        check(lib().b200_swin_forward(plan.handle, ptr(self.flat), ptr(self.wcache), ptr(img), int(is_u8), ptr(emb), ptr(plan.workspace),
                                      plan.workspace_bytes, stream_ptr()), 'swin_forward')
        if training:
            self._token += 1
            self._pending = (plan, self._token)
        return emb

    def backward(self, demb: torch.Tensor, token: int, hooks=None) -> torch.Tensor:
        if self._pending is None or self._pending[1] != token:
            raise B200Error('backward through a Swin forward whose saved activations were overwritten by a later forward '
                            '(one training forward per backward on a given batch size)')
        plan, _ = self._pending
        flat_grad = torch.zeros(self.total, device=demb.device, dtype=torch.float32)
        demb = demb.contiguous().float()
        # stage by stage (head+stage4, stage3, stage2, stage1) so that a caller can overlap gradient
        # all-reduce buckets with the remaining backward (engine/trainer.py: hooks(stage_index, flat_grad))
        for hi, lo in ((4, 3), (2, 2), (1, 1), (0, 0)):
            check(lib().b200_swin_backward(plan.handle, ptr(self.flat), ptr(self.wcache), ptr(demb), ptr(flat_grad),
                                           ptr(plan.workspace), plan.workspace_bytes, hi, lo, stream_ptr()), 'swin_backward')
            if hooks is not None:
                hooks(lo, flat_grad)
        self._pending = None
        self.last_flat_grad = flat_grad
        return flat_grad

    def stage_param_range(self, stage: int) -> Tuple[int, int]:
        """[begin, end) element range of the flat buffers holding stage `stage` (0-3); the head follows stage 3."""
        per_stage = [2 + 12 * n for n in self.spec['layers']]
        first = sum(per_stage[:stage])
        last = first + per_stage[stage]
        begin = self.offsets[first]
        end = self.total if stage == 3 else self.offsets[last]
        return begin, end


class SwinFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: SwinEngine, img, *params):
        # only used when gradients are wanted (grad mode is always off inside Function.forward, so the
        # caller - models/swin.py - decides); inference calls engine.forward(img, False) directly
        emb = engine.forward(img, True)
        ctx.engine = engine
        ctx.token = engine._token
        return emb

    @staticmethod
    def backward(ctx, demb):
        engine = ctx.engine
        flat = engine.backward(demb, ctx.token, getattr(engine, 'grad_hook', None))
        grads = []
        for i, (p, o, n) in enumerate(zip(engine.params, engine.offsets, engine.numels)):
            grads.append(flat[o:o + n].view(p.shape) if ctx.needs_input_grad[2 + i] else None)
        return (None, None) + tuple(grads)
