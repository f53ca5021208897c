"""Data-parallel gradient exchange over NVLink peer memory (csrc/peer.cu, include/b200_fe.h: b200_peer_*).

The reference trains multi-GPU through Lightning's DDPPlugin (utils/__init__.py:114-119): torch DDP, i.e. a bucketed
NCCL all-reduce of the gradients under the backward.  On a B200 node that collective is an SM kernel, and every
persistent 148-CTA kernel of the backward that it overlaps loses SMs to it and runs a second wave.  Here no SM moves a
gradient:

  * every rank owns an arena ``[2 buffers][world slots][total]`` of fp32, shared through CUDA IPC;
  * as soon as a bucket of gradients is final (engine/trainer.py hooks), rank r PUSHES it into slot r of every rank's arena
    with asynchronous peer copies - copy engines over NVLink / NVSwitch - on a side stream, under the rest of the backward;
  * one barrier closes the step, and the optimizer kernel (b200_optimizer_step_sum) adds the `world` slots in slot order
    while it applies the update: the same order on every rank, so the replicas stay bit-identical, and the reduction is
    never a pass of its own.

The arena is double-buffered: a rank that is already one step ahead writes buffer (step + 1) % 2 while a slower rank's
optimizer still reads buffer step % 2; it cannot get two steps ahead because of the barrier.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Tuple

import torch

from .abi import B200Error, check, lib

ALIGN = 64          # elements: every segment starts on a 256-B boundary


def build_layout(segments: Iterable[Tuple[object, int]]) -> Tuple[Dict[object, int], int]:
    """segments: (key, numel) in a fixed order (identical on every rank) -> ({key: first element}, total elements)."""
    offsets, total = {}, 0
    for key, numel in segments:
        if key in offsets:
            raise B200Error('peer layout: duplicate segment key')
        offsets[key] = total
        total += (int(numel) + ALIGN - 1) // ALIGN * ALIGN
    return offsets, total


class PeerGradExchange:
    def __init__(self, total_elems: int, device: torch.device, group=None):
        import torch.distributed as dist
        if device.type != 'cuda':
            raise B200Error('the peer gradient exchange needs CUDA devices (gloo / CPU runs use the all-reduce path)')
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.total = (int(total_elems) + ALIGN - 1) // ALIGN * ALIGN
        self.device = device
        self.buf = 0
        self.stream = torch.cuda.Stream(device=device)
        self._flag = torch.zeros(1, device=device, dtype=torch.float32)
        self._own = C.c_void_p()
        self._opened: List[int] = []
        self.ptrs: List[int] = []
        handle = C.create_string_buffer(64)
        err = None
        with torch.cuda.device(device):
            try:
                check(lib().b200_peer_alloc(self.arena_bytes, C.byref(self._own), handle), 'peer_alloc')
            except B200Error as e:
                err = e
            handles = [None] * self.world
            dist.all_gather_object(handles, (self.rank, self.total, handle.raw, err is None), group=group)
            if all(h[3] for h in handles):
                try:
                    for r, (rr, tot, raw, _) in enumerate(handles):
                        if rr != r or tot != self.total:
                            raise B200Error(f'peer exchange: rank {r} reports layout ({rr}, {tot}), expected ({r}, {self.total})')
                        if r == self.rank:
                            self.ptrs.append(self._own.value)
                            continue
                        p = C.c_void_p()
                        check(lib().b200_peer_open(raw, C.byref(p)), f'peer_open(rank {r})')
                        self._opened.append(p.value)
                        self.ptrs.append(p.value)
                except B200Error as e:
                    err = e
            elif err is None:
                err = B200Error('peer exchange: another rank could not allocate / export its arena')
            # One collective decides for everybody (a rank that failed must not leave the others waiting), and doubles as the
            # barrier before the first push: every rank has opened every arena.  No stale data is ever summed - the optimizer
            # reads only segments that were pushed in the same step.
            ok = torch.tensor([0 if err is not None else 1], device=device, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self._release()
                raise B200Error(f'peer gradient exchange unavailable on this node: {err or "a peer rank failed"}')

    @property
    def arena_bytes(self) -> int:
        return 2 * self.world * self.total * 4

    def grad_ptr(self, offset: int) -> int:
        """Device address of element `offset` in buffer 0, slot 0 of THIS rank's arena: the `grad` pointer of the optimizer
        table (slot r is `total` elements further, buffer b is `shift(b)` further)."""
        return self._own.value + 4 * int(offset)

    def push(self, src: torch.Tensor, offset: int) -> None:
        """Send `src` (contiguous fp32, final on the current stream) into slot `rank` of every rank's arena at `offset`."""
        if not (src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()):
            raise B200Error('peer push: contiguous fp32 CUDA gradients only')
        n = src.numel()
        if offset < 0 or offset % ALIGN or offset + n > self.total:
            raise B200Error(f'peer push: segment [{offset}, {offset + n}) outside the arena layout ({self.total} elements)')
        ev = torch.cuda.Event()
        ev.record()
        self.stream.wait_event(ev)
        slot = (self.buf * self.world + self.rank) * self.total + offset
        L, sp, nbytes, sptr = lib(), self.stream.cuda_stream, 4 * n, src.data_ptr()
        # ring order: at any moment the `world` senders aim at `world` different receivers
        for k in range(1, self.world + 1):
            r = (self.rank + k) % self.world
            check(L.b200_peer_copy(self.ptrs[r] + 4 * slot, sptr, nbytes, sp), 'peer_copy')

    def finish(self) -> int:
        """Close the step: when this returns (in stream order on the current stream) every rank's pushes have landed here.
        Returns the element shift of the buffer the optimizer must read, and flips the buffers."""
        with torch.cuda.stream(self.stream):
            self.dist.all_reduce(self._flag, group=self.group)      # 4-byte barrier behind this rank's copies
        torch.cuda.current_stream().wait_stream(self.stream)
        shift = self.buf * self.world * self.total
        self.buf ^= 1
        return shift

    def slot_view(self, rank: int, buf: int = 0) -> torch.Tensor:
        """Debug / tests: a copy of slot `rank` of buffer `buf` of this rank's arena."""
        out = torch.empty(self.total, device=self.device, dtype=torch.float32)
        check(lib().b200_peer_copy(out.data_ptr(), self._own.value + 4 * (buf * self.world + rank) * self.total, 4 * self.total,
                                   torch.cuda.current_stream().cuda_stream), 'peer_copy')
        return out

    def _release(self) -> None:
        for p in self._opened:
            lib().b200_peer_close(p)
        self._opened = []
        if self._own.value is not None:
            lib().b200_peer_free(self._own)
            self._own = C.c_void_p()
        self.ptrs = []

    def close(self) -> None:
        """Collective: nobody unmaps or frees while a peer may still be writing."""
        if self._own.value is None and not self._opened:
            return
        torch.cuda.synchronize(self.device)
        try:
            self.dist.barrier(group=self.group)
        except Exception:
            pass
        self._release()


class ArenaGradSource:
    """What FusedStep needs to read gradients out of the arena: parameter -> arena offset, slot count / stride, buffer shift."""

    def __init__(self, exchange: PeerGradExchange, offsets: Dict[int, int]):
        self.exchange, self.offsets = exchange, offsets
        self.shift = 0

    @property
    def n_src(self) -> int:
        return self.exchange.world

    @property
    def stride(self) -> int:
        return self.exchange.total

    def ptr_of(self, param) -> int:
        off = self.offsets.get(id(param))
        if off is None:
            raise B200Error('a parameter with a gradient is missing from the peer arena layout')
        return self.exchange.grad_ptr(off)
