"""Multi-GPU plumbing: one process per GPU (torchrun), batch sharded over ranks,
no data-path collective except the two the path really has:

  1. all-reduce(sum) of the SHARED patch gradient, with the scalar attack loss
     riding in the tail of the same buffer -- one collective per PGD step;
  2. all-reduce(sum) of the per-scale loss numerators once per training step
     (denominators are static: global B*H*W).

Everything else is per-image and needs no communication.  Works with any
torch.distributed backend (nccl on the GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(batch: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Items [lo, hi) of a global batch owned by `rank` (equal shards; the batch
    must divide evenly, as `BackprojectDepth(batch_size, ...)` bakes the shard
    size in -- `M2/layers.py:154-161`)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if batch % world_size != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (batch, world_size))
    per = batch // world_size
    return rank * per, (rank + 1) * per


def shard_batch(tensors: Dict, batch: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Dict:
    """Slice every tensor whose leading dim is the global batch; others are replicated."""
    lo, hi = shard_range(batch, rank, world_size)
    out = {}
    for k, v in tensors.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == batch:
            out[k] = v[lo:hi].contiguous()
        else:
            out[k] = v
    return out


def resolve_group(group):
    """True / None -> the default process group; a ProcessGroup -> itself."""
    return None if group is True or group is None else group


def broadcast_patch_state(tensors: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Make the shared patch's starting point (random start, L0 patterns) rank `src`'s on every rank, in place."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for t in tensors:
        dist.broadcast(t, src=src, group=group)


def allreduce_patch_grad(grad: torch.Tensor, scalars: Sequence[torch.Tensor] = (), average: bool = True, group=None,
                         out: Optional[torch.Tensor] = None):
    """Sum `grad` (the shared patch's gradient) over ranks; `scalars` (0-dim
    tensors, e.g. the attack loss) are appended to the same flat buffer so the
    step costs ONE collective.  Returns (grad, [scalars...]) -- identical on
    every rank, so the sign / Adam / threshold update that follows stays
    bit-identical across ranks.  Without scalars the all-reduce runs in place on `grad` (no staging launches)."""
    if not (dist.is_available() and dist.is_initialized()):
        return grad, list(scalars)
    w = dist.get_world_size(group)
    if w == 1:
        return grad, list(scalars)
    n = grad.numel()
    if len(scalars) == 0 and grad.is_contiguous():
        buf = grad.view(-1)                      # in place: no staging buffer, no copy
    else:
        buf = out if out is not None else torch.empty(n + len(scalars), device=grad.device, dtype=grad.dtype)
        buf[:n].copy_(grad.reshape(-1))
        for i, s in enumerate(scalars):
            buf[n + i] = s.to(grad.dtype)
    dist.all_reduce(buf, group=group)
    if average:
        buf.div_(w)
    return buf[:n].view_as(grad), [buf[n + i] for i in range(len(scalars))]


def allreduce_loss_sums(sums: torch.Tensor) -> torch.Tensor:
    """Sum the per-scale loss numerators (1-D tensor) over ranks, in place."""
    _, w = world()
    if w > 1:
        dist.all_reduce(sums)
    return sums


class PeerReducer:
    """The all-reduce of the shared patch gradient as ONE kernel over NVLink peer memory (csrc/peer_reduce.cu,
    dmh_peer_allreduce) instead of an NCCL call + a scaling launch: every rank's backward kernel accumulates into
    `self.buffer` -- a symmetric allocation mapped into every process (torch.distributed._symmetric_memory is the
    plumbing: CUDA VMM handles exchanged once at construction) --, `allreduce()` sums the ranks' buffers in rank order
    (the same bits on every rank: the update that follows keeps the universal patch bit-identical), scales by 1/world
    and, for the L-inf attack, applies the PGD update in the same launch.  CUDA-graph capturable (the step counter
    lives on the device).  Construct it on every rank of `group` (collective: rendezvous + barrier).

    `PeerReducer.available()` is False without CUDA peer access / symmetric memory (gloo, a single GPU, other
    backends): callers then keep `allreduce_patch_grad` (NCCL)."""

    FLAG_WORDS = 64          # 2 * world <= 32 used; one 256-byte block behind the gradient

    @staticmethod
    def available(group=None) -> bool:
        if not (dist.is_available() and dist.is_initialized() and torch.cuda.is_available()):
            return False
        if dist.get_backend(group) != "nccl" or dist.get_world_size(group) < 2 or dist.get_world_size(group) > 16:
            return False
        try:
            import importlib
            importlib.import_module("torch.distributed._symmetric_memory")
        except Exception:
            return False
        return True

    def __init__(self, numel: int, device, group=None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.n = int(numel)
        self.n_pad = (self.n + 3) // 4 * 4
        self._alloc = symm.empty(self.n_pad + self.FLAG_WORDS, dtype=torch.float32, device=device)
        self._hdl = symm.rendezvous(self._alloc, self.group)
        self._alloc.zero_()
        self.buffer = self._alloc[:self.n]                       # what the backward kernel accumulates into
        self.out = torch.zeros(self.n_pad, device=device, dtype=torch.float32)
        self.state = torch.zeros(2, device=device, dtype=torch.int32)
        ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        self._bufs = (C.c_void_p * self.world)(*ptrs)
        self._flags = (C.c_void_p * self.world)(*[p + 4 * self.n_pad for p in ptrs])
        self._lib = _lib.load()
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)                           # every rank's flags are zero before the first step

    def allreduce(self, average: bool = True, linf=None):
        """Enqueue the kernel on the current stream.  Returns the reduced gradient (flat, n floats; identical on
        every rank).  linf = (adv, clean, alpha, eps, adv_out): also apply the L-inf update to the first
        adv.numel() elements."""
        from ._lib import check, ptr, stream
        scale = 1.0 / self.world if average else 1.0
        if linf is not None:
            adv, clean, alpha, eps, adv_out = linf
            check(self._lib.dmh_peer_allreduce(self._bufs, self._flags, self.rank, self.world, self.n, scale, ptr(self.out),
                                               ptr(self.state), ptr(adv), ptr(clean), adv.numel(), float(alpha), float(eps),
                                               ptr(adv_out), stream()), "peer_allreduce")
        else:
            check(self._lib.dmh_peer_allreduce(self._bufs, self._flags, self.rank, self.world, self.n, scale, ptr(self.out),
                                               ptr(self.state), None, None, 0, 0.0, 0.0, None, stream()), "peer_allreduce")
        return self.out[:self.n]
