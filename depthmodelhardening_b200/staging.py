"""Host -> device staging of a training batch as ONE transfer.

The reference moves its batch dictionary tensor by tensor (`inputs[key] = ipt.to(self.device)`,
`M2/trainer.py:338-339`): ~20 separate copies per step, each paying its own launch and PCIe ramp.
`BatchArena` lays the dictionary out in a pinned host buffer and a device buffer of the same layout (256-byte
aligned entries), one pair per slot; `upload()` is a single `cudaMemcpyAsync`, the tensors handed to the kernels are
views.

Double-buffering contract: slot k's pinned buffer may be refilled only after its previous upload has finished --
`upload(k)` records an event, `host_views(k)` waits for it (host-side) before handing the views out, so filling
batch k+1 while batch k uploads or computes can never tear a batch.
"""
from __future__ import annotations

from typing import Dict, Hashable

import torch

_ALIGN = 256


class BatchArena:
    def __init__(self, example: Dict[Hashable, torch.Tensor], device, slots: int = 2):
        """example: key -> host tensor giving shape/dtype of every entry; `slots` device copies (double buffering)."""
        self.layout = {}
        off = 0
        for k, v in example.items():
            n = v.numel() * v.element_size()
            self.layout[k] = (off, tuple(v.shape), v.dtype, n)
            off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
        self.nbytes = off
        self.hosts = [torch.empty(off, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.dev = [torch.empty(off, dtype=torch.uint8, device=device) for _ in range(slots)]
        self._uploaded = [None] * slots                    # event of the last H2D copy that read hosts[k]

    def _views(self, buf):
        out = {}
        for k, (off, shape, dtype, n) in self.layout.items():
            out[k] = buf[off:off + n].view(dtype).view(shape)
        return out

    @property
    def host(self):
        return self.hosts[0]

    def host_views(self, slot: int = 0) -> Dict[Hashable, torch.Tensor]:
        """Pinned host tensors of `slot` to fill (e.g. as DataLoader collate targets).  Blocks until the last upload
        from this slot's pinned buffer has completed (the DMA engine may still be reading it)."""
        ev = self._uploaded[slot]
        if ev is not None:
            ev.synchronize()
        return self._views(self.hosts[slot])

    def device_views(self, slot: int) -> Dict[Hashable, torch.Tensor]:
        return self._views(self.dev[slot])

    def upload(self, slot: int, stream=None) -> None:
        """One async H2D copy of the whole batch into device slot `slot` (on `stream` or the current stream)."""
        if stream is None:
            stream = torch.cuda.current_stream(self.dev[slot].device)
        with torch.cuda.stream(stream):
            self.dev[slot].copy_(self.hosts[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        self._uploaded[slot] = ev


def unpack_u8(src: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """uint8 CUDA tensor -> fp32 `src / 255` (IEEE division: torchvision `to_tensor`, the conversion every colour
    frame of the reference goes through in its loaders, `mono_dataset.py:79,133-144`) by `dmh_unpack_u8`.

    Lets a batch cross PCIe as bytes -- a quarter of the fp32 volume -- and arrive bit-identical to the fp32 tensors
    the reference's DataLoader would have produced from the same 8-bit images.  `out`: optional preallocated fp32
    tensor of the same shape (16-byte aligned, as torch allocations and `BatchArena` slots are)."""
    from . import _lib
    if src.dtype != torch.uint8:
        raise RuntimeError("unpack_u8: expected a uint8 tensor, got %s" % src.dtype)
    lib = _lib.load()
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    elif out.shape != src.shape or out.dtype != torch.float32:
        raise RuntimeError("unpack_u8: `out` must be fp32 of shape %s" % (tuple(src.shape),))
    if src.numel() == 0:
        if not src.is_cuda:
            raise RuntimeError("dmh_b200: tensor is on %s; the hot path is CUDA-only (no CPU fallback)" % src.device)
        return out
    _lib.check(lib.dmh_unpack_u8(_lib.ptr(src), src.numel(), _lib.ptr(out), _lib.stream()), "unpack_u8")
    return out
