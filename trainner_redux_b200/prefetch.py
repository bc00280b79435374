"""Side-stream host-to-device prefetcher for the OTF feed — the role of ``CUDAPrefetcher``
(traiNNer/data/prefetch_dataloader.py:418-499) in the reference's training loop (train.py:526-537, :975).

Same protocol: ``preload()`` takes the next batch from the loader and starts its upload on a private copy stream,
``next()`` makes the caller's stream wait for that upload (``wait_stream``, :488-493), hands the batch out and preloads
the following one, ``reset()`` restarts the loader.  The batch is expected to be consumed on the stream that is current
when ``next()`` returns it (as in the reference's loop).  Two differences, both on purpose:

  * the device tensors are STATIC slots (``slots`` per key, reused round-robin) instead of fresh allocations, so the
    addresses ``RealESRGANFeed.feed_data`` sees repeat and its captured chains (chain_graph.py) are replayed instead of
    re-recorded.  A slot is overwritten only after every kernel the consumer issued for the batch that last lived in it
    (the copy stream waits for the consumer's stream before reusing a slot);
  * nothing is converted on the way: a ``uint8`` GT batch stays ``uint8`` (a quarter of the PCIe bytes; ``feed_data``
    normalises it on the device), fp32 stays fp32 — the reference's format.

The upload step itself is ONE library call (``otf_upload_async``, csrc/hostrt.cu): order the copy stream behind the
consumer, one ``cudaMemcpyAsync`` per tensor, record the batch's "ready" event — the Python-level stream context,
``wait_stream`` pair and per-tensor ``copy_`` of the first version cost more host time than the whole chain's launch.
Host batches are kept referenced until their copies have completed (the raw copies are invisible to torch's pinned
memory allocator).

No error swallowing / retry logic: worker failures are the loader's business (out of scope).
"""

from __future__ import annotations

import ctypes as C
from collections import deque
from typing import Any, Iterable

import torch
from torch import Tensor

from . import _lib


class CUDAPrefetcher:
    def __init__(self, loader: Iterable[dict], device: torch.device | str = "cuda", slots: int = 2) -> None:
        self.ori_loader = loader
        self.loader = iter(loader)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CUDAPrefetcher uploads to a CUDA device")
        _lib.load()
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(self.device)
        self._copy_stream = C.c_void_p(self.stream.cuda_stream)
        self.slots = max(2, int(slots))
        self._bufs: dict[tuple, list[Tensor]] = {}
        self._turn = 0
        with torch.cuda.device(self.device):
            self._ready = [self._new_event() for _ in range(self.slots)]  # per slot: its upload has been issued up to here
            self._consumed = self._new_event()
        self._hold: deque[tuple[int, Any]] = deque()  # (slot, host batch) until the slot's copies have completed
        self._pending: int | None = None  # slot of the batch waiting to be handed out
        self._slot_stream: list[int | None] = [None] * self.slots  # raw handle of the stream each slot's batch was handed out on
        self.batch: dict | None = None
        self.h2d_bytes = 0  # bytes of the batch being handed out (bench.py reports it)
        self._next_bytes = 0
        self.preload()

    @staticmethod
    def _new_event() -> C.c_void_p:
        ev = C.c_void_p()
        _lib.call("otf_event_create", C.byref(ev))
        return ev

    def __del__(self) -> None:
        try:
            for ev in [*getattr(self, "_ready", []), getattr(self, "_consumed", None)]:
                if ev is not None:
                    _lib.call("otf_event_destroy", ev)
        except Exception:  # noqa: BLE001  interpreter shutdown
            pass

    def _slot(self, key: str, v: Tensor) -> Tensor:
        k = (key, v.shape, v.dtype)
        ring = self._bufs.get(k)
        if ring is None:
            ring = self._bufs[k] = [torch.empty(v.shape, dtype=v.dtype, device=self.device) for _ in range(self.slots)]
        return ring[self._turn % self.slots]

    def _stage(self, key: str, v: Any, todo: list) -> Any:
        """The device-side stand-in of one batch entry; host tensors are appended to ``todo`` as (dst, src)."""
        if isinstance(v, Tensor):
            if v.is_cuda:
                return v
            if not v.is_contiguous():
                v = v.contiguous()
            dst = self._slot(key, v)
            todo.append((dst, v))
            return dst
        if isinstance(v, (tuple, list)) and v and all(isinstance(t, Tensor) for t in v):
            return type(v)(self._stage(f"{key}[{i}]", t, todo) for i, t in enumerate(v))
        return v

    def _release_done(self) -> None:
        done = C.c_int(0)
        while self._hold:
            _lib.call("otf_event_query", self._ready[self._hold[0][0]], C.byref(done))
            if not done.value:
                break
            self._hold.popleft()

    def _issue(self, n: int, dst: Any, src: Any, nbytes: Any, slot: int) -> None:
        self._release_done()
        # the consumer of this slot's previous batch: the stream that was current when ``next()`` handed it out (a loop
        # that alternates its batches over several streams keeps working: the copy waits for the right one)
        consumer = self._slot_stream[slot]
        _lib.call("otf_upload_async", n, dst, src, nbytes, self._copy_stream, _lib.stream() if consumer is None else C.c_void_p(consumer),
                  self._consumed, self._ready[slot])

    def preload(self) -> None:
        try:
            batch = next(self.loader)
        except StopIteration:
            self.batch = None
            self._pending = None
            return
        todo: list[tuple[Tensor, Tensor]] = []
        self.batch = {k: self._stage(k, v, todo) for k, v in batch.items()}
        n = len(todo)
        slot = self._turn % self.slots
        dst = (C.c_void_p * n)(*[d.data_ptr() for d, _ in todo])
        src = (C.c_void_p * n)(*[s.data_ptr() for _, s in todo])
        sizes = [s.numel() * s.element_size() for _, s in todo]
        nbytes = (C.c_uint64 * n)(*sizes)
        # slot reuse: everything the consumer has issued so far (which includes all work on the batch that last used
        # this slot, handed out `slots` calls ago) must be done before the copies overwrite it
        if torch.cuda.current_device() != self.device.index:  # (multi-device processes only: a context costs microseconds)
            with torch.cuda.device(self.device):
                self._issue(n, dst, src, nbytes, slot)
        else:
            self._issue(n, dst, src, nbytes, slot)
        self._hold.append((slot, (batch, [src_t for _, src_t in todo])))  # (the sources as uploaded: a .contiguous() copy is not in `batch`)
        self._pending = slot
        self._next_bytes = sum(sizes)
        self._turn += 1

    def next(self) -> dict | None:
        if self._pending is not None:  # the caller's stream waits for the upload: prefetch_dataloader.py:488-493
            cur = _lib.stream()
            _lib.call("otf_stream_wait_event", cur, self._ready[self._pending])
            self._slot_stream[self._pending] = cur.value or 0
        batch = self.batch
        self.h2d_bytes = self._next_bytes
        self.preload()
        return batch

    def reset(self, loader: Iterable[dict] | None = None) -> None:
        """Restart the loader (prefetch_dataloader.py:495-499); ``loader`` swaps in another one and KEEPS the static
        slots, so the chains captured against their addresses stay valid from one epoch / loader to the next."""
        if loader is not None:
            self.ori_loader = loader
        self.loader = iter(self.ori_loader)
        self.preload()


class CUDAReadback:
    """Side-stream device-to-host read of a step's result (what a training loop does with ``.cpu()`` on a loss or a
    visual, e.g. traiNNer/models/sr_model.py ``get_current_visuals``) that does not sit on the compute stream: the copy
    runs on a private stream behind the producer, into a ring of pinned host buffers.

    ``read(t)`` first makes the CURRENT stream wait for the previous read — at most one read is in flight, so a device
    buffer the producer rewrites two or more calls later (the static outputs of a captured chain) is never overwritten
    under a copy — then issues the copy of ``t`` and returns the pinned host tensor it lands in; ``wait()`` blocks the
    host until the last read has completed."""

    def __init__(self, device: torch.device | str = "cuda", depth: int = 2) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CUDAReadback reads from a CUDA device")
        _lib.load()
        self.stream = torch.cuda.Stream(self.device)
        self._copy_stream = C.c_void_p(self.stream.cuda_stream)
        self.depth = max(1, int(depth))
        self._host: dict[tuple, list[Tensor]] = {}
        self._turn = 0
        with torch.cuda.device(self.device):
            self._produced = CUDAPrefetcher._new_event()
            self._done = CUDAPrefetcher._new_event()
        self._issued = False

    def __del__(self) -> None:
        try:
            for ev in (getattr(self, "_produced", None), getattr(self, "_done", None)):
                if ev is not None:
                    _lib.call("otf_event_destroy", ev)
        except Exception:  # noqa: BLE001  interpreter shutdown
            pass

    def read(self, t: Tensor, as_u8: bool = False) -> Tensor:
        """``as_u8=True``: ``t`` is an fp32 image on the 8-bit lattice (a finished LQ batch, realesrgan_model.py:616) and
        crosses PCIe as ``clamp(round(t * 255), 0, 255)`` bytes (one launch of ``otf_f32_to_u8`` on the current stream into
        a device staging buffer of the ring); ``host.float() / 255`` restores ``t`` bit for bit.  A quarter of the bytes: at
        eight ranks per host the fp32 read-back competes with the uploads for the path to host memory
        (profiles/r02_e2e_readback_n8.json)."""
        _lib.require_cuda(t)
        if not t.is_contiguous():
            t = t.contiguous()
        cur = _lib.stream()
        if self._issued:  # (the copy stream is in order: the last read done = every earlier one done, staging buffers included)
            _lib.call("otf_stream_wait_event", cur, self._done)
        if as_u8:
            if t.dtype != torch.float32:
                raise TypeError("CUDAReadback.read(as_u8=True) expects a float32 tensor")
            if t.data_ptr() % 16:  # (a view that starts inside a 16-byte group: the kernel reads float4s)
                t = t.clone()
            ku = ("u8dev", t.shape)
            dring = self._host.get(ku)
            if dring is None:  # (device-side staging, one per ring position: the copy of call i may still run during call i + 1)
                dring = self._host[ku] = [torch.empty(t.shape, dtype=torch.uint8, device=t.device) for _ in range(self.depth)]
            u = dring[self._turn % self.depth]
            _lib.call("otf_f32_to_u8", _lib.ptr(t), t.numel(), _lib.ptr(u), cur)
            t = u
        k = (t.shape, t.dtype)
        ring = self._host.get(k)
        if ring is None:
            ring = self._host[k] = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for _ in range(self.depth)]
        dst = ring[self._turn % self.depth]
        self._turn += 1
        _lib.call("otf_download_async", C.c_void_p(dst.data_ptr()), _lib.ptr(t), t.numel() * t.element_size(), self._copy_stream, cur,
                  self._produced, self._done)
        self._issued = True
        return dst

    def wait(self) -> None:
        if self._issued:
            _lib.call("otf_event_synchronize", self._done)
