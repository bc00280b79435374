"""Side-stream host-to-device prefetcher for the OTF feed — the role of ``CUDAPrefetcher``
(traiNNer/data/prefetch_dataloader.py:418-499) in the reference's training loop (train.py:526-537, :975).

Same protocol: ``preload()`` takes the next batch from the loader and starts its upload on a private copy stream,
``next()`` makes the caller's stream wait for that upload (``wait_stream``, :488-493), hands the batch out and preloads
the following one, ``reset()`` restarts the loader.  Two differences, both on purpose:

  * the device tensors are STATIC slots (``slots`` per key, reused round-robin) instead of fresh allocations, so the
    addresses ``RealESRGANFeed.feed_data`` sees repeat and its captured chains (chain_graph.py) are replayed instead of
    re-recorded.  A slot is overwritten only after every kernel the consumer issued for the batch that last lived in it
    (the copy stream waits for the consumer's stream before reusing a slot);
  * nothing is converted on the way: a ``uint8`` GT batch stays ``uint8`` (a quarter of the PCIe bytes; ``feed_data``
    normalises it on the device), fp32 stays fp32 — the reference's format.

No error swallowing / retry logic: worker failures are the loader's business (out of scope).
"""

from __future__ import annotations

from typing import Any, Iterable

import torch
from torch import Tensor


class CUDAPrefetcher:
    def __init__(self, loader: Iterable[dict], device: torch.device | str = "cuda", slots: int = 2) -> None:
        self.ori_loader = loader
        self.loader = iter(loader)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CUDAPrefetcher uploads to a CUDA device")
        self.stream = torch.cuda.Stream(self.device)
        self.slots = max(2, int(slots))
        self._bufs: dict[tuple, list[Tensor]] = {}
        self._turn = 0
        self.batch: dict | None = None
        self.h2d_bytes = 0  # bytes of the batch being handed out (bench.py reports it)
        self.preload()

    def _slot(self, key: str, v: Tensor) -> Tensor:
        k = (key, tuple(v.shape), v.dtype)
        ring = self._bufs.get(k)
        if ring is None:
            ring = self._bufs[k] = [torch.empty(v.shape, dtype=v.dtype, device=self.device) for _ in range(self.slots)]
        return ring[self._turn % self.slots]

    def _upload(self, key: str, v: Any) -> Any:
        if isinstance(v, Tensor):
            if v.is_cuda:
                return v
            self._bytes += v.numel() * v.element_size()
            return self._slot(key, v).copy_(v, non_blocking=True)
        if isinstance(v, (tuple, list)) and v and all(isinstance(t, Tensor) for t in v):
            return type(v)(self._upload(f"{key}[{i}]", t) for i, t in enumerate(v))
        return v

    def preload(self) -> None:
        try:
            batch = next(self.loader)
        except StopIteration:
            self.batch = None
            return
        # slot reuse: everything the consumer has issued so far (which includes all work on the batch that last used
        # this slot, handed out `slots` calls ago) must be done before the copy overwrites it
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self._bytes = 0
        with torch.cuda.stream(self.stream):
            self.batch = {k: self._upload(k, v) for k, v in batch.items()}
        self._next_bytes = self._bytes
        self._turn += 1

    def next(self) -> dict | None:
        torch.cuda.current_stream(self.device).wait_stream(self.stream)  # prefetch_dataloader.py:488-493
        batch = self.batch
        self.h2d_bytes = getattr(self, "_next_bytes", 0)
        self.preload()
        return batch

    def reset(self, loader: Iterable[dict] | None = None) -> None:
        """Restart the loader (prefetch_dataloader.py:495-499); ``loader`` swaps in another one and KEEPS the static
        slots, so the chains captured against their addresses stay valid from one epoch / loader to the next."""
        if loader is not None:
            self.ori_loader = loader
        self.loader = iter(self.ori_loader)
        self.preload()
