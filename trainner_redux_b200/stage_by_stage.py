"""The recording interface of ``stages.StageList`` executed AT ONCE, one public per-primitive call per stage.

``RealESRGANFeed._record`` describes a chain (traiNNer/models/realesrgan_model.py:512-616 for the fork's order,
traiNNer/utils/redux_options.py:720-851 for the classical one) by calling ``analyse / filter2d / resize / gaussian_noise /
jpeg / ...`` on whatever object it is handed: a ``StageList`` collects launch records for the native executor, a
``chain_graph.ParamCollector`` collects the per-step numbers of a captured chain — and this class runs every stage right
away through the package's own Python wrappers (``filter2d``, ``resize_pt``, ``DiffJPEG``, the ``*_noise_pt`` functions,
``paragon_otf``): the per-stage path behind the stage hooks (per-stage CUDA-event timing, stage closures for bench.py,
intermediate taps for parity localisation).  One description of the chain, three consumers; the native path and this
one are compared bit for bit in tests/test_chain_native_gpu.py.
"""

from __future__ import annotations

from typing import Any, Callable, Sequence

import numpy as np
from torch import Tensor

from . import _lib
from . import degradations as D
from .img_process_util import KernelAnalysis, filter2d


class StageByStage:
    def __init__(self, feed: Any, img: Tensor) -> None:
        """``feed``: the RealESRGANFeed whose hooks (``_timed``) and modules (``jpeger``) are used."""
        self.feed = feed
        self.img = _lib.dense_f32(img)
        self.device = self.img.device
        self._ka: KernelAnalysis | None = None
        self._name = "stage"

    # -- extent of the running image (what _record reads) ---------------------------------------
    @property
    def b(self) -> int:
        return self.img.size(0)

    @property
    def c(self) -> int:
        return self.img.size(1)

    @property
    def h(self) -> int:
        return self.img.size(2)

    @property
    def w(self) -> int:
        return self.img.size(3)

    def at(self, name: str) -> "StageByStage":
        """Name of the stage the next call runs (key of the timing / closure / tap hooks)."""
        self._name = name
        return self

    def _do(self, fn: Callable[[Tensor], Tensor]) -> None:
        self.img = self.feed._timed(self._name, lambda o=self.img: fn(o))

    def _dev(self, v: Any) -> Any:
        return v.to(self.device, non_blocking=True) if isinstance(v, Tensor) else v

    # -- the classical chain's stages ---------------------------------------------------------------
    def analyse(self, kernels: Sequence[Tensor]) -> None:
        self._ka = KernelAnalysis(list(kernels))  # one launch for all kernel tensors of the step

    def filter2d(self, kernel: Tensor, analysed_set: int | None = None) -> None:
        an = (self._ka, analysed_set) if analysed_set is not None and self._ka is not None else None
        self._do(lambda o: filter2d(o, kernel, _analysis=an))

    def usm(self, sharpener: Any, weight: float, threshold: float) -> None:
        self._do(lambda o: sharpener(o, weight, threshold))

    def resize(self, mode: str, scale_factor: float = 0, size: tuple[int, int] = (0, 0)) -> None:
        self._do(lambda o: D.resize_pt(o, mode, scale_factor=scale_factor, size=size))

    def resize_raw(self, mode_id: int, oh: int, ow: int, clamp: bool) -> None:
        self._do(lambda o: D._resize_call(_lib.dense_f32(o), int(oh), int(ow), mode_id, clamp))

    def gaussian_noise(self, sigma: Any, gray: Any, gen: D.PhiloxState, noise: Tensor | None = None,
                       noise_gray: Tensor | None = None) -> None:
        self._do(lambda o: D.add_gaussian_noise_pt(o, self._dev(sigma), self._dev(gray), clip=True, rounds=False, noise=noise,
                                                   noise_gray=noise_gray, generator=gen))

    def poisson_noise(self, scale: Any, gray: Any, gen: D.PhiloxState, counts: Tensor | None = None,
                      counts_gray: Tensor | None = None) -> None:
        self._do(lambda o: D.add_poisson_noise_pt(o, self._dev(scale), True, False, self._dev(gray), poisson_counts=counts,
                                                  poisson_counts_gray=counts_gray, generator=gen))

    def noise_field(self, field: Tensor) -> None:
        self._do(lambda o: D.add_noise_field_pt(o, field, clip=True, rounds=False))

    def jpeg(self, quality: float | Tensor, differentiable: bool = False, clamp_in: bool = True, round8: bool = False) -> None:
        if isinstance(quality, Tensor):  # raw per-sample qualities: the kernel converts them, nothing is mutated
            q = self._dev(quality)
            self._do(lambda o: self.feed.jpeger(o, quality=q, _clamp_in=clamp_in, _round8=round8, _keep_quality=True))
        else:
            self._do(lambda o: self.feed.jpeger(o, quality=float(quality), _clamp_in=clamp_in, _round8=round8))

    def clamp_round(self) -> None:
        from .realesrgan_feed import clamp_round

        self._do(clamp_round)

    # -- the fork's extras (paragon_otf.py) ---------------------------------------------------------
    def warp(self, mode: int, p0: float) -> None:
        from . import paragon_otf as PO

        self._do(lambda o: PO._warp(o, mode, p0))

    def taps_zero(self, kernel: np.ndarray, epilogue: int = _lib.TAPS_NONE, strength: float = 0.0) -> None:
        from . import paragon_otf as PO

        self._do(lambda o: PO._taps_zero(o, kernel, epilogue, strength))

    def gain(self, g: tuple[float, float, float], clamp: bool = True) -> None:
        from . import paragon_otf as PO

        self._do(lambda o: PO._gain(o, g, clamp))

    def sensor_noise(self, std: float, gen: D.PhiloxState, noise: Tensor | None = None) -> None:
        from . import paragon_otf as PO

        self._do(lambda o: PO.sensor_noise(o, std, noise, gen))

    def demosaic(self) -> None:
        from . import paragon_otf as PO

        self._do(PO.demosaic)

    def trunc8(self) -> None:
        from . import paragon_otf as PO

        self._do(PO.trunc8)

    def libjpeg(self, quality: int) -> None:
        from . import paragon_otf as PO

        self._do(lambda o: PO.jpeg_round(o, quality))
