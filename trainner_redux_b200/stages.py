"""Builder for the native stage executor (``otf_run_stages_f32``, include/otf_b200.h).

``feed_data`` decides in Python what runs — it mirrors the reference's branches and host random draws
(traiNNer/models/realesrgan_model.py:455-650) — and records each stage here instead of launching it;
``run()`` then hands the whole list to the library, which launches every kernel back to back from C++.
One interpreter → library crossing per batch instead of ~20: with a fresh random plan per iteration the
eager path costs ≈0.65 ms of host time per step whatever the batch size (profiles/host_overhead.py),
more than the GPU time of the chain at the reference's usual batch sizes.

The stages are the same entry points with the same arguments as the per-primitive Python API
(img_process_util.filter2d, degradations.resize_pt / add_*_noise_pt, DiffJPEG, transforms.crop_pair),
so a chain run through here is bit-identical to the same chain run call by call
(tests/test_chain_native_gpu.py).
"""

from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib
from . import degradations as D

JPEG_IS_QUALITY, JPEG_DIFFERENTIABLE, JPEG_CLAMP_IN, JPEG_ROUND8 = 1, 2, 4, 8
_MISSED_ONCE: set[tuple] = set()  # resize-table keys that missed the cache once (see StageList._resize)


class StageList:
    """Accumulates ``OtfStage`` records for one batch; tracks the image extent as stages are added."""

    def __init__(self, img: Tensor, params: Any | None = None) -> None:
        """``params`` (a ``chain_graph.ParamBlock``): per-sample vectors, the Philox offset and the crop offsets are read
        by the kernels from that device block instead of being baked into the launches — the form a chain takes when it
        is captured into a CUDA graph (the caller fills and uploads the block before every run / replay)."""
        _lib.require_cuda(img)
        self.params = params
        if params is not None:
            params.begin()
        self.img = _lib.dense_f32(img)
        self.b, self.c, self.h, self.w = self.img.shape
        self.device = self.img.device
        self.stages: list[_lib.Stage] = []
        self.launches = 0
        self._keep: list[Any] = []  # tensors / host arrays the stages point into
        self._host_vecs: list[Tensor] = []  # per-sample CPU vectors, uploaded together by run()
        self._host_slots: list[tuple[int, str]] = []  # (stage index, field) to patch with the uploaded address
        self._new_tables: list[tuple[tuple, Tensor]] = []  # resize weight tables this run builds (published by run())
        self.unit_range = False  # the running image is known to lie in [0, 1] (after a clamping stage)

    # -- helpers ------------------------------------------------------------------------------
    def _add(self, op: int, **kw: Any) -> _lib.Stage:
        s = _lib.Stage()
        s.op = op
        for k, v in kw.items():
            setattr(s, k, v)
        self.stages.append(s)
        return s

    def _dev_ptr(self, t: Tensor | None) -> int | None:
        if t is None:
            return None
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        _lib.require_cuda(t)
        self._keep.append(t)
        return t.data_ptr()

    def _per_sample(self, stage: _lib.Stage, fld: str, v: float | Tensor | None) -> None:
        """Point ``stage.fld`` at a length-B fp32 device vector: device tensors are used in place, CPU tensors
        and scalars are collected and uploaded by ONE copy in run()."""
        if v is None:
            return
        if self.params is not None:
            if isinstance(v, Tensor) and v.is_cuda:
                raise RuntimeError("a chain with a parameter block takes its per-sample vectors from the host")
            setattr(stage, fld, self.params.add_row(v))
            return
        if isinstance(v, Tensor) and v.is_cuda:
            t = v.reshape(self.b).to(torch.float32).contiguous()
            self._keep.append(t)
            setattr(stage, fld, t.data_ptr())
            return
        t = torch.full((self.b,), float(v), dtype=torch.float32) if isinstance(v, (int, float)) else v.reshape(self.b).to(torch.float32)
        self._host_slots.append((len(self.stages) - 1, fld))
        self._host_vecs.append(t)

    # -- stages (each mirrors the per-primitive wrapper it replaces) ----------------------------
    def analyse(self, kernels: Sequence[Tensor]) -> None:
        """Joint analysis of up to four kernel tensors (img_process_util.KernelAnalysis)."""
        kb, k = kernels[0].size(0), kernels[0].size(-1)
        assert 1 <= len(kernels) <= 4 and all(t.size(0) == kb and t.size(-1) == k for t in kernels)
        ptrs = [self._dev_ptr(t) for t in kernels]
        self._sets = list(ptrs)
        self._sets_shape = (kb, k)
        self._add(_lib.OP_ANALYSE, n=len(kernels), kb=kb, K=k, **{f"p{i}": p for i, p in enumerate(ptrs)})
        self.launches += 1

    def filter2d(self, kernel: Tensor, analysed_set: int | None = None) -> None:
        k, kb = kernel.size(-1), kernel.size(0)
        if k % 2 != 1:
            raise ValueError("Wrong kernel size")
        if kb not in (1, self.b):
            raise RuntimeError(f"kernel batch {kb} does not match image batch {self.b}")
        if k // 2 >= self.h or k // 2 >= self.w:
            raise RuntimeError(
                f"Padding size should be less than the corresponding input dimension, but got: padding ({k // 2}, {k // 2}) "
                f"at dimension 3 of input {[self.b, self.c, self.h, self.w]}")
        if analysed_set is not None and k <= 21:
            self._add(_lib.OP_FILTER2D, p0=self._sets[analysed_set], kb=kb, K=k, n=analysed_set)
            self.launches += 1
            self.unit_range = False
        else:
            self._add(_lib.OP_FILTER2D, p0=self._dev_ptr(kernel), kb=kb, K=k, n=-1)
            self.launches += 1 if k > 21 else 2
            self.unit_range = False

    def at(self, name: str) -> "StageList":
        """Stage name (used by the per-stage executor, stage_by_stage.StageByStage); nothing to record here."""
        return self

    def usm(self, taps: Any, weight: float, threshold: float) -> None:
        taps = getattr(taps, "_taps", taps)  # a USMSharp module or its 1-D taps
        n = len(taps)
        if n // 2 >= self.h or n // 2 >= self.w:
            raise RuntimeError(f"Padding size should be less than the corresponding input dimension, but got: padding ({n // 2}, {n // 2})")
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        self._keep.append(taps)
        self._add(_lib.OP_USM, p0=taps.ctypes.data, n=n, f0=float(weight), f1=float(threshold))
        self.launches += 4
        self.unit_range = False

    def resize(self, mode: str, scale_factor: float = 0, size: tuple[int, int] = (0, 0)) -> None:
        """degradations.resize_pt: ``size`` wins, else round(extent * scale_factor); always clamps."""
        if scale_factor == 0 and tuple(size) == (0, 0):
            raise ValueError("scale_factor or size is required")
        if scale_factor != 0:
            size = (round(self.h * scale_factor), round(self.w * scale_factor))
        oh, ow = int(size[0]), int(size[1])
        if (oh, ow) == (self.h, self.w) and self.unit_range and (mode in D._MODE_ID or mode == "lanczos"):
            # A same-size resample is the identity in every mode of resize_pt — the window weights evaluate to exactly
            # (1, 0, ...) (cubic pieces vanish at 1 and 2 in fp32, area / nearest windows have one tap, the lanczos
            # prefilter only runs on shrinking axes) — and the trailing clamp(0, 1) is the identity on an image a clamping
            # stage produced: no launch (realesrgan_model.py:564-571 with ori_h // scale equal to the current extent).
            return
        if mode == "lanczos" and _lib.load().otf_resize_workspace_bytes(self.h, self.w, oh, ow, _lib.RESIZE_LANCZOS) > 0:
            mode_id = _lib.RESIZE_LANCZOS  # prefilter and bicubic sample composed into one set of weight tables: one launch
        elif mode == "lanczos":  # extreme down-scales: prefilter the shrinking axes, then plain bicubic (degradations.py:982-1001)
            for axis, (o, i) in enumerate(((oh, self.h), (ow, self.w))):
                if o < i:
                    taps = D._lanczos_taps(o / i)
                    self._keep.append(taps)
                    self._add(_lib.OP_SEPCONV, p0=taps.ctypes.data, n=len(taps), mode=axis)
                    self.launches += 1
                    self.unit_range = False
            mode_id = _lib.RESIZE_BICUBIC
        elif mode in D._MODE_ID:
            mode_id = D._MODE_ID[mode]
        else:
            raise NotImplementedError(f"resize_pt: unsupported mode {mode!r}")
        self._resize(mode_id, oh, ow, True)

    def _resize(self, mode_id: int, oh: int, ow: int, clamp: bool) -> None:
        """Weight tables depend on (H, W, OH, OW, mode) only and are shared with degradations._resize_call's cache:
        pinned tables are complete by construction and cached ones are ordered behind the event of the launch that
        filled them.  On a miss the executor builds the table in its scratch; only a key that misses twice (a
        shape-stable workload, not the usual freshly drawn scale) gets a table of its own, published by run()."""
        key = (self.device.index, self.h, self.w, oh, ow, mode_id)
        tab = D._TABLE_PINNED.get(key)
        if tab is None and self.params is not None:  # about to be captured: build the tables now, keep them for good
            tab = D.pinned_resize_table(self.device, self.h, self.w, oh, ow, mode_id)
        ready = tab is not None
        if tab is None and key in D._TABLE_CACHE and not torch.cuda.is_current_stream_capturing():
            tab, ev, sid = D._TABLE_CACHE[key]
            cur = torch.cuda.current_stream()
            if sid != cur.cuda_stream:
                cur.wait_event(ev)
                tab.record_stream(cur)  # an eviction must not hand the block back while this stream still reads it
            ready = True
        elif tab is None and key in _MISSED_ONCE and not torch.cuda.is_current_stream_capturing():
            nbytes = _lib.load().otf_resize_workspace_bytes(self.h, self.w, oh, ow, mode_id)
            if nbytes > 0:
                tab = torch.empty(nbytes // 4, dtype=torch.int32, device=self.device)
                self._new_tables.append((key, tab))
        elif tab is None:
            if len(_MISSED_ONCE) > 256:
                _MISSED_ONCE.clear()
            _MISSED_ONCE.add(key)
        if tab is not None:
            self._keep.append(tab)
        self._add(_lib.OP_RESIZE, mode=mode_id, oh=oh, ow=ow, flags=int(clamp) | (2 if ready else 0),
                  p0=None if tab is None else tab.data_ptr())
        self.launches += 1 if ready else 2
        self.h, self.w = oh, ow
        self.unit_range = bool(clamp)

    def gaussian_noise(self, sigma: float | Tensor, gray: float | Tensor | None, gen: D.PhiloxState, clip: bool = True,
                       rounds: bool = False, noise: Tensor | None = None, noise_gray: Tensor | None = None) -> None:
        """degradations.add_gaussian_noise_pt (same defaulting of the gray field as _gaussian)."""
        if isinstance(gray, (int, float)) and gray <= 0:
            gray = None
        if noise is not None and noise_gray is None:
            gray = None
        s = self._add(_lib.OP_GAUSS, seed=gen.seed, flags=D._flags(clip, rounds), **self._offset(gen),
                      p2=self._dev_ptr(noise), p3=self._dev_ptr(noise_gray) if noise is not None else None)
        self._per_sample(s, "p0", sigma)
        self._per_sample(s, "p1", gray)
        self.launches += 1
        self.unit_range = bool(clip)

    def _offset(self, gen: D.PhiloxState) -> dict:
        """Philox position of a noise stage: the generator's running offset, or — with a parameter block — the stage's
        index inside the chain plus the base the block carries (the caller advances the generator per run)."""
        if self.params is None:
            return {"offset": gen.next_offset()}
        return {"offset": self.params.next_noise_index(), "p4": self.params.offset_ptr}

    def noise_field(self, field: Tensor, clip: bool = True, rounds: bool = False) -> None:
        """degradations.add_noise_field_pt: ``tail(img + field)`` for a field generated elsewhere (parity tests)."""
        f = field.to(torch.float32).expand(self.b, self.c, self.h, self.w)
        self._add(_lib.OP_GAUSS, flags=D._flags(clip, rounds) | _lib.NOISE_RAW_FIELD, p2=self._dev_ptr(f))
        self.launches += 1
        self.unit_range = bool(clip)

    def poisson_noise(self, scale: float | Tensor, gray: float | Tensor | None, gen: D.PhiloxState, clip: bool = True,
                      rounds: bool = False, counts: Tensor | None = None, counts_gray: Tensor | None = None) -> None:
        """degradations.add_poisson_noise_pt."""
        if isinstance(gray, (int, float)) and gray <= 0:
            gray = None
        if counts is not None and counts_gray is None:
            gray = None
        tables = D.poisson_tables(self.device) if counts is None else None  # exact table inversion when the counts are drawn here
        if tables is not None:
            s = self._add(_lib.OP_POISSON, seed=gen.seed, flags=D._flags(clip, rounds) | 8, **self._offset(gen),
                          p2=tables.data_ptr())  # (raw table block: owned by degradations._POISSON_TABLES for the life of the process)
        else:
            s = self._add(_lib.OP_POISSON, seed=gen.seed, flags=D._flags(clip, rounds), **self._offset(gen),
                          p2=self._dev_ptr(counts), p3=self._dev_ptr(counts_gray) if counts is not None else None)
        self._per_sample(s, "p0", scale)
        self._per_sample(s, "p1", gray)
        self.launches += 2
        self.unit_range = bool(clip)

    def jpeg(self, quality: float | Tensor, differentiable: bool = False, clamp_in: bool = True, round8: bool = False) -> None:
        """DiffJPEG.forward(x, quality) with the conversion of the quality fused into the kernel."""
        flags = JPEG_IS_QUALITY | (JPEG_DIFFERENTIABLE if differentiable else 0) | (JPEG_CLAMP_IN if clamp_in else 0) | \
            (JPEG_ROUND8 if round8 else 0)
        if isinstance(quality, Tensor) or self.params is not None:  # (a parameter block carries scalars as rows too)
            s = self._add(_lib.OP_JPEG, flags=flags)
            self._per_sample(s, "p0", quality)
        else:
            self._add(_lib.OP_JPEG, flags=flags, f0=float(quality))
        self.launches += 1
        self.unit_range = True  # diffjpeg.py:476-479: clamp to [0, 255], then / 255

    def clamp_round(self) -> None:
        self._add(_lib.OP_CLAMP_ROUND)
        self.launches += 1
        self.unit_range = True

    # -- the fork's extra stages (paragon_otf.py): same entry points, same arguments ------------
    def warp(self, mode: int, p0: float) -> None:
        """paragon_otf._warp: lens distortion / rolling shutter / chromatic aberration."""
        self._add(_lib.OP_WARP, mode=mode, f0=float(p0))
        self.launches += 1
        self.unit_range = mode == _lib.WARP_CHROMA  # (the chromatic warp clamps; the others only interpolate)

    def taps_zero(self, kernel: np.ndarray, epilogue: int = 0, strength: float = 0.0) -> None:
        """paragon_otf._taps_zero: zero-padded correlation with a small host kernel (motion blur, oversharpen)."""
        k = int(kernel.shape[0])
        kh = np.ascontiguousarray(kernel, dtype=np.float32)
        self._keep.append(kh)
        self._add(_lib.OP_TAPS_ZERO, p0=kh.ctypes.data, K=k, flags=int(epilogue), f0=float(strength))
        self.launches += 1
        pad = k // 2
        self.h, self.w = self.h + 2 * pad - k + 1, self.w + 2 * pad - k + 1
        self.unit_range = epilogue == _lib.TAPS_OVERSHARPEN

    def gain(self, g: tuple[float, float, float], clamp: bool = True) -> None:
        """paragon_otf._gain: exposure / colour temperature."""
        self._add(_lib.OP_GAIN, f0=float(g[0]), f1=float(g[1]), f2=float(g[2]), flags=int(clamp))
        self.launches += 1
        self.unit_range = bool(clamp)

    def sensor_noise(self, std: float, gen: D.PhiloxState, noise: Tensor | None = None) -> None:
        """paragon_otf.sensor_noise: clamp(img + N * std, 0, 1)."""
        if noise is not None:
            if tuple(noise.shape) != (self.b, self.c, self.h, self.w):
                raise ValueError("sensor_noise: injected field must have the image's shape")
            self._add(_lib.OP_SENSOR, f0=float(std), p0=self._dev_ptr(noise))
        else:
            self._add(_lib.OP_SENSOR, f0=float(std), seed=gen.seed, offset=gen.next_offset())
        self.launches += 1
        self.unit_range = True

    def demosaic(self) -> None:
        if self.c != 3:
            raise RuntimeError(f"demosaic expects 3 channels, got {self.c}")
        self._add(_lib.OP_DEMOSAIC)
        self.launches += 1
        self.unit_range = True

    def trunc8(self) -> None:
        self._add(_lib.OP_TRUNC8)
        self.launches += 1
        self.unit_range = True

    def libjpeg(self, quality: int) -> None:
        """paragon_otf.jpeg_round: uint8 truncation + libjpeg's baseline round trip at ``quality`` + / 255 (two launches)."""
        if self.c != 3:
            raise RuntimeError(f"the JPEG round expects 3 channels, got {self.c}")
        self._add(_lib.OP_LIBJPEG, n=int(quality))
        self.launches += 2
        self.unit_range = True

    def resize_raw(self, mode_id: int, oh: int, ow: int, clamp: bool) -> None:
        """degradations._resize_call with an explicit mode id (the aliasing stage's legacy nearest, no clamp)."""
        self._resize(mode_id, int(oh), int(ow), clamp)

    # -- execution ----------------------------------------------------------------------------
    def run(self, crop: tuple[Tensor, int, int, int, int] | None = None, gt_view: bool = False) -> Tensor | tuple[Tensor | None, Tensor]:
        """Launch the recorded stages.  Returns the final image, or with ``crop=(gt, gt_patch, scale, top, left)``
        the (gt_crop, lq_crop) pair of transforms.crop_pair, cropped by the chain's last launch.  ``gt_view``: no GT
        copy — the first element is None and the caller slices ``gt`` itself (the reference's crop returns that view)."""
        if not self.stages:
            raise ValueError("empty stage list")
        outs: tuple[Tensor, ...]
        if crop is not None:
            gt, gt_patch, scale, top, left = crop
            p = gt_patch // scale
            gt_out = None if gt_view else torch.empty((self.b, self.c, p * scale, p * scale), dtype=torch.float32, device=self.device)
            lq_out = torch.empty((self.b, self.c, p, p), dtype=torch.float32, device=self.device)
            self._add(_lib.OP_CROP_PAIR, p0=self._dev_ptr(gt), p1=None if gt_out is None else gt_out.data_ptr(), p2=lq_out.data_ptr(), oh=top, ow=left, n=p, mode=scale,
                      p4=None if self.params is None else self.params.crop_ptr)
            self.launches += 1
            outs = (gt_out, lq_out)
        else:
            last = next(s for s in reversed(self.stages) if s.op != _lib.OP_ANALYSE)
            out = torch.empty((self.b, self.c, self.h, self.w), dtype=torch.float32, device=self.device)
            last.dst = out.data_ptr()
            outs = (out,)
        if self._host_vecs:  # every per-sample CPU vector of the batch in one upload
            packed = torch.stack(self._host_vecs).to(self.device, non_blocking=True)
            self._keep.append(packed)
            base = packed.data_ptr()
            for row, (idx, fld) in enumerate(self._host_slots):
                setattr(self.stages[idx], fld, base + 4 * self.b * row)
        n = len(self.stages)
        arr = (_lib.Stage * n)(*self.stages)
        lib = _lib.load()
        b, c, h0, w0 = self.img.shape
        ws_bytes = lib.otf_run_stages_workspace_bytes(b, c, h0, w0, arr, n)
        if ws_bytes < 0:
            raise _lib.OtfError(f"otf_run_stages_workspace_bytes failed ({ws_bytes}): {_lib.last_error()}")
        ws = torch.empty(max(ws_bytes, 4) // 4, dtype=torch.int32, device=self.device)
        _lib.call("otf_run_stages_f32", _lib.ptr(self.img), b, c, h0, w0, arr, n, _lib.ptr(ws), ws_bytes, None, None, _lib.stream(),
                  launches=0)
        # the executor fuses adjacent stages where it has a fused kernel (csrc/chain.cu): what it launched is its count,
        # self.launches stays the unfused figure (one entry point per stage)
        self.launched = lib.otf_run_stages_launches()
        _lib.launch_count += self.launched
        if self._new_tables:  # one event covers every table this run built
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(cur)
            for key, tab in self._new_tables:
                if len(D._TABLE_CACHE) >= D._TABLE_CACHE_MAX:
                    D._TABLE_CACHE.pop(next(iter(D._TABLE_CACHE)))
                D._TABLE_CACHE[key] = (tab, ev, cur.cuda_stream)
        return outs if crop is not None else outs[0]
