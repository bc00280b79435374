"""Drop-in for the ``*_pt`` functions of traiNNer/data/degradations.py.

Gaussian noise  (degradations.py:569-633, :668-698), Poisson noise (:762-842, :879-909) and
``resize_pt`` (:958-1021).  Same names, positional arguments and defaults as the reference.
Keyword-only extras (not in the reference) let tests inject the random fields:
``noise=`` / ``noise_gray=`` for Gaussian, ``poisson_counts=`` / ``poisson_counts_gray=`` for
Poisson, and ``generator=`` — a :class:`PhiloxState` holding the (seed, offset) the kernels use.

Host-side draw order is the reference's: ``torch.rand(B)`` for
sigma/scale, ``torch.rand(B)`` for the gray flags — both on the device's torch generator — and
only the bulk fields (randn/poisson) come from the kernels' Philox streams.
"""

from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
from torch import Tensor

from . import _lib


class PhiloxState:
    """(seed, offset) for the kernels' Philox4x32-10 streams; ``offset`` advances once per call,
    so a fixed seed reproduces a run."""

    def __init__(self, seed: int = 0, offset: int = 0) -> None:
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = int(offset)

    def next_offset(self) -> int:
        o = self.offset
        self.offset += 1
        return o

    def manual_seed(self, seed: int) -> "PhiloxState":
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = 0
        return self


_default_generator = PhiloxState(0x0B200B200)


def default_generator() -> PhiloxState:
    return _default_generator


def _flags(clip: bool, rounds: bool) -> int:
    return (_lib.NOISE_CLIP if clip else 0) | (_lib.NOISE_ROUNDS if rounds else 0)


def _per_sample(v: float | Tensor, b: int, device: torch.device) -> Tensor:
    if isinstance(v, (int, float)):
        return torch.full((b,), float(v), dtype=torch.float32, device=device)
    return v.reshape(b).to(device=device, dtype=torch.float32).contiguous()


# ------------------------------------------------------------------ Gaussian ----


def add_noise_field_pt(img: Tensor, field: Tensor, clip: bool = True, rounds: bool = False) -> Tensor:
    """``tail(img + field)`` — the last three statements of ``add_gaussian_noise_pt`` / ``add_poisson_noise_pt``
    (degradations.py:625-632, :833-841) for a noise field that was generated elsewhere (e.g. by the reference, in a
    parity test that injects the finished field)."""
    _lib.require_cuda(img, field)
    x = _lib.dense_f32(img)
    b, c, h, w = x.shape
    f = field.to(torch.float32).expand(b, c, h, w).contiguous()
    out = torch.empty_like(x)
    _lib.call("otf_gaussian_noise_f32", _lib.ptr(x), b, c, h, w, None, None, _lib.ptr(f), None, 0, 0, None,
              _flags(clip, rounds) | _lib.NOISE_RAW_FIELD, _lib.ptr(out), _lib.stream())
    return out


def _gaussian(img, sigma, gray_noise, clip, rounds, add, noise=None, noise_gray=None, generator=None) -> Tensor:
    _lib.require_cuda(img)
    x = _lib.dense_f32(img)
    b, c, h, w = x.shape
    sg = _per_sample(sigma, b, x.device)
    if isinstance(gray_noise, (int, float)):
        gray = _per_sample(gray_noise, b, x.device) if gray_noise > 0 else None
    else:
        gray = _per_sample(gray_noise, b, x.device)
    if noise is not None:
        noise = noise.to(torch.float32).contiguous()
        if noise_gray is not None:
            noise_gray = noise_gray.to(torch.float32).contiguous()
        else:
            gray = None  # the reference draws no gray field when no flag is set (:592-598)
    gen = generator or _default_generator
    out = torch.empty_like(x)
    _lib.call(
        "otf_gaussian_noise_f32", _lib.ptr(x), b, c, h, w, _lib.ptr(sg), _lib.ptr(gray), _lib.ptr(noise),
        _lib.ptr(noise_gray), gen.seed, gen.next_offset(), None, _flags(clip, rounds) if add else _lib.NOISE_FIELD_ONLY,
        _lib.ptr(out), _lib.stream(),
    )
    return out


def generate_gaussian_noise_pt(img: Tensor, sigma: float | Tensor = 10, gray_noise: float | Tensor = 0, *,
                               noise: Tensor | None = None, noise_gray: Tensor | None = None,
                               generator: PhiloxState | None = None) -> Tensor:
    """degradations.py:569-605 — returns the noise field only."""
    return _gaussian(img, sigma, gray_noise, False, False, False, noise, noise_gray, generator)


def add_gaussian_noise_pt(img: Tensor, sigma: float | Tensor = 10, gray_noise: float | Tensor = 0, clip: bool = True,
                          rounds: bool = False, *, noise: Tensor | None = None, noise_gray: Tensor | None = None,
                          generator: PhiloxState | None = None) -> Tensor:
    """degradations.py:608-633."""
    return _gaussian(img, sigma, gray_noise, clip, rounds, True, noise, noise_gray, generator)


def _draw_range_and_gray(img: Tensor, rng: tuple[float, float], gray_prob: float | Tensor) -> tuple[Tensor, Tensor]:
    # degradations.py:673-679 / :884-890: two torch.rand(B) draws on the image's device
    val = torch.rand(img.size(0), dtype=img.dtype, device=img.device) * (rng[1] - rng[0]) + rng[0]
    gray = torch.rand(img.size(0), dtype=img.dtype, device=img.device)
    gray = (gray < gray_prob).float()
    return val, gray


def random_generate_gaussian_noise_pt(img: Tensor, sigma_range: tuple[float, float] = (0, 10),
                                      gray_prob: float | Tensor = 0, *, generator: PhiloxState | None = None) -> Tensor:
    """degradations.py:668-680."""
    sigma, gray = _draw_range_and_gray(img, sigma_range, gray_prob)
    return generate_gaussian_noise_pt(img, sigma, gray, generator=generator)


def random_add_gaussian_noise_pt(img: Tensor, sigma_range: tuple[float, float] = (0, 1.0), gray_prob: float | Tensor = 0,
                                 clip: bool = True, rounds: bool = False, *,
                                 generator: PhiloxState | None = None) -> Tensor:
    """degradations.py:683-698."""
    sigma, gray = _draw_range_and_gray(img, sigma_range, gray_prob)
    return add_gaussian_noise_pt(img, sigma, gray, clip, rounds, generator=generator)


# ------------------------------------------------------------------- Poisson ----


# Universal Poisson CDF tables (2.96 MB, one per device, read-only once built): lambda = level/255 * 2^v can take only
# 2304 values, so the table-inversion sampler never rebuilds anything.  Same stream discipline as the resize tables:
# the entry remembers the event that follows the build; other streams wait on it; a capturing stream only uses tables
# whose build has completed (poisson_tables() is called eagerly by RealESRGANFeed.__init__).
_POISSON_TABLES: dict[int, tuple[Tensor, "torch.cuda.Event", int, bool]] = {}


def poisson_tables(device: torch.device, wait: bool = False) -> Tensor | None:
    """The device's table block, building it on first use.  ``wait=True`` (RealESRGANFeed.__init__) blocks until the
    build has finished so that later CUDA-graph captures can use it; inside a capture an unfinished table is never
    touched (event queries are illegal there): the call returns None and the rejection sampler runs instead."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    hit = _POISSON_TABLES.get(idx)
    if hit is not None and hit[3]:
        return hit[0]
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing:
        return None
    cur = torch.cuda.current_stream(device)
    if hit is None:
        tab = torch.empty(_lib.load().otf_poisson_tables_bytes() // 4, dtype=torch.int32, device=device)
        _lib.call("otf_poisson_build_tables", _lib.ptr(tab), _lib.stream())
        ev = torch.cuda.Event()
        ev.record(cur)
        hit = (tab, ev, cur.cuda_stream, False)
        _POISSON_TABLES[idx] = hit
    tab, ev, sid, _ = hit
    if wait:
        ev.synchronize()
    if ev.query():
        _POISSON_TABLES[idx] = (tab, ev, sid, True)
    elif sid != cur.cuda_stream:
        cur.wait_event(ev)
    return tab


def _poisson(img, scale, gray_noise, clip, rounds, add, counts=None, counts_gray=None, generator=None,
             export: dict | None = None) -> Tensor:
    _lib.require_cuda(img)
    x = _lib.dense_f32(img)
    b, c, h, w = x.shape
    sc = _per_sample(scale, b, x.device)
    if isinstance(gray_noise, (int, float)):
        gray = _per_sample(gray_noise, b, x.device) if gray_noise > 0 else None
    else:
        gray = _per_sample(gray_noise, b, x.device)
    if counts is not None:
        counts = counts.to(torch.float32).contiguous()
        if counts_gray is not None:
            counts_gray = counts_gray.to(torch.float32).contiguous()
        else:
            gray = None  # reference skips the gray branch when no flag is set (:784-796)
    masks = torch.empty(b * 16, dtype=torch.int32, device=x.device)
    # counts drawn on the device: exact table inversion (the tables are built once per device)
    tables = poisson_tables(x.device) if counts is None and counts_gray is None and export is None else None
    vals = lam_c = lam_g = None
    if export is not None:
        vals = torch.empty(b, 2, dtype=torch.float32, device=x.device)
        lam_c = torch.empty_like(x)
        lam_g = torch.empty((b, 1, h, w), dtype=torch.float32, device=x.device) if gray is not None else None
    gen = generator or _default_generator
    out = torch.empty_like(x)
    _lib.call(
        "otf_poisson_noise_f32", _lib.ptr(x), b, c, h, w, _lib.ptr(sc), _lib.ptr(gray), _lib.ptr(counts),
        _lib.ptr(counts_gray), gen.seed, gen.next_offset(), None, _flags(clip, rounds) if add else _lib.NOISE_FIELD_ONLY, _lib.ptr(masks),
        _lib.ptr(tables), _lib.ptr(vals), _lib.ptr(lam_c), _lib.ptr(lam_g), _lib.ptr(out), _lib.stream(),
    )
    if export is not None:
        export.update(vals=vals, lambda_color=lam_c, lambda_gray=lam_g)
    return out


def generate_poisson_noise_pt(img: Tensor, scale: float | Tensor = 1.0, gray_noise: float | Tensor = 0, *,
                              poisson_counts: Tensor | None = None, poisson_counts_gray: Tensor | None = None,
                              generator: PhiloxState | None = None) -> Tensor:
    """degradations.py:762-811 — returns the noise field only."""
    return _poisson(img, scale, gray_noise, False, False, False, poisson_counts, poisson_counts_gray, generator)


def add_poisson_noise_pt(img: Tensor, scale: float = 1.0, clip: bool = True, rounds: bool = False,
                         gray_noise: float | Tensor = 0, *, poisson_counts: Tensor | None = None,
                         poisson_counts_gray: Tensor | None = None, generator: PhiloxState | None = None,
                         _export: dict | None = None) -> Tensor:
    """degradations.py:814-842 (note the reference's argument order: clip, rounds, gray_noise)."""
    return _poisson(img, scale, gray_noise, clip, rounds, True, poisson_counts, poisson_counts_gray, generator, _export)


def random_generate_poisson_noise_pt(img: Tensor, scale_range: tuple[float, float] = (0, 1.0),
                                     gray_prob: float | Tensor = 0, *, generator: PhiloxState | None = None) -> Tensor:
    """degradations.py:879-891."""
    scale, gray = _draw_range_and_gray(img, scale_range, gray_prob)
    return generate_poisson_noise_pt(img, scale, gray, generator=generator)


def random_add_poisson_noise_pt(img: Tensor, scale_range: tuple[float, float] = (0, 1.0), gray_prob: float | Tensor = 0,
                                clip: bool = True, rounds: bool = False, *,
                                generator: PhiloxState | None = None) -> Tensor:
    """degradations.py:894-909."""
    scale, gray = _draw_range_and_gray(img, scale_range, gray_prob)
    return add_poisson_noise_pt(img, scale, clip, rounds, gray, generator=generator)


# -------------------------------------------------------------------- resize ----

ANTIALIAS_MODES = {"bicubic", "bilinear"}
_MODE_ID = {
    "bilinear": _lib.RESIZE_BILINEAR_AA,
    "bicubic": _lib.RESIZE_BICUBIC_AA,
    "area": _lib.RESIZE_AREA,
    "nearest-exact": _lib.RESIZE_NEAREST_EXACT,
}


def _lanczos_taps(ratio: float, a: int = 3) -> np.ndarray:
    """Host-side taps of degradations.py:961-978: positions accumulated as the reference does
    (python float sum stored to fp32), sinc(t)*sinc(t/a) in fp32, normalised."""
    n = math.ceil(a / ratio + 1)
    half = np.empty(n, dtype=np.float32)
    cur = 0.0
    for i in range(n):
        half[i] = cur
        cur += ratio
    pos = np.concatenate([-half[1:][::-1], half])[1:-1].astype(np.float32)

    def sinc(t: np.ndarray) -> np.ndarray:
        pt = (np.float32(math.pi) * t).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = (np.sin(pt) / pt).astype(np.float32)
        return np.where(t != 0, v, np.float32(1.0)).astype(np.float32)

    inside = np.logical_and(-a < pos, pos < a)
    w = np.where(inside, sinc(pos) * sinc((pos / np.float32(a)).astype(np.float32)), np.float32(0)).astype(np.float32)
    return np.ascontiguousarray((w / w.sum(dtype=np.float32)).astype(np.float32))


# Resize weight tables depend on (H, W, OH, OW, mode) only: keep the last few per device.  An entry
# remembers the event that follows the launch which filled it; another stream waits on that event, and
# a capturing stream only takes entries pinned beforehand with pin_resize_tables() (a captured graph
# keeps reading the same memory, so pinned tables live for the life of the process).
_TABLE_CACHE: dict[tuple, tuple[Tensor, "torch.cuda.Event", int]] = {}
_TABLE_PINNED: dict[tuple, Tensor] = {}
_TABLE_CACHE_MAX = 4096  # a few KB each; random schedules revisit a few hundred (H, W, OH, OW, mode) keys per stage


def pin_resize_tables() -> int:
    """Wait for every cached table to be complete and pin it, so that a CUDA-graph capture that follows
    can read it (captured graphs keep reading the same memory, so pinned tables are never evicted)."""
    for key, (tab, ev, _) in list(_TABLE_CACHE.items()):
        ev.synchronize()
        _TABLE_PINNED[key] = tab
    return len(_TABLE_PINNED)


def pinned_resize_table(device: torch.device, h: int, w: int, oh: int, ow: int, mode_id: int) -> Tensor:
    """The weight tables of one (H, W, OH, OW, mode), built now if need be and kept for the life of the process — for
    chains that are about to be captured into a CUDA graph (the capture starts with a device synchronise, so a table
    launched here is complete before any captured kernel can read it)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), h, w, oh, ow, mode_id)
    tab = _TABLE_PINNED.get(key)
    if tab is None:
        hit = _TABLE_CACHE.get(key)
        if hit is not None:
            hit[1].synchronize()
            tab = hit[0]
        else:
            nbytes = _lib.load().otf_resize_workspace_bytes(h, w, oh, ow, mode_id)
            if nbytes <= 0:
                raise _lib.OtfError(f"resize tables: mode {mode_id} or extents ({h}, {w}) -> ({oh}, {ow}) not supported")
            tab = torch.empty(nbytes // 4, dtype=torch.int32, device=device)
            _lib.call("otf_resize_tables_f32", h, w, oh, ow, mode_id, _lib.ptr(tab), nbytes, _lib.stream())
            torch.cuda.current_stream().synchronize()
        _TABLE_PINNED[key] = tab
    return tab


def _resize_call(x: Tensor, oh: int, ow: int, mode_id: int, clamp: bool) -> Tensor:
    b, c, h, w = x.shape
    out = torch.empty((b, c, oh, ow), dtype=torch.float32, device=x.device)
    key = (x.device.index, h, w, oh, ow, mode_id)
    cur = torch.cuda.current_stream()
    capturing = torch.cuda.is_current_stream_capturing()
    ws = _TABLE_PINNED.get(key)
    if ws is None:
        hit = _TABLE_CACHE.get(key)
        if hit is not None:
            tab, ev, sid = hit
            if not capturing:  # (event queries are illegal under capture: only pinned tables are used there)
                if sid != cur.cuda_stream:
                    cur.wait_event(ev)
                    tab.record_stream(cur)  # an eviction must not hand the block back while this stream still reads it
                ws = tab
    ready = ws is not None
    if ws is None:
        ws_bytes = _lib.load().otf_resize_workspace_bytes(h, w, oh, ow, mode_id)
        ws = torch.empty(ws_bytes // 4, dtype=torch.int32, device=x.device)
    _lib.call("otf_resize_f32", _lib.ptr(x), b * c, h, w, _lib.ptr(out), oh, ow, mode_id, int(clamp), _lib.ptr(ws),
              ws.numel() * 4, int(ready), _lib.stream(), launches=1 if ready else 2)
    if not ready and not capturing:
        ev = torch.cuda.Event()
        ev.record(cur)
        if len(_TABLE_CACHE) >= _TABLE_CACHE_MAX:
            _TABLE_CACHE.pop(next(iter(_TABLE_CACHE)))
        _TABLE_CACHE[key] = (ws, ev, cur.cuda_stream)
    return out


def resize_pt(img: Tensor, mode: str, scale_factor: float = 0, size: tuple[int, int] = (0, 0)) -> Tensor:
    """degradations.py:1004-1021: ``size`` wins, ``scale_factor`` is turned into a size with Python's
    ``round``; bilinear/bicubic are antialiased; every mode ends in ``clamp(0, 1)``."""
    if scale_factor == 0 and tuple(size) == (0, 0):
        raise ValueError("scale_factor or size is required")
    if scale_factor != 0:
        size = (round(img.shape[2] * scale_factor), round(img.shape[3] * scale_factor))
    _lib.require_cuda(img)
    x = _lib.dense_f32(img)
    b, c, h, w = x.shape
    oh, ow = int(size[0]), int(size[1])
    if mode == "lanczos" and _lib.load().otf_resize_workspace_bytes(h, w, oh, ow, _lib.RESIZE_LANCZOS) > 0:
        # degradations.py:961-1001 in ONE pass: the Lanczos prefilter of the shrinking axes composed with the bicubic sample
        return _resize_call(x, oh, ow, _lib.RESIZE_LANCZOS, True)
    if mode == "lanczos":
        # extreme down-scales (prefilter radius > 62): prefilter passes, then plain bicubic + clamp
        cur = x
        if oh < h:
            taps = _lanczos_taps(oh / h)
            nxt = torch.empty_like(cur)
            _lib.call("otf_sepconv_reflect_f32", _lib.ptr(cur), b * c, h, w, taps.ctypes.data_as(C.c_void_p), len(taps), 0,
                      _lib.ptr(nxt), _lib.stream())
            cur = nxt
        if ow < w:
            taps = _lanczos_taps(ow / w)
            nxt = torch.empty_like(cur)
            _lib.call("otf_sepconv_reflect_f32", _lib.ptr(cur), b * c, h, w, taps.ctypes.data_as(C.c_void_p), len(taps), 1,
                      _lib.ptr(nxt), _lib.stream())
            cur = nxt
        return _resize_call(cur, oh, ow, _lib.RESIZE_BICUBIC, True)
    if mode not in _MODE_ID:
        raise NotImplementedError(f"resize_pt: unsupported mode {mode!r}")
    return _resize_call(x, oh, ow, _MODE_ID[mode], True)
