"""Drop-in for the tensor-math half of traiNNer/models/paragon_otf_degradations.py (``ParagonOTF``) — the extra
stages this fork's ``RealESRGANModel.feed_data`` runs around the classical primitives.

Two layers:
  * explicit-parameter functions (``lens_distortion(img, strength)`` ...): one library call each, no RNG;
  * ``ParagonOTF``: the reference's static methods with the reference's names, ``(img_tensor, opt)`` arguments,
    ``hasattr`` guards and draw order (numpy ``Generator`` gate first, then the parameters), so
    ``ParagonOTF.apply_lens_distortion(img, opt)`` can replace the reference call one for one.  The generators are
    the process-wide ones of ``set_rng`` (the reference's ``RNG.get_rng()`` / ``random`` singletons) unless ``rng=``
    is given.

Out of scope (host codecs): WebP / AVIF / HEIF / ffmpeg video rounds.  The cv2 Bayer demosaic IS
reproduced on the device (bit for bit), and so is the unified pipeline's "jpeg" choice: ``jpeg_round`` runs libjpeg's
baseline round trip — what ``PIL.Image.save(format="JPEG")`` + ``Image.open`` decode to — in integer arithmetic on the
device, bit for bit (csrc/libjpeg.cu; oracle/libjpeg_oracle.py is pinned against PIL itself).  Other formats pass
through, which is what the reference does when a codec plugin is missing (paragon_otf_degradations.py:125-134).
"""

from __future__ import annotations

import ctypes as C
import math
import random as _random
import warnings
from typing import Any

import numpy as np
import torch
from torch import Tensor

from . import _lib
from . import degradations as D

_RNG: Any = None


def set_rng(rng: Any) -> None:
    """Install the host generators (an object with ``.np`` numpy Generator, ``.py`` ``random.Random`` and
    ``.philox``) used when a ``ParagonOTF`` method is called without ``rng=``."""
    global _RNG
    _RNG = rng


def _rng(rng: Any) -> Any:
    if rng is not None:
        return rng
    if _RNG is None:
        raise RuntimeError("paragon_otf: no host generators installed (call set_rng(HostRNG(seed)) or pass rng=)")
    return _RNG


def _f32(img: Tensor) -> Tensor:
    return _lib.dense_f32(img)


# ------------------------------------------------------------- explicit stages ----
def _warp(img: Tensor, mode: int, p0: float) -> Tensor:
    x = _f32(img)
    b, c, h, w = x.shape
    out = torch.empty_like(x)
    _lib.call("otf_warp_f32", _lib.ptr(x), b, c, h, w, mode, float(p0), _lib.ptr(out), _lib.stream())
    return out


def lens_distortion(img: Tensor, strength: float) -> Tensor:
    """paragon_otf_degradations.py:297-342 (barrel / pincushion: r' = r (1 + k r^2), reflection padding)."""
    return _warp(img, _lib.WARP_LENS, strength)


def rolling_shutter(img: Tensor, strength: float) -> Tensor:
    """:417-455 (x' = x + strength * H / W * y)."""
    h, w = img.shape[2:4]
    return _warp(img, _lib.WARP_SHUTTER, strength * h / w)


def chromatic_aberration(img: Tensor) -> Tensor:
    """:485-523 / realesrgan_model.py:244-310 (R x1.001, B x0.999, zeros padding, clamp); non-RGB input is returned as is."""
    if img.size(1) != 3:
        return img
    return _warp(img, _lib.WARP_CHROMA, 0.0)


def motion_blur_kernel(kernel_size: int, angle: float) -> np.ndarray:
    """:276-294 as a host array (K x K fp32): ones within half a pixel of the line through the centre, normalised."""
    center = kernel_size // 2
    c, s = math.cos(math.radians(angle)), math.sin(math.radians(angle))
    i = np.arange(kernel_size, dtype=np.float64)[:, None] - center
    j = np.arange(kernel_size, dtype=np.float64)[None, :] - center
    k = (np.abs(i * c + j * s) < 0.5).astype(np.float32)
    return np.ascontiguousarray(k / k.sum(dtype=np.float32))


def _taps_zero(img: Tensor, kernel: np.ndarray, epilogue: int = _lib.TAPS_NONE, strength: float = 0.0) -> Tensor:
    x = _f32(img)
    b, c, h, w = x.shape
    k = kernel.shape[0]
    pad = k // 2
    out = torch.empty((b, c, h + 2 * pad - k + 1, w + 2 * pad - k + 1), dtype=torch.float32, device=x.device)
    kh = np.ascontiguousarray(kernel, dtype=np.float32)
    _lib.call("otf_taps_zero_f32", _lib.ptr(x), b * c, h, w, k, kh.ctypes.data_as(C.c_void_p), epilogue, float(strength),
              _lib.ptr(out), _lib.stream())
    return out


def motion_blur(img: Tensor, kernel_size: int, angle: float) -> Tensor:
    """:251-273 — zero-padded correlation with the line kernel (an even K grows the image by one, as the reference)."""
    return _taps_zero(img, motion_blur_kernel(kernel_size, angle))


_BOX5 = np.full((5, 5), np.float32(1.0) / np.float32(25.0), dtype=np.float32)


def oversharpen(img: Tensor, strength: float) -> Tensor:
    """:458-482 / realesrgan_model.py:193-242 — clamp(img + (img - box5(img)) * strength, 0, 1), one launch."""
    return _taps_zero(img, _BOX5, _lib.TAPS_OVERSHARPEN, strength)


def _gain(img: Tensor, g: tuple[float, float, float], clamp: bool = True) -> Tensor:
    x = _f32(img)
    b, c, h, w = x.shape
    out = torch.empty_like(x)
    _lib.call("otf_channel_gain_f32", _lib.ptr(x), b, c, h * w, float(g[0]), float(g[1]), float(g[2]), int(clamp), _lib.ptr(out),
              _lib.stream())
    return out


def exposure(img: Tensor, factor: float) -> Tensor:
    """:345-362 — clamp(img * factor, 0, 1)."""
    return _gain(img, (factor, factor, factor))


def color_temperature_gains(shift: float) -> tuple[float, float, float]:
    """Per-channel gains of :365-394 — warm: R x(1+0.3s), G x(1+0.1s); cool: B x(1-0.3s), G x(1-0.1s)."""
    if shift > 0:
        return (1 + shift * 0.3, 1 + shift * 0.1, 1.0)
    return (1.0, 1 - shift * 0.1, 1 - shift * 0.3)


def color_temperature(img: Tensor, shift: float) -> Tensor:
    """:365-394 — the gains above, then clamp."""
    if img.size(1) != 3:
        return img
    return _gain(img, color_temperature_gains(shift))


def sensor_noise(img: Tensor, std: float, noise: Tensor | None = None, generator: D.PhiloxState | None = None) -> Tensor:
    """:397-414 — clamp(img + N * std, 0, 1); N from the kernel's Philox stream unless ``noise`` injects it."""
    x = _f32(img)
    out = torch.empty_like(x)
    if noise is not None:
        noise = _f32(noise)
        if noise.shape != x.shape:
            raise ValueError("sensor_noise: injected field must have the image's shape")
        seed = off = 0
    else:
        gen = generator or D.default_generator()
        seed, off = gen.seed, gen.next_offset()
    _lib.call("otf_sensor_noise_f32", _lib.ptr(x), x.numel(), float(std), _lib.ptr(noise), seed, off, _lib.ptr(out), _lib.stream())
    return out


def aliasing(img: Tensor, scale: float) -> Tensor:
    """:555-572 / realesrgan_model.py:365-401 — legacy `nearest` down to int(h*s) x int(w*s), then back up."""
    h, w = img.shape[2:4]
    down = D._resize_call(_f32(img), int(h * scale), int(w * scale), _lib.RESIZE_NEAREST, False)
    return D._resize_call(down, h, w, _lib.RESIZE_NEAREST, False)


def demosaic(img: Tensor) -> Tensor:
    """:526-552 — Bayer mosaic of the uint8-truncated image + OpenCV's bilinear demosaic, bit for bit, in one launch
    (the reference does this on the host with cv2, image by image)."""
    x = _f32(img)
    b, c, h, w = x.shape
    if c != 3:
        raise RuntimeError(f"demosaic expects 3 channels, got {c}")
    out = torch.empty_like(x)
    _lib.call("otf_demosaic_f32", _lib.ptr(x), b, h, w, _lib.ptr(out), _lib.stream())
    return out


def trunc8(img: Tensor) -> Tensor:
    """``(img.clamp(0,1) * 255).astype(uint8) / 255`` — the truncation in front of every codec round (:114-115)."""
    x = _f32(img)
    out = torch.empty_like(x)
    _lib.call("otf_trunc8_f32", _lib.ptr(x), x.numel(), _lib.ptr(out), _lib.stream())
    return out


_WARNED: set[str] = set()


def jpeg_round(img: Tensor, quality: float) -> Tensor:
    """The JPEG round of `_compress_with_format` (:119-149) on the device, BIT FOR BIT what the reference gets from PIL:
    clamp, uint8 truncation, libjpeg's baseline round trip at ``int(quality)`` (4:2:0, Annex-K tables, integer DCT, fancy
    up-sampling — ``otf_libjpeg_roundtrip_f32``), ``/ 255``.  3-channel images only (the reference turns a 1-channel
    batch into RGB here; not reproduced)."""
    x = _f32(img)
    b, c, h, w = x.shape
    if c != 3:
        raise RuntimeError(f"the JPEG round expects 3 channels, got {c}")
    out = torch.empty_like(x)
    nbytes = _lib.load().otf_libjpeg_workspace_bytes(b, h, w)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _lib.call("otf_libjpeg_roundtrip_f32", _lib.ptr(x), b, h, w, int(quality), _lib.ptr(ws), nbytes, _lib.ptr(out), _lib.stream())
    return out


def codec_runs_jpeg(format_name: str, quality: float | None, fallback: str = "passthrough") -> bool:
    """Whether a drawn (format, quality) runs the JPEG round here (see compress_with_format); warns once per host codec
    that passes through."""
    if quality is None:
        return False
    if format_name == "jpeg" or (fallback == "jpeg" and format_name in ("webp", "avif", "heif")):
        return True
    if format_name not in _WARNED:
        _WARNED.add(format_name)
        warnings.warn(f"paragon_otf: {format_name!r} is a host codec outside the GPU path; the image passes through unchanged "
                      "(as the reference does when the codec plugin is missing); set codec_fallback='jpeg' to run a JPEG round "
                      "at the drawn quality instead", stacklevel=2)
    return False


def compress_with_format(img: Tensor, format_name: str, quality: float | None, fallback: str = "passthrough") -> Tensor:
    """`_compress_with_format` (:95-158) for one drawn (format, quality).  "jpeg" = ``jpeg_round``: the reference's own
    PIL / libjpeg result, bit for bit (parity, not a stand-in).  WebP / AVIF / HEIF are host codecs:
    ``fallback="passthrough"`` (default) returns the image unchanged with a one-time warning — what the reference does
    when the codec plugin is missing; ``fallback="jpeg"`` runs the JPEG round at the drawn quality instead, so the
    share of compressed batches stays what the option file asks for."""
    if codec_runs_jpeg(format_name, quality, fallback):
        return jpeg_round(img, quality)
    return img


class ParagonOTF:
    """Static methods with the reference's names and draw order (paragon_otf_degradations.py:35-572)."""

    @staticmethod
    def apply_motion_blur(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "motion_blur_prob"):
            return img_tensor
        r = _rng(rng)
        if r.np.uniform() >= opt.motion_blur_prob:
            return img_tensor
        ks = r.py.randint(opt.motion_blur_kernel_size[0], opt.motion_blur_kernel_size[1])
        angle = r.np.uniform(opt.motion_blur_angle_range[0], opt.motion_blur_angle_range[1])
        return motion_blur(img_tensor, ks, angle)

    @staticmethod
    def _create_motion_blur_kernel(kernel_size: int, angle: float) -> Tensor:
        return torch.from_numpy(motion_blur_kernel(kernel_size, angle)).unsqueeze(0).unsqueeze(0)

    @staticmethod
    def apply_lens_distortion(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "lens_distort_prob") or not hasattr(opt, "lens_distort_strength_range"):
            return img_tensor
        r = _rng(rng)
        if r.np.uniform() >= opt.lens_distort_prob:
            return img_tensor
        return lens_distortion(img_tensor, r.np.uniform(*opt.lens_distort_strength_range))

    @staticmethod
    def apply_exposure_errors(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "exposure_prob") or not hasattr(opt, "exposure_factor_range"):
            return img_tensor
        r = _rng(rng)
        if r.np.uniform() >= opt.exposure_prob:
            return img_tensor
        return exposure(img_tensor, r.np.uniform(*opt.exposure_factor_range))

    @staticmethod
    def apply_color_temperature_shift(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "color_temp_prob") or not hasattr(opt, "color_temp_shift_range"):
            return img_tensor
        r = _rng(rng)
        if r.np.uniform() >= opt.color_temp_prob:
            return img_tensor
        if img_tensor.size(1) != 3:
            return img_tensor
        return color_temperature(img_tensor, r.np.uniform(*opt.color_temp_shift_range))

    @staticmethod
    def apply_sensor_noise(img_tensor: Tensor, opt: Any, rng: Any = None, noise: Tensor | None = None) -> Tensor:
        if not hasattr(opt, "sensor_noise_prob") or not hasattr(opt, "sensor_noise_std_range"):
            return img_tensor
        r = _rng(rng)
        if r.np.uniform() >= opt.sensor_noise_prob:
            return img_tensor
        return sensor_noise(img_tensor, r.np.uniform(*opt.sensor_noise_std_range), noise, getattr(r, "philox", None))

    @staticmethod
    def apply_rolling_shutter(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "rolling_shutter_prob") or not hasattr(opt, "rolling_shutter_strength_range"):
            return img_tensor
        r = _rng(rng)
        if r.np.uniform() >= opt.rolling_shutter_prob:
            return img_tensor
        return rolling_shutter(img_tensor, r.np.uniform(*opt.rolling_shutter_strength_range))

    @staticmethod
    def apply_oversharpening(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "oversharpen_prob") or not hasattr(opt, "oversharpen_strength"):
            return img_tensor
        r = _rng(rng)
        if r.np.uniform() >= opt.oversharpen_prob:
            return img_tensor
        return oversharpen(img_tensor, r.np.uniform(opt.oversharpen_strength[0], opt.oversharpen_strength[1]))

    @staticmethod
    def apply_chromatic_aberration(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "chromatic_aberration_prob"):
            return img_tensor
        if _rng(rng).np.uniform() >= opt.chromatic_aberration_prob:
            return img_tensor
        return chromatic_aberration(img_tensor)

    @staticmethod
    def apply_demosaicing_artifacts(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "demosaic_prob"):
            return img_tensor
        if _rng(rng).np.uniform() >= opt.demosaic_prob:
            return img_tensor
        return demosaic(img_tensor)

    @staticmethod
    def apply_aliasing_artifacts(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        if not hasattr(opt, "aliasing_prob") or not hasattr(opt, "aliasing_scale_range"):
            return img_tensor
        r = _rng(rng)
        if r.np.uniform() >= opt.aliasing_prob:
            return img_tensor
        return aliasing(img_tensor, r.np.uniform(*opt.aliasing_scale_range))

    @staticmethod
    def _choose_compression_format(formats: list[str], weights: list[float], rng: Any = None) -> str:
        return str(_rng(rng).np.choice(formats, p=weights))

    @staticmethod
    def _compress_with_format(img_tensor: Tensor, format_name: str, opt: Any, round: int = 1, rng: Any = None) -> Tensor:  # noqa: A002
        attr = f"compression_{format_name}_range"
        if not hasattr(opt, attr):
            return img_tensor
        lo, hi = getattr(opt, attr)
        return compress_with_format(img_tensor, format_name, _rng(rng).np.uniform(lo, hi), getattr(opt, "codec_fallback", "passthrough"))

    @staticmethod
    def apply_realistic_compression_pipeline(img_tensor: Tensor, opt: Any, rng: Any = None) -> Tensor:
        """:39-87 — one format per round; the optional second round models platform recompression."""
        if not hasattr(opt, "compression_formats"):
            return img_tensor  # the legacy per-codec probabilities (:68-72) are host codecs only
        r = _rng(rng)
        fmt = ParagonOTF._choose_compression_format(list(opt.compression_formats), list(opt.compression_weights), r)
        img_tensor = ParagonOTF._compress_with_format(img_tensor, fmt, opt, 1, r)
        if r.np.uniform() < opt.recompression_prob:
            fmt = ParagonOTF._choose_compression_format(list(opt.recompression_formats), list(opt.recompression_weights), r)
            img_tensor = ParagonOTF._compress_with_format(img_tensor, fmt, opt, 2, r)
        return img_tensor


def default_host_rng(seed: int = 0) -> Any:
    """A minimal generator bundle for callers outside ``RealESRGANFeed``."""
    from types import SimpleNamespace

    return SimpleNamespace(np=np.random.default_rng(seed), py=_random.Random(seed), philox=D.PhiloxState(seed))
