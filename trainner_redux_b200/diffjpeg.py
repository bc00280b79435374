"""Drop-in for traiNNer/utils/diffjpeg.py: the ``DiffJPEG`` module (diffjpeg.py:485-527).

One fused sm_100a kernel replaces the 12 sub-modules; the quantisation tables and DCT
basis live in the kernel's constant bank.
"""

from __future__ import annotations

import torch
from torch import Tensor, nn

from . import _lib


def quality_to_factor(quality: float) -> float:
    """diffjpeg.py:48-61 for Python numbers."""
    quality = 5000.0 / quality if quality < 50 else 200.0 - quality * 2
    return quality / 100.0


class DiffJPEG(nn.Module):
    """Batched JPEG simulation. ``differentiable=True`` selects the reference's cubic rounding
    surrogate (forward values only — this path runs under no_grad in feed_data).

    The reference module is an autograd graph (diffjpeg.py:485-527); this kernel is forward-only, so a call that
    would need a gradient — grad mode on and an input that requires grad — raises instead of silently returning a
    tensor with no ``grad_fn``."""

    def __init__(self, differentiable: bool = True) -> None:
        super().__init__()
        self.differentiable = bool(differentiable)

    def forward(self, x: Tensor, quality: float | Tensor, *, _clamp_in: bool = False, _round8: bool = False,
                _keep_quality: bool = False) -> Tensor:
        """
        Args:
            x (Tensor): Input image, bchw, rgb, [0, 1]
            quality (float | Tensor[b]): JPEG quality. A tensor is overwritten in place with the
                per-sample factors, as the reference does (diffjpeg.py:512-514).
        """
        _lib.require_cuda(x)
        if torch.is_grad_enabled() and (x.requires_grad or (isinstance(quality, Tensor) and quality.requires_grad)):
            raise RuntimeError(
                "trainner_redux_b200.DiffJPEG is forward-only (the OTF feed runs it under torch.no_grad, "
                "realesrgan_model.py:455); call it under torch.no_grad() or detach the input — it cannot backpropagate")
        img = _lib.dense_f32(x)
        b, c, h, w = img.shape
        if c != 3:
            raise RuntimeError(f"DiffJPEG expects 3 channels, got {c}")
        out = torch.empty_like(img)
        raw = 0
        if isinstance(quality, (int, float)):
            fac_t, fac_s = None, float(quality_to_factor(quality))
        else:
            _lib.require_cuda(quality)
            if quality.dtype != torch.float32 or not quality.is_contiguous() or quality.numel() != b:
                raise RuntimeError("quality must be a contiguous fp32 tensor of shape (b,)")
            if _keep_quality:
                raw = 1  # the kernel converts quality -> factor itself; the caller's tensor is left alone
            else:
                _lib.call("otf_quality_to_factor_f32", _lib.ptr(quality), b, _lib.stream())  # quirk Q1: in place
            fac_t, fac_s = quality, 0.0
        _lib.call(
            "otf_diffjpeg_f32", _lib.ptr(img), b, h, w, _lib.ptr(fac_t), fac_s, raw, int(self.differentiable),
            int(_clamp_in), int(_round8), _lib.ptr(out), _lib.stream(),
        )
        return out
