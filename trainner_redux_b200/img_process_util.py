"""Drop-in for traiNNer/utils/img_process_util.py: ``filter2d`` and ``USMSharp``.

Same names, arguments, defaults and error behaviour as the reference
(img_process_util.py:8-32 and :35-55); the work runs in libotf_b200's sm_100a kernels.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib

_SMALL_GAUSSIAN = {
    1: [1.0],
    3: [0.25, 0.5, 0.25],
    5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
    7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
    9: [4 / 256, 13 / 256, 30 / 256, 51 / 256, 60 / 256, 51 / 256, 30 / 256, 13 / 256, 4 / 256],
}


def gaussian_kernel_1d(ksize: int, sigma: float = 0.0) -> np.ndarray:
    """What ``cv2.getGaussianKernel(ksize, sigma)`` returns (float64, shape (ksize,)) —
    the call made at img_process_util.py:41.  sigma<=0 means 0.3*((ksize-1)*0.5-1)+0.8, with
    OpenCV's fixed tables for odd ksize<=9."""
    if sigma <= 0 and ksize in _SMALL_GAUSSIAN:
        return np.asarray(_SMALL_GAUSSIAN[ksize], dtype=np.float64)
    s = sigma if sigma > 0 else 0.3 * ((ksize - 1) * 0.5 - 1.0) + 0.8
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * s * s))
    return k / k.sum()


class KernelAnalysis:
    """Device-side analysis (true support, launch order, rank-1 factors) of up to four kernel
    tensors, produced by ONE pair of launches; ``filter2d(..., _analysis=(ka, i))`` then skips its own."""

    def __init__(self, kernels: list[Tensor]) -> None:
        assert 1 <= len(kernels) <= 4
        kb, k = kernels[0].size(0), kernels[0].size(-1)
        assert all(t.size(0) == kb and t.size(-1) == k for t in kernels), "kernel sets must share batch and size"
        self.kernels = [t.to(torch.float32).contiguous() for t in kernels]
        _lib.require_cuda(*self.kernels)
        self.words = _lib.load().otf_filter2d_scratch_words(kb)
        self.scratch = torch.empty(len(kernels) * self.words, dtype=torch.int32, device=kernels[0].device)
        ptrs = (C.c_void_p * len(kernels))(*[t.data_ptr() for t in self.kernels])
        _lib.call("otf_filter2d_analyse_f32", ptrs, len(kernels), kb, k, _lib.ptr(self.scratch), _lib.stream(),
                  launches=1)

    def scratch_ptr(self, i: int) -> C.c_void_p:
        return C.c_void_p(self.scratch.data_ptr() + 4 * i * self.words)


def filter2d(img: Tensor, kernel: Tensor, *, _analysis: tuple[KernelAnalysis, int] | None = None) -> Tensor:
    """PyTorch version of cv2.filter2D (img_process_util.py:8-32).

    Args:
        img (Tensor): (b, c, h, w)
        kernel (Tensor): (b, k, k), or (1, k, k) to share one kernel across the batch
    """
    k = kernel.size(-1)
    if k % 2 != 1:
        raise ValueError("Wrong kernel size")
    _lib.require_cuda(img, kernel)
    x = _lib.dense_f32(img)
    b, c, h, w = x.shape
    kb = kernel.size(0)
    if kb not in (1, b):
        raise RuntimeError(f"kernel batch {kb} does not match image batch {b}")
    if k // 2 >= h or k // 2 >= w:
        # F.pad(mode="reflect") raises the same way at img_process_util.py:18
        raise RuntimeError(
            f"Padding size should be less than the corresponding input dimension, but got: padding ({k // 2}, {k // 2}) "
            f"at dimension 3 of input {list(img.shape)}"
        )
    out = torch.empty_like(x)
    if _analysis is not None and k <= 21:
        ka, i = _analysis
        _lib.call("otf_filter2d_f32", _lib.ptr(x), b, c, h, w, _lib.ptr(ka.kernels[i]), kb, k, ka.scratch_ptr(i), 1, _lib.ptr(out),
                  _lib.stream(), launches=1)
        return out
    kern = kernel.to(torch.float32).contiguous()
    # per-kernel analysis scratch: true radii, launch order, rank-1 flags and factors
    scratch = torch.empty(_lib.load().otf_filter2d_scratch_words(kb), dtype=torch.int32, device=x.device)
    _lib.call("otf_filter2d_f32", _lib.ptr(x), b, c, h, w, _lib.ptr(kern), kb, k, _lib.ptr(scratch), 0, _lib.ptr(out), _lib.stream(),
              launches=1 if k > 21 else 2)
    return out


class USMSharp(nn.Module):
    """Unsharp-mask sharpening, same constructor and forward as img_process_util.py:35-55.

    The ``kernel`` buffer is kept (shape (1,K,K), fp32) for state-dict compatibility; the
    computation uses its exact rank-1 factors (51+51 taps instead of 2601)."""

    def __init__(self, radius: int = 50, sigma: int = 0) -> None:
        super().__init__()
        if radius % 2 == 0:
            radius += 1
        self.radius = radius
        k1 = gaussian_kernel_1d(radius, sigma)
        kernel = torch.FloatTensor(np.dot(k1.reshape(-1, 1), k1.reshape(1, -1))).unsqueeze_(0)
        self.register_buffer("kernel", kernel)
        self._taps = np.ascontiguousarray(k1.astype(np.float32))

    def forward(self, img: Tensor, weight: float = 0.5, threshold: int = 10) -> Tensor:
        _lib.require_cuda(img)
        x = _lib.dense_f32(img)
        b, c, h, w = x.shape
        n = len(self._taps)
        if n // 2 >= h or n // 2 >= w:
            raise RuntimeError(
                f"Padding size should be less than the corresponding input dimension, but got: padding ({n // 2}, {n // 2}) "
                f"at dimension 3 of input {list(img.shape)}"
            )
        if n > 1023:
            raise _lib.OtfError("USMSharp: radius above 1023 is not supported by the sm_100a kernel")
        lib = _lib.load()
        ws_bytes = lib.otf_usm_workspace_bytes(b * c, h, w)
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=x.device)
        out = torch.empty_like(x)
        _lib.call(
            "otf_usm_sharp_f32", _lib.ptr(x), b * c, h, w, self._taps.ctypes.data_as(C.c_void_p), n,
            float(weight), float(threshold), _lib.ptr(ws), ws_bytes, _lib.ptr(out), _lib.stream(),
            launches=lib.otf_usm_launch_count(b * c, h, w),
        )
        return out
