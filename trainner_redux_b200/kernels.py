"""On-GPU blur / sinc kernel synthesis.

The reference builds ``kernel1``, ``kernel2`` and ``sinc_kernel`` with numpy/scipy inside its
dataset workers (traiNNer/data/realesrgan_dataset.py:149-211 calling
traiNNer/data/degradations.py:22-507) and ships three (B,21,21) tensors to the GPU.  Here the
host only draws the *parameters* — in the reference's order, from the same two generators (Python
``random`` for sizes / kernel types, the numpy ``Generator`` for everything else) — and one small
kernel launch evaluates, normalises, pads and rounds all B kernels on the device.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib

TYPE_ID = {"iso": 0, "aniso": 1, "generalized_iso": 2, "generalized_aniso": 3, "plateau_iso": 4, "plateau_aniso": 5,
           "sinc": 6, "pulse": 7}
_DEFAULT_LIST = ("iso", "aniso", "generalized_iso", "generalized_aniso", "plateau_iso", "plateau_aniso")
_DEFAULT_PROB = (0.45, 0.25, 0.12, 0.03, 0.12, 0.03)


@dataclass
class KernelOptions:
    """The blur-kernel fields of ``DatasetOptions`` with the reference's names and defaults
    (traiNNer/utils/redux_options.py:101-142)."""

    kernel_list: Sequence[str] = _DEFAULT_LIST
    kernel_prob: Sequence[float] = _DEFAULT_PROB
    kernel_range: tuple[int, int] = (5, 17)
    sinc_prob: float = 0
    blur_sigma: tuple[float, float] = (0.2, 2)
    betag_range: tuple[float, float] = (0.5, 4)
    betap_range: tuple[float, float] = (1, 2)
    kernel_list2: Sequence[str] = _DEFAULT_LIST
    kernel_prob2: Sequence[float] = _DEFAULT_PROB
    kernel_range2: tuple[int, int] = (5, 17)
    sinc_prob2: float = 0
    blur_sigma2: tuple[float, float] = (0.2, 1)
    betag_range2: tuple[float, float] = (0.5, 4)
    betap_range2: tuple[float, float] = (1, 2)
    final_sinc_prob: float = 0
    final_kernel_range: tuple[int, int] = (5, 17)


def _row(kind: str, k: int, sx=0.0, sy=0.0, theta=0.0, beta=0.0, wc=0.0, pad=21) -> list[float]:
    return [float(TYPE_ID[kind]), float(k), sx, sy, theta, beta, wc, float(pad)]


def _draw_blur(py, nprng, sizes, sinc_prob, klist, kprob, sigma, betag, betap) -> list[float]:
    """One blur kernel: realesrgan_dataset.py:150-169 (and :175-194 for the second one)."""
    k = py.choice(sizes)
    if nprng.uniform() < sinc_prob:
        wc = nprng.uniform(np.pi / 3, np.pi) if k < 13 else nprng.uniform(np.pi / 5, np.pi)
        return _row("sinc", k, wc=wc)
    kind = py.choices(list(klist), list(kprob))[0]  # degradations.py:405
    sx = nprng.uniform(sigma[0], sigma[1])  # :240 / :292 / :344
    if kind.endswith("aniso"):
        sy = nprng.uniform(sigma[0], sigma[1])
        theta = nprng.uniform(-math.pi, math.pi)
    else:
        sy, theta = sx, 0.0
    beta = 0.0
    if kind.startswith(("generalized", "plateau")):
        lo, hi = betag if kind.startswith("generalized") else betap
        beta = nprng.uniform(lo, 1) if nprng.uniform() < 0.5 else nprng.uniform(1, hi)  # :302-305 / :355-358
    return _row(kind, k, sx, sy, theta, beta)


def draw_kernel_params(opt: Any, batch: int, py, nprng) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Parameter tables (batch, 8) float64 for kernel1, kernel2 and the final sinc kernel, drawn sample by
    sample in the order of ``RealESRGANDataset.__getitem__`` (realesrgan_dataset.py:149-206).  ``py`` is a
    ``random.Random``-like object, ``nprng`` a numpy ``Generator``."""
    r1 = list(range(opt.kernel_range[0], opt.kernel_range[1] + 1, 2))
    r2 = list(range(opt.kernel_range2[0], opt.kernel_range2[1] + 1, 2))
    r3 = list(range(opt.final_kernel_range[0], opt.final_kernel_range[1] + 1, 2))
    p1, p2, p3 = [], [], []
    for _ in range(batch):
        p1.append(_draw_blur(py, nprng, r1, opt.sinc_prob, opt.kernel_list, opt.kernel_prob, opt.blur_sigma, opt.betag_range, opt.betap_range))
        p2.append(_draw_blur(py, nprng, r2, opt.sinc_prob2, opt.kernel_list2, opt.kernel_prob2, opt.blur_sigma2, opt.betag_range2, opt.betap_range2))
        if nprng.uniform() < opt.final_sinc_prob:  # :200-206
            k = py.choice(r3)
            p3.append(_row("sinc", k, wc=nprng.uniform(np.pi / 3, np.pi)))
        else:
            p3.append(_row("pulse", 21))
    return np.asarray(p1, np.float64), np.asarray(p2, np.float64), np.asarray(p3, np.float64)


def synthesize_kernels(params: np.ndarray | Tensor, device: torch.device | str = "cuda", out: Tensor | None = None) -> Tensor:
    """(B,8) parameter table -> (B,21,21) fp32 kernels on the device (one launch).  ``out``: write into this buffer
    (feed_data keeps one per upload slot so that the kernels' addresses repeat and its captured chains are replayed)."""
    t = torch.as_tensor(params, dtype=torch.float64)
    if t.dim() != 2 or t.size(1) != 8:
        raise ValueError("params must have shape (B, 8)")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("kernel synthesis runs on a CUDA device only (no CPU fallback)")
    t = t.contiguous().to(dev, non_blocking=True)
    if out is None or tuple(out.shape) != (t.size(0), 21, 21) or out.dtype != torch.float32 or not out.is_contiguous():
        out = torch.empty((t.size(0), 21, 21), dtype=torch.float32, device=dev)
    _lib.call("otf_synth_kernels_f32", _lib.ptr(t), t.size(0), _lib.ptr(out), _lib.stream())
    return out
