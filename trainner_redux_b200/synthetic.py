"""Synthetic inputs for bench.py / profiles (SURVEY.md §8d "Synthetic inputs"): seeded GT batches and kernel
parameter tables.  Product-side data generation only — nothing here touches oracle/."""

from __future__ import annotations

import random

import numpy as np
import torch
from torch import Tensor
from torch.nn import functional as F  # noqa: N812

from .kernels import KernelOptions, draw_kernel_params


def synth_gt(b: int, h: int, w: int, kind: str = "uniform", seed: int = 1234) -> Tensor:
    """(b,3,h,w) fp32 in [0,1] on the CPU: "uniform" = i.i.d. U[0,1) (worst case for JPEG cliffs and Poisson vals);
    "natural" = 5x5 box-blurred noise plus a linear ramp."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 3, h, w, generator=g)
    if kind == "uniform":
        return x
    if kind == "natural":
        x = F.avg_pool2d(F.pad(x, (2, 2, 2, 2), mode="reflect"), 5, stride=1)
        ramp = torch.linspace(0, 1, w).view(1, 1, 1, w) * 0.5 + torch.linspace(0, 1, h).view(1, 1, h, 1) * 0.3
        return (x * 0.6 + ramp * 0.5).clamp(0, 1)
    raise ValueError(kind)


def bench_kernel_options(kernel_list=None, sinc_prob: float = 0.1, final_sinc_prob: float = 0.8) -> KernelOptions:
    """Real-ESRGAN's kernel settings on the reference's default kernel_list / kernel_prob
    (traiNNer/utils/redux_options.py:102-119): odd sizes 7..21, sigma [0.2,3] / [0.2,1.5], betag [0.5,4], betap [1,2]."""
    kw = {}
    if kernel_list is not None:
        kw = {"kernel_list": list(kernel_list), "kernel_prob": [1.0 / len(kernel_list)] * len(kernel_list),
              "kernel_list2": list(kernel_list), "kernel_prob2": [1.0 / len(kernel_list)] * len(kernel_list)}
    return KernelOptions(kernel_range=(7, 21), kernel_range2=(7, 21), final_kernel_range=(7, 21), sinc_prob=sinc_prob,
                         sinc_prob2=sinc_prob, final_sinc_prob=final_sinc_prob, blur_sigma=(0.2, 3), blur_sigma2=(0.2, 1.5),
                         betag_range=(0.5, 4), betag_range2=(0.5, 4), betap_range=(1, 2), betap_range2=(1, 2), **kw)


def synth_kernel_params(batch: int, seed: int, opt: KernelOptions | None = None):
    """Three (batch, 8) float64 tables (kernel1, kernel2, final sinc), drawn in the dataset's order."""
    return draw_kernel_params(opt or bench_kernel_options(), batch, random.Random(100 + seed), np.random.default_rng(200 + seed))
