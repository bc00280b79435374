"""Independent numpy restatement of ATen's separable resampling rules (test infrastructure).

otf_oracle.resize_pt calls ATen itself; this file restates the *published algorithm* of the
third-party dependency so that the index/weight rules the CUDA kernel implements are pinned by
something other than the kernel's author reading the same code twice:
  aten/src/ATen/native/cpu/UpSampleKernel.cpp  `_compute_indices_min_size_weights_aa`
  (torch 2.11.0; reached from traiNNer/data/degradations.py:1018-1021 via F.interpolate).
Arithmetic is float32, one rounding per operation (the C++ promotes through double at its 0.5 /
1.0 literals; reproduced where it matters).  Pure-Python loops: use on small images only.
"""

from __future__ import annotations

import numpy as np

f32 = np.float32


def _cubic(x: np.float32, a: float) -> np.float32:
    x = abs(x)
    a = f32(a)
    if x < 1:
        return ((a + f32(2)) * x - (a + f32(3))) * x * x + f32(1)
    if x < 2:
        return ((a * x - f32(5) * a) * x + f32(8) * a) * x - f32(4) * a
    return f32(0)


def _linear(x: np.float32) -> np.float32:
    x = abs(x)
    return f32(1) - x if x < 1 else f32(0)


def aa_weights(in_n: int, out_n: int, mode: str):
    """Per output index: (first source index, normalised taps) for antialias=True."""
    scale = f32(in_n) / f32(out_n)
    interp = 2 if mode == "bilinear" else 4
    support = f32(interp * 0.5) * scale if scale >= 1 else f32(interp * 0.5)
    invscale = f32(1) / scale if scale >= 1 else f32(1)
    table = []
    for i in range(out_n):
        center = scale * (f32(i) + f32(0.5))
        lo = max(int(center - support + f32(0.5)), 0)
        n = min(int(center + support + f32(0.5)), in_n) - lo
        ws, tot = [], f32(0)
        for j in range(n):
            arg = f32((np.float64(f32(j + lo) - center) + 0.5) * np.float64(invscale))
            w = _linear(arg) if mode == "bilinear" else _cubic(arg, -0.5)
            ws.append(f32(w))
            tot = f32(tot + w)
        table.append((lo, [f32(w / tot) for w in ws]))
    return table


def area_windows(in_n: int, out_n: int):
    return [((o * in_n) // out_n, -((-(o + 1) * in_n) // out_n)) for o in range(out_n)]


def nearest_exact_index(in_n: int, out_n: int):
    scale = f32(in_n) / f32(out_n)
    return [min(int(np.floor((f32(o) + f32(0.5)) * scale)), in_n - 1) for o in range(out_n)]


def _apply(img: np.ndarray, table, axis: int) -> np.ndarray:
    img = np.moveaxis(img, axis, -1)
    out = np.zeros(img.shape[:-1] + (len(table),), f32)
    for o, (lo, ws) in enumerate(table):
        acc = img[..., lo] * ws[0]
        for j in range(1, len(ws)):
            acc = (acc + img[..., lo + j] * ws[j]).astype(f32)
        out[..., o] = acc
    return np.moveaxis(out, -1, axis)


def resize_aa(img: np.ndarray, oh: int, ow: int, mode: str) -> np.ndarray:
    """(...,H,W) float32 -> (...,oh,ow): horizontal pass first, then vertical, as ATen does."""
    tmp = _apply(img.astype(f32), aa_weights(img.shape[-1], ow, mode), -1)
    return _apply(tmp, aa_weights(img.shape[-2], oh, mode), -2)
