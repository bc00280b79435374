"""MoA batch augment on the finished pair — drop-in for traiNNer/ops/batchaug.py.

Same names, arguments, defaults, host random draws (in the same order, from the same three
generators) and exceptions as the reference; the tensor work runs in libotf_b200:

  mixup      batchaug.py:129-158   otf_mixup_f32 (one launch per tensor, permutation in the parameter bank)
  cutmix     :162-229              otf_copy_box_f32 x2 per tensor (stage the boxes, paste them permuted)
  resizemix  :233-321              otf_resize_f32 of the un-permuted batch + permuted paste
  cutblur    :350-403              box crop + otf_resize_f32 (bicubic, antialias) + paste into LQ
  downup     :406-444              two otf_resize_f32
  up         :447-509              box crops + otf_resize_f32

There is no CPU path: tensors must be CUDA fp32.  The reference draws from three GLOBAL generators
(`random`, `RNG.get_rng()`, torch's default CPU generator); pass ``rng=`` (anything with ``.py``,
``.np``, ``.torch`` — e.g. realesrgan_feed.HostRNG) to use explicit ones, or call ``init_rng(seed)``
once to seed the module's numpy generator the way traiNNer/utils/rng.py:19-25 does.
The reference's debug image dump (`moa_debug`) is not reproduced.
"""

from __future__ import annotations

import ctypes as C
import math
import random as _random
from types import SimpleNamespace
from typing import Any

import numpy as np
import torch
from torch import Size, Tensor

from . import _lib
from . import degradations as D

_SAMPLERS = (("bicubic", True), ("bilinear", True), ("nearest-exact", False))
_MODE = {"bicubic": _lib.RESIZE_BICUBIC_AA, "bilinear": _lib.RESIZE_BILINEAR_AA, "nearest-exact": _lib.RESIZE_NEAREST_EXACT}
_np_rng: np.random.Generator | None = None


def init_rng(seed: int) -> None:
    """Seed the module-level numpy generator (the reference's RNG.init_rng)."""
    global _np_rng
    _np_rng = np.random.default_rng(seed)


def _rng(rng: Any | None) -> Any:
    if rng is not None:
        return rng
    if _np_rng is None:
        raise RuntimeError("Manual seed is not set.")  # traiNNer/utils/rng.py:22-24
    return SimpleNamespace(py=_random, np=_np_rng, torch=None)


def _randperm(n: int, rng: Any) -> Tensor:
    g = getattr(rng, "torch", None)
    return torch.randperm(n) if g is None else torch.randperm(n, generator=g)


def _perm_arg(perm: Tensor) -> tuple[np.ndarray, C.c_void_p]:
    arr = np.ascontiguousarray(perm.cpu().numpy().astype(np.int32))
    return arr, arr.ctypes.data_as(C.c_void_p)


def _check_pair(img_gt: Tensor, img_lq: Tensor, scale: int) -> None:
    if img_gt.size()[3] != img_lq.size()[3] * scale or img_gt.size()[2] != img_lq.size()[2] * scale:
        raise ValueError("img_gt and img_lq have to be the same resolution.")


def _dense(*ts: Tensor) -> list[Tensor]:
    _lib.require_cuda(*ts)
    return [_lib.dense_f32(t) for t in ts]


def _copy_box(src: Tensor, sy: int, sx: int, dst: Tensor, dy: int, dx: int, bh: int, bw: int, perm: Tensor | None = None) -> None:
    """dst[b, :, dy:dy+bh, dx:dx+bw] = src[perm[b] or b, :, sy:sy+bh, sx:sx+bw] (dense NCHW tensors)."""
    if bh <= 0 or bw <= 0:
        return
    b, c = dst.shape[:2]
    keep, pp = _perm_arg(perm) if perm is not None else (None, None)
    _lib.call("otf_copy_box_f32", _lib.ptr(src), src.shape[2], src.shape[3], sy, sx, _lib.ptr(dst), dst.shape[2], dst.shape[3],
              dy, dx, bh, bw, b, c, pp, _lib.stream())
    del keep


def _crop(img: Tensor, y0: int, y1: int, x0: int, x1: int) -> Tensor:
    """Dense copy of img[:, :, y0:y1, x0:x1] with Python's slice clipping."""
    h, w = img.shape[2:]
    y0, y1, x0, x1 = min(y0, h), min(y1, h), min(x0, w), min(x1, w)
    out = torch.empty((img.shape[0], img.shape[1], max(y1 - y0, 0), max(x1 - x0, 0)), dtype=torch.float32, device=img.device)
    _copy_box(img, y0, x0, out, 0, 0, out.shape[2], out.shape[3])
    return out


def _paste(dst: Tensor, y0: int, y1: int, x0: int, x1: int, src: Tensor, perm: Tensor | None = None) -> None:
    """dst[:, :, y0:y1, x0:x1] = src[perm]; the shapes must agree as a slice assignment requires."""
    h, w = dst.shape[2:]
    y0, y1, x0, x1 = min(y0, h), min(y1, h), min(x0, w), min(x1, w)
    bh, bw = max(y1 - y0, 0), max(x1 - x0, 0)
    if tuple(src.shape[2:]) != (bh, bw):
        raise RuntimeError(f"The expanded size of the tensor ({bh}, {bw}) must match the existing size {tuple(src.shape[2:])}")
    _copy_box(src, 0, 0, dst, y0, x0, bh, bw, perm)


def _interp(x: Tensor, size: tuple[int, int], sampler: tuple[str, bool], clamp: bool = False) -> Tensor:
    if size[0] <= 0 or size[1] <= 0:
        raise RuntimeError(f"Input and output sizes should be greater than 0, but got output (H: {size[0]}, W: {size[1]})")
    return D._resize_call(x, int(size[0]), int(size[1]), _MODE[sampler[0]], clamp)


def _centre_box(rng: Any, w: int, h: int, cut_w: int, cut_h: int, scale: int) -> tuple[int, int, int, int]:
    cx = int(rng.np.integers(w, dtype=int))
    cy = int(rng.np.integers(h, dtype=int))
    lo_x, hi_x = int(np.clip(cx - cut_w // 2, 0, w)), int(np.clip(cx + cut_w // 2, 0, w))
    lo_y, hi_y = int(np.clip(cy - cut_h // 2, 0, h)), int(np.clip(cy + cut_h // 2, 0, h))
    return lo_x * scale, lo_y * scale, hi_x * scale, hi_y * scale


class BatchAugment:
    """batchaug.py:21-44: reads ``moa_augs`` / ``moa_probs`` from the train options."""

    def __init__(self, scale: int, train_opt: Any, rng: Any | None = None) -> None:
        self.moa_augs = train_opt.moa_augs
        self.moa_probs = train_opt.moa_probs
        self.scale = scale
        self.debug = getattr(train_opt, "moa_debug", False)
        self.debug_limit = getattr(train_opt, "moa_debug_limit", 0)
        self.rng = rng

    def __call__(self, img1: Tensor, img2: Tensor) -> tuple[Tensor, Tensor]:
        return batch_aug(img1, img2, self.scale, self.moa_augs, self.moa_probs, self.debug, self.debug_limit, rng=self.rng)


def batch_aug(img_gt: Tensor, img_lq: Tensor, scale: int, augs: list[str], probs: list[float], debug: bool = False,
              debug_limit: int = 0, *, rng: Any | None = None) -> tuple[Tensor, Tensor]:
    """batchaug.py:47-128: pick ONE augmentation with ``random.choices`` and apply it to the batch."""
    if debug:
        raise NotImplementedError("moa_debug image dumps are not part of the B200 path")
    if len(augs) != len(probs):
        raise ValueError("Length of 'augmentation' and aug_prob don't match!")
    if img_gt.shape[0] == 1:
        raise ValueError("Augmentations need batch >1 to work.")
    rng = _rng(rng)
    aug = augs[rng.py.choices(range(len(augs)), weights=probs)[0]]
    if aug == "none":
        return img_gt, img_lq
    if aug == "cutmix":
        return cutmix(img_gt, img_lq, scale, rng=rng)
    if aug == "mixup":
        return mixup(img_gt, img_lq, scale, rng=rng)
    if aug == "resizemix":
        return resizemix(img_gt, img_lq, scale, rng=rng)
    if aug == "cutblur":
        return cutblur(img_gt, img_lq, scale, rng=rng)
    if aug == "downup":
        return downup(img_gt, img_lq, rng=rng)
    if aug == "up":
        return up(img_gt, img_lq, scale, rng=rng)
    raise ValueError(f"{aug} is not invalid.")


@torch.no_grad()
def mixup(img_gt: Tensor, img_lq: Tensor, scale: int, alpha_min: float = 0.4, alpha_max: float = 0.6, *,
          rng: Any | None = None) -> tuple[Tensor, Tensor]:
    """batchaug.py:129-158: lam * x + (1 - lam) * x[randperm], lam ~ U(alpha_min, alpha_max)."""
    rng = _rng(rng)
    lam = float(rng.np.uniform(alpha_min, alpha_max))
    perm = _randperm(img_gt.size(0), rng)
    img_gt, img_lq = _dense(img_gt, img_lq)
    keep, pp = _perm_arg(perm)
    outs = []
    for x in (img_gt, img_lq):
        out = torch.empty_like(x)
        _lib.call("otf_mixup_f32", _lib.ptr(x), pp, x.shape[0], x[0].numel(), lam, 1 - lam, _lib.ptr(out), _lib.stream())
        outs.append(out)
    del keep
    return outs[0], outs[1]


@torch.no_grad()
def cutmix(img_gt: Tensor, img_lq: Tensor, scale: int, alpha: float = 0.9, *, rng: Any | None = None) -> tuple[Tensor, Tensor]:
    """batchaug.py:162-229: paste a random box from the permuted batch (in place on dense inputs).
    As in the reference the box's x range comes from dim 2 and indexes dim 2."""
    _check_pair(img_gt, img_lq, scale)
    rng = _rng(rng)
    size: Size = img_gt.size()
    lam = float(rng.np.uniform(0, alpha))
    perm = _randperm(size[0], rng)
    w, h = size[2] // scale, size[3] // scale
    cut_rat = np.sqrt(1.0 - lam)
    x1, y1, x2, y2 = _centre_box(rng, w, h, int(w * cut_rat), int(h * cut_rat), scale)
    img_gt, img_lq = _dense(img_gt, img_lq)
    for img, (a1, b1, a2, b2) in ((img_gt, (x1, y1, x2, y2)), (img_lq, (x1 // scale, y1 // scale, x2 // scale, y2 // scale))):
        _paste(img, a1, a2, b1, b2, _crop(img, a1, a2, b1, b2), perm)
    return img_gt, img_lq


@torch.no_grad()
def resizemix(img_gt: Tensor, img_lq: Tensor, scale: int, scope: tuple[float, float] = (0.5, 0.9), *,
              rng: Any | None = None) -> tuple[Tensor, Tensor]:
    """batchaug.py:233-321: shrink the permuted batch (bicubic, antialias, clamp) into a random box.
    The resize is per sample, so the un-permuted batch is resized and the permutation applied by the paste."""
    _check_pair(img_gt, img_lq, scale)
    rng = _rng(rng)
    size: Size = img_gt.size()
    perm = _randperm(size[0], rng)
    tao = float(rng.np.uniform(scope[0], scope[1]))
    w, h = size[2] // scale, size[3] // scale
    x1, y1, x2, y2 = _centre_box(rng, w, h, int(w * tao), int(h * tao), scale)
    img_gt, img_lq = _dense(img_gt, img_lq)
    for img, (a1, b1, a2, b2) in ((img_gt, (x1, y1, x2, y2)), (img_lq, (x1 // scale, y1 // scale, x2 // scale, y2 // scale))):
        small = _interp(img, (b2 - b1, a2 - a1), _SAMPLERS[0], clamp=True)
        _paste(img, b1, b2, a1, a2, small, perm)
    return img_gt, img_lq


@torch.no_grad()
def cutblur(img_gt: Tensor, img_lq: Tensor, scale: int, alpha: float = 0.7, *, rng: Any | None = None) -> tuple[Tensor, Tensor]:
    """batchaug.py:350-403: replace a random LQ box with the bicubic-antialias downscale of the GT box."""
    _check_pair(img_gt, img_lq, scale)
    rng = _rng(rng)
    size: Size = img_gt.size()
    lam = float(rng.np.uniform(0.2, alpha))
    w, h = size[2] // scale, size[3] // scale
    x1, y1, x2, y2 = _centre_box(rng, w, h, int(w * lam), int(h * lam), scale)
    img_gt, img_lq = _dense(img_gt, img_lq)
    box = _crop(img_gt, x1, x2, y1, y2)
    # F.interpolate(scale_factor=1/scale): output extent = floor(extent * (1/scale)) in double
    oh, ow = math.floor(float(box.shape[2]) * (1 / scale)), math.floor(float(box.shape[3]) * (1 / scale))
    _paste(img_lq, x1 // scale, x2 // scale, y1 // scale, y2 // scale, _interp(box, (oh, ow), _SAMPLERS[0]))
    return img_gt, img_lq


def downup(img_gt: Tensor, img_lq: Tensor, scope: tuple[float, float] = (0.5, 0.9), *, rng: Any | None = None) -> tuple[Tensor, Tensor]:
    """batchaug.py:406-444: LQ down by U(scope) and back up with random samplers (never nearest twice)."""
    rng = _rng(rng)
    nearest = _SAMPLERS[2][0]
    down_sample, up_sample = rng.py.choice(_SAMPLERS), rng.py.choice(_SAMPLERS)
    if down_sample[0] == nearest and up_sample[0] == nearest:
        if rng.np.random() > 0.5:
            while up_sample[0] == nearest:
                up_sample = rng.py.choice(_SAMPLERS)
        else:
            while down_sample[0] == nearest:
                down_sample = rng.py.choice(_SAMPLERS)
    scale_factor = rng.np.uniform(scope[0], scope[1])
    (lq,) = _dense(img_lq)
    base = tuple(lq.shape[2:])
    small = tuple(int(v) for v in np.round(np.array(base) * scale_factor).astype(int))
    lq = _interp(_interp(lq, small, down_sample), base, up_sample)
    return img_gt, lq


def up(img_gt: Tensor, img_lq: Tensor, scale: int, scope: tuple[float, float] = (0.5, 0.9), *, rng: Any | None = None) -> tuple[Tensor, Tensor]:
    """batchaug.py:447-509: crop a random box and upscale it back to full size (GT bicubic, LQ random sampler)."""
    rng = _rng(rng)
    size: Size = img_gt.size()
    lam = float(rng.np.uniform(scope[0], scope[1]))
    w, h = size[2] // scale, size[3] // scale
    pad_w, pad_h = int(w * lam) // 2, int(h * lam) // 2
    cx = int(rng.np.integers(pad_w, w - pad_w, dtype=int))
    cy = int(rng.np.integers(pad_h, h - pad_w, dtype=int))  # (sic) batchaug.py:455
    x1, y1, x2, y2 = (cx - pad_w) * scale, (cy - pad_h) * scale, (cx + pad_w) * scale, (cy + pad_h) * scale
    gt, lq = _dense(img_gt, img_lq)
    gt_c = _crop(gt, x1, x2, y1, y2)
    lq_c = _crop(lq, x1 // scale, x2 // scale, y1 // scale, y2 // scale)
    assert gt_c.shape[2] == gt_c.shape[3], f"Expected crop to be square, got shape {tuple(gt_c.shape)}"
    lq_up_sample = rng.py.choice(_SAMPLERS)
    return _interp(gt_c, tuple(gt.shape[2:]), _SAMPLERS[0]), _interp(lq_c, tuple(lq.shape[2:]), lq_up_sample)
