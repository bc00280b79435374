"""Drop-in for the tensor branch of traiNNer/data/transforms.py:69-144 (``paired_random_crop``)."""

from __future__ import annotations

import random

import torch
from torch import Tensor

from . import _lib


def paired_random_crop(
    img_gt: Tensor, img_lq: Tensor, gt_patch_size: int, scale: int, gt_path: str | None = None
) -> tuple[Tensor, Tensor]:
    """One random (top,left) for the whole batch (transforms.py:119-120, Python ``random``);
    same ValueErrors as transforms.py:106-116.  Returns dense copies of both windows (the
    reference returns views and calls ``.contiguous()`` later, realesrgan_model.py:627)."""
    if not isinstance(img_gt, Tensor) or not isinstance(img_lq, Tensor):
        raise TypeError("tensor inputs only: the numpy branch of the reference is outside the hot path")
    _lib.require_cuda(img_gt, img_lq)
    h_lq, w_lq = img_lq.size()[-2:]
    h_gt, w_gt = img_gt.size()[-2:]
    lq_patch_size = gt_patch_size // scale
    if h_gt != h_lq * scale or w_gt != w_lq * scale:
        raise ValueError(
            f"Scale mismatches. GT ({h_gt}, {w_gt}) is not {scale}x ",
            f"multiplication of LQ ({h_lq}, {w_lq}). {gt_path}",
        )
    if h_lq < lq_patch_size or w_lq < lq_patch_size:
        raise ValueError(
            f"LQ ({h_lq}, {w_lq}) is smaller than patch size "
            f"({lq_patch_size}, {lq_patch_size}). "
            f"Please remove {gt_path}."
        )
    top = random.randint(0, h_lq - lq_patch_size)
    left = random.randint(0, w_lq - lq_patch_size)
    return crop_pair(img_gt, img_lq, gt_patch_size, scale, top, left)


def crop_pair(img_gt: Tensor, img_lq: Tensor, gt_patch_size: int, scale: int, top: int, left: int) -> tuple[Tensor, Tensor]:
    gt = _lib.dense_f32(img_gt)
    lq = _lib.dense_f32(img_lq)
    b, c, hg, wg = gt.shape
    hl, wl = lq.shape[-2:]
    p = gt_patch_size // scale
    g = p * scale
    if g != gt_patch_size:
        raise _lib.OtfError(f"gt_patch_size {gt_patch_size} must be a multiple of scale {scale}")
    gt_out = torch.empty((b, c, g, g), dtype=torch.float32, device=gt.device)
    lq_out = torch.empty((b, c, p, p), dtype=torch.float32, device=gt.device)
    _lib.call(
        "otf_crop_pair_f32", _lib.ptr(gt), b * c, hg, wg, _lib.ptr(lq), hl, wl, top, left, None, p, scale, 0,
        _lib.ptr(gt_out), _lib.ptr(lq_out), _lib.stream(),
    )
    return gt_out, lq_out
