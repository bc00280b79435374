"""TEST INFRASTRUCTURE — CPU oracle of the MoA batch augment.

Restates traiNNer/ops/batchaug.py:21-509 as two steps the reference interleaves: ``draw`` consumes
the host random numbers in the reference's order and returns a plan (augmentation, ratio,
permutation, box); ``apply`` performs the tensor work with the same ATen calls on the CPU, so the
result is bit-identical to the reference given the same generator states
(`oracle/make_moa_goldens.py` asserts that and writes `tests/golden/moa_goldens.npz`).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

The reference's generators are globals (`random`, `RNG.get_rng()`, torch's default generator);
here they are explicit: ``py`` (a ``random.Random`` or the module), ``nprng`` (numpy Generator),
``tgen`` (torch.Generator or None for the global one).

Quirks kept on purpose (they change results for non-square inputs):
  * cutmix/cutblur take the box "width" from dim 2 and the "height" from dim 3 and index dim 2 with
    the x range (batchaug.py:189-206, :222-227, :372-401);
  * resizemix derives the box the same way but indexes dim 2 with the y range (:318-319);
  * `up` draws cy from [pad_h, h - pad_w) (:455).
"""

from __future__ import annotations

import math
from typing import Any

import numpy as np
import torch
from torch import Tensor
from torch.nn import functional as F  # noqa: N812

SAMPLERS = (("bicubic", True), ("bilinear", True), ("nearest-exact", False))  # batchaug.py:407, :445


def _check_pair(gt: Tensor, lq: Tensor, scale: int) -> None:
    # batchaug.py:178-183 (same test in resizemix and cutblur)
    if gt.size(3) != lq.size(3) * scale or gt.size(2) != lq.size(2) * scale:
        raise ValueError("img_gt and img_lq have to be the same resolution.")


def _centre_box(nprng: np.random.Generator, w: int, h: int, cut_w: int, cut_h: int) -> tuple[int, int, int, int]:
    """Box of (cut_w, cut_h) around a uniform centre, clipped to [0,w]x[0,h] — batchaug.py:195-204."""
    cx = int(nprng.integers(w, dtype=int))
    cy = int(nprng.integers(h, dtype=int))
    x1, x2 = int(np.clip(cx - cut_w // 2, 0, w)), int(np.clip(cx + cut_w // 2, 0, w))
    y1, y2 = int(np.clip(cy - cut_h // 2, 0, h)), int(np.clip(cy + cut_h // 2, 0, h))
    return x1, y1, x2, y2


def _randperm(n: int, tgen: torch.Generator | None) -> Tensor:
    return torch.randperm(n) if tgen is None else torch.randperm(n, generator=tgen)


def draw(aug: str, gt_shape: tuple[int, ...], lq_shape: tuple[int, ...], scale: int, py: Any, nprng: np.random.Generator,
         tgen: torch.Generator | None = None) -> dict:
    """Consume the host draws of one augmentation in the reference's order."""
    b = gt_shape[0]
    w, h = gt_shape[2] // scale, gt_shape[3] // scale  # (sic) batchaug.py:189-190
    plan: dict[str, Any] = {"aug": aug, "scale": scale}
    if aug == "mixup":  # :150-153
        plan["lam"] = float(nprng.uniform(0.4, 0.6))
        plan["perm"] = _randperm(b, tgen)
    elif aug == "cutmix":  # :208-209, then the box :218
        lam = float(nprng.uniform(0, 0.9))
        plan["perm"] = _randperm(b, tgen)
        rat = np.sqrt(1.0 - lam)
        plan["box"] = tuple(v * scale for v in _centre_box(nprng, w, h, int(w * rat), int(h * rat)))
    elif aug == "resizemix":  # :277, :284, :290
        plan["perm"] = _randperm(b, tgen)
        tao = float(nprng.uniform(0.5, 0.9))
        plan["box"] = tuple(v * scale for v in _centre_box(nprng, w, h, int(w * tao), int(h * tao)))
    elif aug == "cutblur":  # :390-391
        lam = float(nprng.uniform(0.2, 0.7))
        plan["box"] = tuple(v * scale for v in _centre_box(nprng, w, h, int(w * lam), int(h * lam)))
    elif aug == "downup":  # :409-425
        down, up = py.choice(SAMPLERS), py.choice(SAMPLERS)
        if down[0] == "nearest-exact" and up[0] == "nearest-exact":
            if nprng.random() > 0.5:
                while up[0] == "nearest-exact":
                    up = py.choice(SAMPLERS)
            else:
                while down[0] == "nearest-exact":
                    down = py.choice(SAMPLERS)
        plan["down"], plan["up"] = down, up
        plan["factor"] = float(nprng.uniform(0.5, 0.9))
    elif aug == "up":  # :466-469, :485
        lam = float(nprng.uniform(0.5, 0.9))
        pad_w, pad_h = int(w * lam) // 2, int(h * lam) // 2
        cx = int(nprng.integers(pad_w, w - pad_w, dtype=int))
        cy = int(nprng.integers(pad_h, h - pad_w, dtype=int))  # (sic) :455
        plan["box"] = ((cx - pad_w) * scale, (cy - pad_h) * scale, (cx + pad_w) * scale, (cy + pad_h) * scale)
        plan["lq_up"] = py.choice(SAMPLERS)
    elif aug != "none":
        raise ValueError(f"{aug} is not invalid.")  # (sic) :109
    return plan


def apply(gt: Tensor, lq: Tensor, plan: dict) -> tuple[Tensor, Tensor]:
    """Tensor work of one augmentation (in place where the reference is in place)."""
    aug, scale = plan["aug"], plan["scale"]
    if aug == "none":
        return gt, lq
    if aug == "mixup":  # :154-158
        lam, idx = plan["lam"], plan["perm"]
        return lam * gt + (1 - lam) * gt[idx], lam * lq + (1 - lam) * lq[idx]
    if aug == "cutmix":  # :211-227
        _check_pair(gt, lq, scale)
        idx = plan["perm"]
        x1, y1, x2, y2 = plan["box"]
        gt_p, lq_p = gt[idx], lq[idx]
        gt[:, :, x1:x2, y1:y2] = gt_p[:, :, x1:x2, y1:y2]
        a1, b1, a2, b2 = (v // scale for v in plan["box"])
        lq[:, :, a1:a2, b1:b2] = lq_p[:, :, a1:a2, b1:b2]
        return gt, lq
    if aug == "resizemix":  # :277-319
        _check_pair(gt, lq, scale)
        idx = plan["perm"]
        x1, y1, x2, y2 = plan["box"]
        a1, b1, a2, b2 = (v // scale for v in plan["box"])
        gt_small = F.interpolate(gt.clone()[idx], (y2 - y1, x2 - x1), mode="bicubic", antialias=True).clamp(0, 1)
        lq_small = F.interpolate(lq.clone()[idx], (b2 - b1, a2 - a1), mode="bicubic", antialias=True).clamp(0, 1)
        gt[:, :, y1:y2, x1:x2] = gt_small
        lq[:, :, b1:b2, a1:a2] = lq_small
        return gt, lq
    if aug == "cutblur":  # :393-401
        _check_pair(gt, lq, scale)
        x1, y1, x2, y2 = plan["box"]
        lq[:, :, x1 // scale : x2 // scale, y1 // scale : y2 // scale] = F.interpolate(
            gt[:, :, x1:x2, y1:y2], scale_factor=1 / scale, mode="bicubic", antialias=True)
        return gt, lq
    if aug == "downup":  # :427-444
        base = lq.shape[2:]
        small = list(np.round(np.array(base) * plan["factor"]).astype(int))
        lq = F.interpolate(lq, size=small, mode=plan["down"][0], antialias=plan["down"][1])
        lq = F.interpolate(lq, size=base, mode=plan["up"][0], antialias=plan["up"][1])
        return gt, lq
    if aug == "up":  # :471-507
        gt_base, lq_base = gt.shape[2:], lq.shape[2:]
        x1, y1, x2, y2 = plan["box"]
        a1, b1, a2, b2 = (v // scale for v in plan["box"])
        gt_c, lq_c = gt[:, :, x1:x2, y1:y2], lq[:, :, a1:a2, b1:b2]
        assert gt_c.shape[2] == gt_c.shape[3], "Expected crop to be square"
        gt = F.interpolate(gt_c, size=gt_base, mode="bicubic", antialias=True)
        lq = F.interpolate(lq_c, size=lq_base, mode=plan["lq_up"][0], antialias=plan["lq_up"][1])
        return gt, lq
    raise ValueError(f"{aug} is not invalid.")


def batch_aug(gt: Tensor, lq: Tensor, scale: int, augs: list[str], probs: list[float], py: Any, nprng: np.random.Generator,
              tgen: torch.Generator | None = None) -> tuple[Tensor, Tensor, dict]:
    """batchaug.py:47-128 without the debug image dumps. Returns the plan too."""
    if len(augs) != len(probs):
        raise ValueError("Length of 'augmentation' and aug_prob don't match!")
    if gt.shape[0] == 1:
        raise ValueError("Augmentations need batch >1 to work.")
    aug = augs[py.choices(range(len(augs)), weights=probs)[0]]
    with torch.no_grad():
        plan = draw(aug, tuple(gt.shape), tuple(lq.shape), scale, py, nprng, tgen)
        gt, lq = apply(gt, lq, plan)
    return gt, lq, plan


def floor_size(n: int, factor: float) -> int:
    """Output extent F.interpolate derives from a scale_factor: floor(n * factor) in double."""
    return int(math.floor(float(n) * factor))
