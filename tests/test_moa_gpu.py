"""GPU parity of the MoA batch augment row through the C ABI: against the reference's
golden vectors (tests/golden/moa_goldens.npz) and, on larger seeded batches, against the CPU oracle.
Copies and mixup must be bit-exact; anything that goes through a resize is held to the fp32 stage
bound of the path (<= 1e-5 max-abs)."""

from __future__ import annotations

import os
import random

import numpy as np
import pytest
import torch

from oracle import batchaug_oracle as BO
from oracle import otf_oracle as O
from trainner_redux_b200 import batchaug as BA

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "moa_goldens.npz")
TOL = 1e-5
EXACT = {"mixup", "cutmix", "none"}


class _Rng:
    def __init__(self, seed: int) -> None:
        self.py, self.np, self.torch = random.Random(seed), np.random.default_rng(seed), torch.Generator().manual_seed(seed)


@pytest.fixture(scope="module")
def moa():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


def _compare(aug: str, got: torch.Tensor, want: torch.Tensor, what: str) -> None:
    got = got.cpu()
    assert got.shape == want.shape, what
    if aug in EXACT:
        assert torch.equal(got, want), f"{what}: max-abs {(got - want).abs().max().item()}"
    else:
        assert (got - want).abs().max().item() <= TOL, f"{what}: max-abs {(got - want).abs().max().item()}"


def test_goldens(moa, dev):
    for key in map(str, moa["cases"]):
        aug, tag, seed = key.split("_")
        gt, lq = torch.from_numpy(moa[f"in_{tag}_gt"]), torch.from_numpy(moa[f"in_{tag}_lq"])
        scale = int(moa[f"in_{tag}_scale"])
        want_gt = torch.from_numpy(moa[f"{key}_gt"]) if f"{key}_gt" in moa else gt
        g, l = BA.batch_aug(gt.to(dev), lq.to(dev), scale, [aug, "none"], [1.0, 0.0], rng=_Rng(int(seed)))
        _compare(aug, g, want_gt, key + " gt")
        _compare(aug, l, torch.from_numpy(moa[f"{key}_lq"]), key + " lq")


@pytest.mark.parametrize("aug", ["mixup", "cutmix", "resizemix", "cutblur", "downup", "up"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_against_oracle_at_training_sizes(dev, aug, seed):
    b, scale = (8, 4) if seed % 2 == 0 else (5, 2)
    gt_size = 256 if seed < 2 else 132  # 132/scale is not a multiple of 4: the scalar copy path
    gt = O.synth_gt(b, gt_size, gt_size, "natural", seed=seed)
    lq = torch.nn.functional.interpolate(gt, size=(gt_size // scale,) * 2, mode="area").clamp(0, 1)
    ora = _Rng(100 + seed)
    want_gt, want_lq, plan = BO.batch_aug(gt.clone(), lq.clone(), scale, [aug], [1.0], ora.py, ora.np, ora.torch)
    g, l = BA.batch_aug(gt.to(dev), lq.to(dev), scale, [aug], [1.0], rng=_Rng(100 + seed))
    _compare(aug, g, want_gt.contiguous(), f"{aug} seed {seed} gt {plan}")
    _compare(aug, l, want_lq.contiguous(), f"{aug} seed {seed} lq {plan}")


def test_in_place_semantics_and_untouched_outside_box(dev):
    """cutmix / resizemix / cutblur write into the tensors they are given (as the reference does) and leave
    everything outside the box bit-identical; mixup / downup / up return new tensors."""
    gt = torch.rand(6, 3, 128, 128, device=dev)
    lq = torch.rand(6, 3, 32, 32, device=dev)
    for fn in (BA.cutmix, BA.resizemix, BA.cutblur):
        g0, l0 = gt.clone(), lq.clone()
        g, l = fn(g0, l0, 4, rng=_Rng(7))
        assert g.data_ptr() == g0.data_ptr() and l.data_ptr() == l0.data_ptr()
        changed = (l != lq).any(dim=0).any(dim=0)
        ys, xs = torch.nonzero(changed, as_tuple=True)
        assert len(ys) > 0
        box = torch.zeros_like(changed)
        box[ys.min() : ys.max() + 1, xs.min() : xs.max() + 1] = True
        assert torch.equal(l[:, :, ~box], lq[:, :, ~box])
    g, l = BA.mixup(gt, lq, 4, rng=_Rng(7))
    assert g.data_ptr() != gt.data_ptr() and l.data_ptr() != lq.data_ptr()


def test_mixup_linearity_property(dev):
    """Size-independent property at the full training batch: mixup of a constant batch is that constant, and
    lam*x + (1-lam)*x[perm] summed over the batch keeps the batch mean (within fp32 rounding)."""
    gt = torch.rand(64, 3, 224, 224, device=dev)
    lq = torch.rand(64, 3, 56, 56, device=dev)
    g, l = BA.mixup(gt, lq, 4, rng=_Rng(3))
    assert abs(g.double().mean().item() - gt.double().mean().item()) < 1e-6
    assert abs(l.double().mean().item() - lq.double().mean().item()) < 1e-6
    c = torch.full((64, 3, 56, 56), 0.3, device=dev)
    g, _ = BA.mixup(c, c.clone(), 1, rng=_Rng(4))
    assert (g - 0.3).abs().max().item() < 1e-7


def test_abi_argument_checks(dev):
    import ctypes as C

    from trainner_redux_b200 import _lib

    x = torch.rand(4, 3, 16, 16, device=dev)
    perm = np.array([0, 1, 2, 9], dtype=np.int32)
    with pytest.raises(_lib.OtfError, match="perm"):
        _lib.call("otf_mixup_f32", _lib.ptr(x), perm.ctypes.data_as(C.c_void_p), 4, x[0].numel(), 0.5, 0.5,
                  _lib.ptr(torch.empty_like(x)), _lib.stream())
    good = np.arange(4, dtype=np.int32)
    with pytest.raises(_lib.OtfError, match="in place"):
        _lib.call("otf_mixup_f32", _lib.ptr(x), good.ctypes.data_as(C.c_void_p), 4, x[0].numel(), 0.5, 0.5, _lib.ptr(x), _lib.stream())
    with pytest.raises(_lib.OtfError, match="outside"):
        _lib.call("otf_copy_box_f32", _lib.ptr(x), 16, 16, 8, 8, _lib.ptr(torch.empty_like(x)), 16, 16, 0, 0, 12, 4, 4, 3, None, _lib.stream())
    with pytest.raises(_lib.OtfError, match="in place"):
        _lib.call("otf_copy_box_f32", _lib.ptr(x), 16, 16, 0, 0, _lib.ptr(x), 16, 16, 0, 0, 4, 4, 4, 3, good.ctypes.data_as(C.c_void_p), _lib.stream())


def test_feed_applies_moa_after_the_pool(dev):
    """RealESRGANFeed with use_moa: the pair that leaves feed_data is batch_aug(pair), drawn from the feed's own
    generators after the degradation plan (realesrgan_model.py:649-650)."""
    from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed

    def make(use_moa: bool) -> RealESRGANFeed:
        opt = OTFOptions(scale=4, gt_size=64, use_moa=use_moa, moa_augs=("mixup", "cutmix"), moa_probs=(0.5, 0.5))
        return RealESRGANFeed(opt, device=dev, manual_seed=5, use_pool=False)

    gt = O.synth_gt(4, 96, 96, "natural", seed=1)
    k = O.synth_blur_kernels(4, seed=2)
    data = {"gt": gt, "kernel1": k, "kernel2": k.clone(), "sinc_kernel": O.synth_sinc_or_pulse(4, seed=3)}
    plain, aug = make(False), make(True)
    plain.feed_data(dict(data))
    aug.feed_data(dict(data))
    # same plan, same Philox streams -> same pair before MoA; replay the augmentation on the plain pair
    rng = _Rng(0)
    rng.py, rng.np, rng.torch = plain.rng.py, plain.rng.np, plain.rng.torch
    want_gt, want_lq = BA.batch_aug(plain.gt.clone(), plain.lq.clone(), 4, ["mixup", "cutmix"], [0.5, 0.5], rng=rng)
    assert torch.equal(aug.gt, want_gt) and torch.equal(aug.lq, want_lq)
    assert not torch.equal(aug.gt, plain.gt)
