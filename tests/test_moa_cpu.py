"""CPU tests of the MoA batch augment row: the oracle against the reference's golden
vectors (tests/golden/moa_goldens.npz, written by oracle/make_moa_goldens.py from the reference's own
batchaug.py), and the product's HOST logic — which random numbers it consumes and which boxes /
permutations it hands to the C ABI — with the device calls replaced by recorders."""

from __future__ import annotations

import os
import random

import numpy as np
import pytest
import torch

from oracle import batchaug_oracle as BO
from trainner_redux_b200 import batchaug as BA

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "moa_goldens.npz")


@pytest.fixture(scope="module")
def moa():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


def _case(moa, key):
    aug, tag, seed = key.split("_")
    gt, lq, scale = torch.from_numpy(moa[f"in_{tag}_gt"]), torch.from_numpy(moa[f"in_{tag}_lq"]), int(moa[f"in_{tag}_scale"])
    want_gt = torch.from_numpy(moa[f"{key}_gt"]) if f"{key}_gt" in moa else gt
    return aug, int(seed), gt, lq, scale, want_gt, torch.from_numpy(moa[f"{key}_lq"])


def test_golden_file_covers_every_augmentation(moa):
    augs = {str(k).split("_")[0] for k in moa["cases"]}
    assert augs == {"mixup", "cutmix", "resizemix", "cutblur", "downup", "up"}
    assert len(moa["cases"]) >= 20


def test_oracle_reproduces_reference_goldens(moa):
    torch.set_num_threads(1)
    for key in map(str, moa["cases"]):
        aug, seed, gt, lq, scale, want_gt, want_lq = _case(moa, key)
        py, nprng, tgen = random.Random(seed), np.random.default_rng(seed), torch.Generator().manual_seed(seed)
        o_gt, o_lq, plan = BO.batch_aug(gt.clone(), lq.clone(), scale, [aug, "none"], [1.0, 0.0], py, nprng, tgen)
        assert torch.equal(o_gt.contiguous(), want_gt), key
        assert torch.equal(o_lq.contiguous(), want_lq), key
        if "box" in plan:
            assert list(plan["box"]) == moa[f"{key}_box"].tolist(), key
        if "perm" in plan:
            assert plan["perm"].tolist() == moa[f"{key}_perm"].tolist(), key


def test_oracle_error_behaviour():
    gt, lq = torch.rand(3, 3, 16, 16), torch.rand(3, 3, 8, 8)
    py, nprng = random.Random(0), np.random.default_rng(0)
    with pytest.raises(ValueError, match="batch >1"):
        BO.batch_aug(gt[:1], lq[:1], 2, ["none"], [1.0], py, nprng)
    with pytest.raises(ValueError, match="don't match"):
        BO.batch_aug(gt, lq, 2, ["none", "mixup"], [1.0], py, nprng)
    with pytest.raises(ValueError, match="is not invalid"):
        BO.batch_aug(gt, lq, 2, ["bogus"], [1.0], py, nprng)
    with pytest.raises(ValueError, match="same resolution"):
        BO.apply(gt, lq, {"aug": "cutmix", "scale": 3, "perm": torch.arange(3), "box": (0, 0, 3, 3)})


class _Rng:
    def __init__(self, seed: int) -> None:
        self.py, self.np, self.torch = random.Random(seed), np.random.default_rng(seed), torch.Generator().manual_seed(seed)


def test_product_host_draws_match_reference(moa, monkeypatch):
    """With the device calls stubbed out, the product must consume exactly the reference's random numbers
    and pass the golden boxes / permutations / ratios / samplers to the C ABI."""
    calls = []

    def fake_call(name, *args, **kw):
        calls.append((name, args))

    def fake_resize(x, oh, ow, mode_id, clamp):
        calls.append(("resize", (tuple(x.shape[2:]), (oh, ow), mode_id, clamp)))
        return torch.empty((x.shape[0], x.shape[1], oh, ow))

    monkeypatch.setattr(BA._lib, "call", fake_call)
    monkeypatch.setattr(BA._lib, "ptr", lambda t: None)
    monkeypatch.setattr(BA._lib, "stream", lambda: None)
    monkeypatch.setattr(BA, "_dense", lambda *ts: list(ts))
    monkeypatch.setattr(BA.D, "_resize_call", fake_resize)
    mode_of = {0: BA._lib.RESIZE_BICUBIC_AA, 1: BA._lib.RESIZE_BILINEAR_AA, 2: BA._lib.RESIZE_NEAREST_EXACT}
    for key in map(str, moa["cases"]):
        aug, seed, gt, lq, scale, _, _ = _case(moa, key)
        calls.clear()
        rng = _Rng(seed)
        BA.batch_aug(gt.clone(), lq.clone(), scale, [aug, "none"], [1.0, 0.0], rng=rng)
        # same generator state afterwards as the oracle's (which make_moa_goldens.py tied to the reference's)
        ora = _Rng(seed)
        BO.batch_aug(gt.clone(), lq.clone(), scale, [aug, "none"], [1.0, 0.0], ora.py, ora.np, ora.torch)
        assert rng.py.random() == ora.py.random(), key
        assert rng.np.random() == ora.np.random(), key
        assert torch.equal(torch.rand(2, generator=rng.torch), torch.rand(2, generator=ora.torch)), key
        box = moa[f"{key}_box"].tolist()
        perm = moa[f"{key}_perm"].tolist()
        lam = float(moa[f"{key}_lam"])
        samplers = moa[f"{key}_samplers"].tolist()
        names = [c[0] for c in calls]
        if aug == "mixup":
            assert names == ["otf_mixup_f32"] * 2
            for (_, a), x in zip(calls, (gt, lq)):
                assert a[2] == x.shape[0] and a[3] == x[0].numel() and a[4] == lam and a[5] == 1 - lam
        elif aug == "cutmix":
            x1, y1, x2, y2 = box
            # per tensor: stage the box (no permutation), paste it back permuted, at the same offsets
            assert names == ["otf_copy_box_f32"] * 4
            for i, bx in enumerate(((x1, y1, x2, y2), tuple(v // scale for v in box))):
                a1, b1, a2, b2 = bx
                stage, paste = calls[2 * i][1], calls[2 * i + 1][1]
                assert stage[3:5] == (a1, b1) and stage[10:12] == (a2 - a1, b2 - b1) and stage[14] is None
                assert paste[8:10] == (a1, b1) and paste[10:12] == (a2 - a1, b2 - b1) and paste[14] is not None
        elif aug == "resizemix":
            x1, y1, x2, y2 = box
            assert names == ["resize", "otf_copy_box_f32"] * 2
            assert calls[0][1] == (tuple(gt.shape[2:]), (y2 - y1, x2 - x1), BA._lib.RESIZE_BICUBIC_AA, True)
            assert calls[1][1][8:12] == (y1, x1, y2 - y1, x2 - x1)
            assert calls[2][1][1] == ((y2 - y1) // scale, (x2 - x1) // scale)
        elif aug == "cutblur":
            x1, y1, x2, y2 = box
            assert names == ["otf_copy_box_f32", "resize", "otf_copy_box_f32"]
            assert calls[0][1][3:5] == (x1, y1) and calls[0][1][10:12] == (x2 - x1, y2 - y1)
            assert calls[1][1] == ((x2 - x1, y2 - y1), ((x2 - x1) // scale, (y2 - y1) // scale), BA._lib.RESIZE_BICUBIC_AA, False)
            assert calls[2][1][8:10] == (x1 // scale, y1 // scale)
        elif aug == "downup":
            assert names == ["resize", "resize"]
            small = tuple(int(v) for v in np.round(np.array(lq.shape[2:]) * lam).astype(int))
            assert calls[0][1] == (tuple(lq.shape[2:]), small, mode_of[samplers[0]], False)
            assert calls[1][1] == (small, tuple(lq.shape[2:]), mode_of[samplers[1]], False)
        elif aug == "up":
            x1, y1, x2, y2 = box
            assert names == ["otf_copy_box_f32", "otf_copy_box_f32", "resize", "resize"]
            assert calls[0][1][3:5] == (x1, y1) and calls[0][1][10:12] == (x2 - x1, y2 - y1)
            assert calls[2][1] == ((x2 - x1, y2 - y1), tuple(gt.shape[2:]), BA._lib.RESIZE_BICUBIC_AA, False)
            assert calls[3][1][1:3] == (tuple(lq.shape[2:]), mode_of[samplers[2]])
        del perm  # the permutation's content is checked on the GPU: it travels to the ABI as a host pointer


def test_product_errors_without_device():
    gt, lq = torch.rand(3, 3, 16, 16), torch.rand(3, 3, 8, 8)
    rng = _Rng(0)
    with pytest.raises(ValueError, match="batch >1"):
        BA.batch_aug(gt[:1], lq[:1], 2, ["none"], [1.0], rng=rng)
    with pytest.raises(ValueError, match="don't match"):
        BA.batch_aug(gt, lq, 2, ["none", "mixup"], [1.0], rng=rng)
    with pytest.raises(ValueError, match="is not invalid"):
        BA.batch_aug(gt, lq, 2, ["bogus"], [1.0], rng=rng)
    with pytest.raises(ValueError, match="same resolution"):
        BA.cutmix(gt, lq, 3, rng=rng)
    with pytest.raises(RuntimeError):  # CPU tensors: there is no CPU path
        BA.mixup(gt, lq, 2, rng=rng)
    assert BA.batch_aug(gt, lq, 2, ["none"], [1.0], rng=rng) == (gt, lq)
