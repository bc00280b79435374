"""CPU tests of the fork-extras row: the oracle restatement against the vectors frozen from the
reference's own paragon_otf_degradations.py (tests/golden/paragon_goldens.npz, made by oracle/make_paragon_goldens.py),
and the product's host-side draw order against the oracle's restatement of realesrgan_model.py:512-611."""

from __future__ import annotations

import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import paragon_oracle as P
from trainner_redux_b200.realesrgan_feed import HostRNG, OTFOptions, draw_plan

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "paragon_goldens.npz")


@pytest.fixture(scope="module")
def pg():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("key", ["nat", "uni", "sq"])
def test_oracle_reproduces_reference_vectors(pg, key):
    torch.set_num_threads(1)
    img = _t(pg[f"img_{key}"])
    for s in range(3):
        assert torch.equal(P.lens_distortion(img, float(pg[f"lens_{key}_{s}_p"])), _t(pg[f"lens_{key}_{s}"]))
        assert torch.equal(P.rolling_shutter(img, float(pg[f"shutter_{key}_{s}_p"])), _t(pg[f"shutter_{key}_{s}"]))
        ks, ang = pg[f"motion_{key}_{s}_p"]
        assert torch.equal(P.motion_blur(img, int(ks), float(ang)), _t(pg[f"motion_{key}_{s}"]))
        assert torch.equal(P.exposure(img, float(pg[f"exposure_{key}_{s}_p"])), _t(pg[f"exposure_{key}_{s}"]))
        assert torch.equal(P.color_temperature(img, float(pg[f"ctemp_{key}_{s}_p"])), _t(pg[f"ctemp_{key}_{s}"]))
        assert torch.equal(P.oversharpen(img, float(pg[f"oversharp_{key}_{s}_p"])), _t(pg[f"oversharp_{key}_{s}"]))
        assert torch.equal(P.aliasing(img, float(pg[f"alias_{key}_{s}_p"])), _t(pg[f"alias_{key}_{s}"]))
        got = P.sensor_noise(img, float(pg[f"sensor_{key}_{s}_p"]), _t(pg[f"sensor_{key}_{s}_noise"]))
        assert torch.equal(got, _t(pg[f"sensor_{key}_{s}"]))
    assert torch.equal(P.chromatic_aberration(img), _t(pg[f"chroma_{key}"]))
    assert torch.equal(P.demosaic(img), _t(pg[f"demosaic_{key}"]))  # numpy restatement of OpenCV's bilinear Bayer demosaic


def test_demosaic_restatement_small_and_degenerate_sizes(pg):
    for hh, ww in ((3, 3), (4, 7), (9, 4), (2, 6), (31, 33)):
        got = P.demosaic(_t(pg[f"demosaic_small_{hh}x{ww}_in"]))
        assert torch.equal(got, _t(pg[f"demosaic_small_{hh}x{ww}"])), (hh, ww)
    assert torch.all(_t(pg["demosaic_small_2x6"]) == 0)  # OpenCV leaves images under 3 rows / columns at zero
    try:
        import cv2  # noqa: F401
    except ImportError:
        return
    x = torch.rand(2, 3, 37, 52, generator=torch.Generator().manual_seed(4))
    assert torch.equal(P.demosaic(x), P.demosaic_cv2(x))


def test_ieee_sqrt_variant_is_within_one_coordinate_ulp(pg):
    """ATen's vectorised CPU sqrt is not correctly rounded; the IEEE variant (what CUDA computes) moves a source
    coordinate by at most one ulp: <= 1.5e-5 of a pixel times the local gradient."""
    for key in ("nat", "uni", "sq"):
        for s in range(3):
            d = np.abs(pg[f"lens_{key}_{s}"] - pg[f"lens_{key}_{s}_ieee"]).max()
            assert d <= 3e-5, (key, s, d)


def test_motion_kernel_is_a_normalised_line():
    for ks, ang in ((5, 0.0), (9, 45.0), (15, 120.0), (8, 10.0)):
        k = P.motion_blur_kernel(ks, ang)
        assert k.shape == (ks, ks) and abs(k.sum().item() - 1) < 1e-6
        nz = k[k > 0]
        assert torch.all(nz == nz[0]) and ks <= nz.numel() <= 2 * ks


def _opts(**kw):
    base = dict(order="fork", gt_size=32, blur_prob=0.5, lens_distort_prob=0.5, chromatic_aberration_prob=0.5, motion_blur_prob=0.5,
                sensor_noise_prob=0.5, rolling_shutter_prob=0.5, exposure_prob=0.5, color_temp_prob=0.5, oversharpen_prob=0.5,
                aliasing_prob=0.5, recompression_prob=0.5, editing_prob=0.5, editing_exposure_prob=0.5, editing_oversharpen_prob=0.5)
    base.update(kw)
    return OTFOptions(**base)


def test_product_draws_follow_the_reference_order():
    """draw_plan(order="fork") consumes the numpy / random streams exactly as the oracle's restatement of the
    reference's feed_data does (the restatement itself is pinned to the reference when the goldens are made)."""
    opt = _opts()
    seen = set()
    for seed in range(40):
        rng = HostRNG(0)
        rng.np, rng.py = np.random.default_rng(seed), random.Random(seed + 7)
        mine = draw_plan(opt, 2, 64, 48, rng)
        want_np = np.random.default_rng(seed)
        want_np.uniform()  # the p_clean gate (realesrgan_model.py:487-489): drawn first, even at probability 0
        want = P.draw_extras(opt, want_np, random.Random(seed + 7))
        for k, v in want.items():
            assert mine.get(k) == v, (seed, k)
        seen.update(want)
        # both consumed the same number of draws: the next values agree
        ref_np, ref_py = np.random.default_rng(seed), random.Random(seed + 7)
        ref_np.uniform()
        P.draw_extras(opt, ref_np, ref_py)
        ref_py.randint(0, 64 // 4 - 8), ref_py.randint(0, 48 // 4 - 8)  # the crop offsets draw_plan takes afterwards
        assert rng.np.uniform() == ref_np.uniform() and rng.py.random() == ref_py.random()
    assert {"lens", "chroma", "motion", "sensor", "shutter", "exposure", "color_temp", "oversharpen", "aliasing", "editing_exposure"} <= seen


def test_gates_draw_even_at_probability_zero():
    """Every stage whose option fields exist consumes its gate draw (the reference's hasattr guards): 15 uniforms,
    the p_clean gate first (ReduxOptions always defines it)."""
    opt = OTFOptions(order="fork", gt_size=32, compression_formats=("jpeg",), compression_weights=(1.0,))
    rng = HostRNG(3)
    plan = draw_plan(opt, 2, 64, 64, rng)
    assert not any(k in plan for k in ("lens", "chroma", "motion", "sensor", "shutter", "exposure", "color_temp", "oversharpen", "aliasing"))
    assert plan["compression"][0][0] == "jpeg" and 45 <= plan["compression"][0][1] <= 95
    ref = np.random.default_rng(3)
    for _ in range(12):  # p_clean, lens, chroma, motion, blur, demosaic, sensor, shutter, exposure, colour temp, oversharpen, aliasing
        ref.uniform()
    ref.choice(["jpeg"], p=[1.0]); ref.uniform(45, 95); ref.uniform(); ref.uniform()  # format, quality, recompression, editing
    assert rng.np.uniform() == ref.uniform()


def test_chain_plan_is_stored_with_the_goldens(pg):
    plan = json.loads(bytes(pg["chain_plan_json"]).decode())
    assert plan["scale"] == 4 and "resize3_mode" in plan and sum(k in plan for k in P.draw_extras.__code__.co_consts if isinstance(k, str)) >= 5
