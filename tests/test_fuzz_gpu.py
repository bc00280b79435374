"""Randomised shape/parameter sweep of every primitive against the CPU oracle (bounded: ~60 cases).
Targets the places where shape-dependent code paths switch: TMA vs cp.async fill (W % 4), partial tiles, tile
configurations (64x64 / 64x32 / 32x32), rank-1 vs 2-D taps, packed vs scalar accumulate, resize tile widths and
tap counts, JPEG padding to x16, vector vs scalar tails."""

import random

import numpy as np
import pytest
import torch

import trainner_redux_b200 as T
from oracle import kernel_synth_oracle as KS
from oracle import otf_oracle as O
from trainner_redux_b200 import degradations as D
from trainner_redux_b200.kernels import KernelOptions, draw_kernel_params

pytestmark = pytest.mark.gpu
LSB = 1 / 255


def _shape(rng, big=False):
    b = rng.choice([1, 2, 3, 5])
    if big:
        return b, rng.choice([64, 96, 128, 200, 256, 320]), rng.choice([64, 100, 128, 192, 256, 300])
    return b, rng.randint(24, 150), rng.randint(24, 150)


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_filter2d(seed, dev):
    rng = random.Random(seed)
    b, h, w = _shape(rng, big=seed % 2 == 0)
    c = rng.choice([1, 3])
    img = torch.rand(b, c, h, w, generator=torch.Generator().manual_seed(seed))
    kopt = KernelOptions(kernel_range=(3, 21), kernel_range2=(3, 21), final_kernel_range=(3, 21), sinc_prob=0.2, sinc_prob2=0.2,
                         final_sinc_prob=0.7, blur_sigma=(0.2, 3), blur_sigma2=(0.2, 1.5))
    p1, p2, p3 = draw_kernel_params(kopt, b, random.Random(seed), np.random.default_rng(seed))
    for p in (p1, p2, p3):
        k = torch.from_numpy(KS.synthesize(p))
        if rng.random() < 0.3:
            k = k[:1]  # shared kernel
        got = T.filter2d(img.to(dev), k.to(dev)).cpu()
        d = (got - O.filter2d(img, k)).abs().max().item()
        assert d <= 1e-5, f"seed {seed} shape {(b, c, h, w)} kernel batch {k.size(0)}: {d:.2e}"


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_resize(seed, dev):
    rng = random.Random(100 + seed)
    b, h, w = _shape(rng, big=seed % 3 == 0)
    img = torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(seed))
    for mode in O.RESIZE_MODES:
        if rng.random() < 0.5:
            s = rng.uniform(0.3, 1.6)
            want = O.resize_pt(img, mode, scale_factor=s)
            got = T.resize_pt(img.to(dev), mode, scale_factor=s).cpu()
            what = f"s={s:.3f}"
        else:
            size = (rng.randint(8, 2 * h), rng.randint(8, 2 * w))
            want = O.resize_pt(img, mode, size=size)
            got = T.resize_pt(img.to(dev), mode, size=size).cpu()
            what = f"size={size}"
        assert got.shape == want.shape
        d = (got - want).abs().max().item()
        assert d <= 1e-5, f"seed {seed} {mode} {(b, h, w)} {what}: {d:.2e}"


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_noise_and_jpeg(seed, dev):
    rng = random.Random(200 + seed)
    b, h, w = _shape(rng)
    g = torch.Generator().manual_seed(seed)
    img = O.synth_gt(b, h, w, "natural", seed=seed) if seed % 2 else torch.rand(b, 3, h, w, generator=g)
    sigma = torch.rand(b, generator=g) * 29 + 1
    gray = (torch.rand(b, generator=g) < 0.5).float()
    ncol, ngray = torch.randn(b, 3, h, w, generator=g), torch.randn(h, w, generator=g)
    use_gray = bool(gray.sum() > 0)
    want = O.add_gaussian_noise(img, sigma, gray, ncol, ngray if use_gray else None)
    got = D.add_gaussian_noise_pt(img.to(dev), sigma.to(dev), gray.to(dev), noise=ncol.to(dev),
                                  noise_gray=ngray.to(dev) if use_gray else None).cpu()
    assert (got - want).abs().max().item() <= 1e-7, f"gaussian seed {seed}"
    # Poisson with the oracle's own counts
    scale = torch.rand(b, generator=g) * 2.9 + 0.05
    qc, vc, lc, qg, vg, lg = O.poisson_lambda(img)
    cc, cg = torch.poisson(lc, generator=g), torch.poisson(lg, generator=g)
    want = O.add_poisson_noise(img, scale, gray, counts_color=cc, counts_gray=cg if use_gray else None)
    exp = {}
    got = D.add_poisson_noise_pt(img.to(dev), scale.to(dev), True, False, gray.to(dev), poisson_counts=cc.to(dev),
                                 poisson_counts_gray=cg.to(dev) if use_gray else None, _export=exp).cpu()
    assert torch.equal(exp["vals"][:, 0].cpu(), vc.view(-1)), f"vals seed {seed}"
    assert (got - want).abs().max().item() <= 1e-6, f"poisson seed {seed}"
    # DiffJPEG, both rounding modes
    q = torch.rand(b, generator=g) * 65 + 30
    for diff in (False, True):
        want = O.diffjpeg(img, q.clone(), diff)
        got = T.DiffJPEG(differentiable=diff)(img.to(dev), quality=q.clone().to(dev)).cpu()
        frac = ((got - want).abs() <= LSB + 1e-6).float().mean().item()
        assert frac >= 0.999, f"jpeg seed {seed} diff={diff} {(b, h, w)}: {frac:.5f}"


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_usm(seed, dev):
    rng = random.Random(300 + seed)
    b, h, w = rng.choice([1, 2]), rng.randint(30, 140), rng.randint(30, 140)
    radius = rng.choice([r for r in (3, 7, 9, 15, 25, 50) if r // 2 + 1 < min(h, w)])
    img = O.synth_gt(b, h, w, "natural", seed=seed)
    wgt, thr = rng.uniform(0.3, 1.0), rng.choice([4, 10])
    want = O.usm_sharp(img, O.usm_kernel(radius), wgt, thr)
    got = T.USMSharp(radius=radius).to(dev)(img.to(dev), wgt, thr).cpu()
    d = (got - want).abs()
    # `abs(res)*255 > threshold` is a cliff.  The separable evaluation differs from the reference's
    # 2-D sum by ~1e-7, so a pixel whose residual sits within ~1e-4 grey levels of the threshold may flip its mask
    # bit; the blurred soft mask spreads one flip over K x K pixels at <= ~1e-3.  Find such pixels in float64.
    k64 = O.usm_kernel(radius).double()
    res64 = img.double() - O.filter2d(img.double(), k64)
    ambiguous = int(((res64.abs() * 255 - thr).abs() < 2e-4).sum())
    if ambiguous == 0:
        assert (d <= 1e-5).float().mean().item() >= 0.999 and d.max().item() <= 3e-4, f"usm seed {seed} r={radius}: {d.max():.2e}"
    else:
        k = radius + (radius % 2 == 0)
        budget = min(1.0, ambiguous * k * k / d.numel())  # pixels a flip can reach
        assert (d > 1e-5).float().mean().item() <= budget + 1e-3 and d.max().item() <= 3e-3, \
            f"usm seed {seed} r={radius}: {ambiguous} threshold-ambiguous pixels, max {d.max():.2e}"
