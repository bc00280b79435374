"""Size-independent properties at BASELINE.json's FULL sizes (B=64 of 256^2 x4, B=32 of 512^2 x2),
where the CPU oracle would take minutes: linearity and identities of the filters, idempotence of
the 8-bit lattice, exact resampling identities, zero-noise limits, crop/window consistency, chain
shape/lattice invariants, and spot checks of full-size outputs against the oracle on a few planes."""

import pytest
import torch

import trainner_redux_b200 as T
from oracle import otf_oracle as O
from trainner_redux_b200 import degradations as D
from trainner_redux_b200.realesrgan_feed import HostRNG, OTFOptions, RealESRGANFeed, clamp_round, draw_plan
from trainner_redux_b200.transforms import crop_pair

pytestmark = pytest.mark.gpu
CONFIGS = [(64, 256, 4), (32, 512, 2)]


def rand(b, h, w, seed, dev):
    return torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(seed)).to(dev)


@pytest.mark.parametrize("b,size,scale", CONFIGS)
def test_filter2d_linearity_identity_and_spot_check(b, size, scale, dev):
    x, y = rand(b, size, size, 1, dev), rand(b, size, size, 2, dev)
    k_cpu = O.synth_blur_kernels(b, seed=3)
    k = k_cpu.to(dev)
    fx, fy, fxy = T.filter2d(x, k), T.filter2d(y, k), T.filter2d(0.25 * x + 0.5 * y, k)
    assert (fxy - (0.25 * fx + 0.5 * fy)).abs().max().item() < 1e-5, "linearity"
    pulse = torch.zeros(b, 21, 21, device=dev)
    pulse[:, 10, 10] = 1
    assert torch.equal(T.filter2d(x, pulse), x), "pulse kernel is the identity"
    # kernels sum to 1: a constant image is a fixed point (reflect padding included)
    c = torch.full((b, 3, size, size), 0.37, device=dev)
    assert (T.filter2d(c, k) - 0.37).abs().max().item() < 2e-6
    # spot check three samples of the full-size result against the oracle
    for i in (0, b // 2, b - 1):
        want = O.filter2d(x[i : i + 1].cpu(), k_cpu[i : i + 1])
        assert (fx[i : i + 1].cpu() - want).abs().max().item() < 1e-5, f"sample {i}"


@pytest.mark.parametrize("b,size,scale", CONFIGS)
def test_resize_identities(b, size, scale, dev):
    x = rand(b, size, size, 4, dev)
    for mode in ("bilinear", "bicubic", "area", "nearest-exact", "lanczos"):
        same = T.resize_pt(x, mode, size=(size, size))
        assert (same - x).abs().max().item() < 1e-6, f"{mode}: same-size resize is the identity"
    # area x1/2 == 2x2 mean; nearest-exact x2 == pixel replication
    half = T.resize_pt(x, "area", scale_factor=0.5)
    ref = torch.nn.functional.avg_pool2d(x, 2)
    assert (half - ref).abs().max().item() < 1e-6
    up = T.resize_pt(x[:4], "nearest-exact", scale_factor=2)
    assert torch.equal(up, x[:4].repeat_interleave(2, 2).repeat_interleave(2, 3))
    # constants are preserved by every normalised kernel
    c = torch.full((4, 3, size, size), 0.61, device=dev)
    for mode in ("bilinear", "bicubic", "area", "lanczos"):
        out = T.resize_pt(c, mode, size=(size // scale, size // scale + 3))
        assert (out - 0.61).abs().max().item() < 2e-6, mode
    # spot check against the oracle
    for mode, sz in (("bicubic", (int(size * 0.75),) * 2), ("bilinear", (size // scale,) * 2)):
        got = T.resize_pt(x[:2], mode, size=sz)
        assert (got.cpu() - O.resize_pt(x[:2].cpu(), mode, size=sz)).abs().max().item() < 1e-5, mode


@pytest.mark.parametrize("b,size,scale", CONFIGS)
def test_lattice_noise_limits_and_crop(b, size, scale, dev):
    x = rand(b, size, size, 5, dev) * 1.2 - 0.1
    q = clamp_round(x)
    assert torch.equal(clamp_round(q), q), "8-bit lattice is idempotent"
    assert torch.equal((q * 255).round(), q * 255) and q.min() >= 0 and q.max() <= 1
    zero = torch.zeros(b, device=dev)
    assert torch.equal(D.add_gaussian_noise_pt(x, zero, zero), x.clamp(0, 1)), "sigma = 0 -> clamp only"
    assert torch.equal(D.add_poisson_noise_pt(x.clamp(0, 1), zero, True, False, zero), x.clamp(0, 1)), "scale = 0 -> unchanged"
    lq = rand(b, size // scale, size // scale, 6, dev)
    p = (size // scale) * 7 // 8 // 4 * 4
    g, l = crop_pair(x, lq, p * scale, scale, 3, 5)
    assert torch.equal(g, x[:, :, 3 * scale : 3 * scale + p * scale, 5 * scale : 5 * scale + p * scale])
    assert torch.equal(l, lq[:, :, 3 : 3 + p, 5 : 5 + p])


@pytest.mark.parametrize("b,size,scale", CONFIGS)
def test_diffjpeg_properties(b, size, scale, dev):
    x = O.synth_gt(b, size, size, "natural", seed=7).to(dev)
    jp = T.DiffJPEG(differentiable=False)
    q = torch.linspace(30, 95, b, device=dev)
    y = jp(x, quality=q.clone())
    assert y.shape == x.shape and y.min() >= 0 and y.max() <= 1
    # higher quality -> smaller error, sample by sample (quality is per sample)
    err = (y - x).abs().flatten(1).mean(1)
    assert err[0] > err[-1] and (err[: b // 2].mean() > err[b // 2 :].mean())
    # a flat image has only a DC coefficient; it survives up to the DC quantiser's step
    flat = torch.full((4, 3, size, size), 0.5, device=dev)
    assert (jp(flat, quality=90.0) - 0.5).abs().max().item() < 4.0 / 255
    # spot check against the oracle
    want = O.diffjpeg(x[:2].cpu(), q[:2].cpu().clone(), False)
    d = (y[:2].cpu() - want).abs()
    assert (d <= 1 / 255 + 1e-6).float().mean().item() >= 0.999


@pytest.mark.parametrize("b,size,scale", CONFIGS)
def test_full_size_chain_invariants(b, size, scale, dev):
    modes = ["bilinear", "bicubic", "area"]
    opt = OTFOptions(scale=scale, gt_size=size - 32, blur_prob=1, blur_prob2=1, gaussian_noise_prob=0.5, noise_range=(1, 30),
                     poisson_scale_range=(0.05, 3), gray_noise_prob=0.4, gaussian_noise_prob2=0.5, noise_range2=(1, 25),
                     poisson_scale_range2=(0.05, 2.5), gray_noise_prob2=0.4, jpeg_range=(30, 95), jpeg_range2=(30, 95),
                     resize_mode_list=modes, resize_mode_prob=[1 / 3] * 3, resize_mode_list2=modes, resize_mode_prob2=[1 / 3] * 3,
                     resize_mode_list3=modes, resize_mode_prob3=[1 / 3] * 3, queue_size=2 * b)
    data = {"gt": O.synth_gt(b, size, size, "natural", seed=9), "kernel1": O.synth_blur_kernels(b, seed=1),
            "kernel2": O.synth_blur_kernels(b, seed=2), "sinc_kernel": O.synth_sinc_or_pulse(b, seed=3)}
    feed = RealESRGANFeed(opt, device=dev, manual_seed=3)
    for step in range(4):
        feed.feed_data(data)
        g = size - 32
        assert tuple(feed.gt.shape) == (b, 3, g, g) and tuple(feed.lq.shape) == (b, 3, g // scale, g // scale)
        assert feed.lq.is_contiguous() and torch.isfinite(feed.lq).all()
        lat = feed.lq * 255
        assert torch.equal(lat, lat.round()) and feed.lq.min() >= 0 and feed.lq.max() <= 1
        # the LQ is a degraded view of the GT crop: strongly correlated with its area-downscaled version
        ref = torch.nn.functional.avg_pool2d(feed.gt, scale)
        corr = torch.corrcoef(torch.stack([ref.flatten(), feed.lq.flatten()]))[0, 1].item()
        assert corr > 0.5, corr
