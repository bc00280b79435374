"""Distribution tests of the kernels' Philox4x32-10 generators (the reference's random fields come
from torch's generator, so the GPU RNG cannot be compared value by value — north_star asks for
distribution tests): N(0,1) moments + KS, uniform moments, Poisson chi-square at
lambda in {0.5, 4, 40, 200}, per-sample sigma scaling, the batch-shared gray field, reproducibility
from (seed, offset)."""

import ctypes as C

import numpy as np
import pytest
import torch

from trainner_redux_b200 import _lib
from trainner_redux_b200 import degradations as D

pytestmark = pytest.mark.gpu
scipy_stats = pytest.importorskip("scipy.stats")


def _fill(kind: str, n: int, seed: int, offset: int, dev) -> torch.Tensor:
    out = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.call(f"otf_philox_{kind}_f32", _lib.ptr(out), n, seed, offset, _lib.stream())
    return out


def test_normal_moments_and_ks(dev):
    n = 4_000_000
    x = _fill("normal", n, 1234, 0, dev).double().cpu().numpy()
    assert abs(x.mean()) < 4 / np.sqrt(n)
    assert abs(x.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs(scipy_stats.skew(x)) < 0.01 and abs(scipy_stats.kurtosis(x)) < 0.02
    d = scipy_stats.kstest(x[:1_000_000], "norm").statistic
    assert d < 2.0 / np.sqrt(1_000_000), d
    assert np.abs(x).max() < 6.5  # Box-Muller with 24-bit uniforms tops out at sqrt(-2 ln 2^-24) = 5.77


def test_uniform_moments(dev):
    n = 4_000_000
    u = _fill("uniform", n, 99, 3, dev).double().cpu().numpy()
    assert u.min() > 0 and u.max() <= 1
    assert abs(u.mean() - 0.5) < 4 * np.sqrt(1 / 12 / n)
    assert abs(u.var() - 1 / 12) < 1e-3
    assert scipy_stats.kstest(u[:1_000_000], "uniform").statistic < 2.0 / 1000


@pytest.mark.parametrize("lam", [0.5, 4.0, 9.5, 40.0, 200.0, 255.0])
def test_poisson_chi_square(dev, lam):
    n = 2_000_000
    lam_t = torch.full((n,), lam, dtype=torch.float32, device=dev)
    out = torch.empty_like(lam_t)
    _lib.call("otf_philox_poisson_f32", _lib.ptr(lam_t), _lib.ptr(out), n, 7, 1, _lib.stream())
    k = out.cpu().numpy().astype(np.int64)
    assert k.min() >= 0
    assert abs(k.mean() - lam) < 5 * np.sqrt(lam / n)
    assert abs(k.var() - lam) < 0.01 * lam + 0.01
    kmax = int(k.max())
    obs = np.bincount(k, minlength=kmax + 1).astype(np.float64)
    exp = scipy_stats.poisson.pmf(np.arange(kmax + 1), lam) * n
    keep = exp >= 50
    obs_k, exp_k = obs[keep], exp[keep]
    obs_k = np.append(obs_k, n - obs_k.sum())
    exp_k = np.append(exp_k, n - exp_k.sum())
    if exp_k[-1] < 1:
        obs_k, exp_k = obs_k[:-1], exp_k[:-1]
    chi2 = ((obs_k - exp_k) ** 2 / exp_k).sum()
    dof = len(exp_k) - 1
    assert chi2 < scipy_stats.chi2.ppf(0.9999, dof), (chi2, dof)


def test_gaussian_sigma_scaling_and_gray_sharing(dev):
    b, h, w = 4, 128, 128
    img = torch.zeros(b, 3, h, w, device=dev)
    sigma = torch.tensor([5.0, 10.0, 20.0, 40.0], device=dev)
    gen = D.PhiloxState(5)
    n_col = D.generate_gaussian_noise_pt(img, sigma, torch.zeros(b, device=dev), generator=gen)
    std = n_col.flatten(1).std(1).cpu() * 255
    assert torch.allclose(std, sigma.cpu(), rtol=0.02), std
    # colour noise is independent across channels and samples
    c = torch.corrcoef(torch.stack([n_col[0, 0].flatten(), n_col[0, 1].flatten(), n_col[1, 0].flatten()]))
    assert c[0, 1].abs() < 0.02 and c[0, 2].abs() < 0.02
    # gray noise: ONE (h,w) field shared by the whole batch (degradations.py:593-596), scaled per sample
    n_gray = D.generate_gaussian_noise_pt(img, sigma, torch.ones(b, device=dev), generator=gen)
    base = n_gray[0, 0] / sigma[0]
    for bi in range(b):
        for ch in range(3):
            assert torch.allclose(n_gray[bi, ch] / sigma[bi], base, atol=1e-6)
    # mixed flags: flagged samples get the shared field, the others their own colour noise
    flags = torch.tensor([1.0, 0.0, 1.0, 0.0], device=dev)
    gen2 = D.PhiloxState(5, offset=gen.offset)
    mixed = D.generate_gaussian_noise_pt(img, sigma, flags, generator=gen2)
    assert torch.equal(mixed[0, 0], mixed[0, 2]) and torch.allclose(mixed[0, 0] / sigma[0], mixed[2, 1] / sigma[2], atol=1e-6)
    assert not torch.equal(mixed[1, 0], mixed[1, 1])


def test_reproducible_from_seed_and_offset(dev):
    img = torch.rand(2, 3, 64, 48, device=dev)
    a = D.random_add_gaussian_noise_pt  # draws sigma/gray with torch.rand: fix torch's generator too
    torch.manual_seed(0)
    x1 = a(img, (1, 30), 0.5, generator=D.PhiloxState(11, 4))
    torch.manual_seed(0)
    x2 = a(img, (1, 30), 0.5, generator=D.PhiloxState(11, 4))
    torch.manual_seed(0)
    x3 = a(img, (1, 30), 0.5, generator=D.PhiloxState(11, 5))
    assert torch.equal(x1, x2) and not torch.equal(x1, x3)
    p1 = D.add_poisson_noise_pt(img, 1.0, True, False, torch.ones(2, device=dev), generator=D.PhiloxState(3, 0))
    p2 = D.add_poisson_noise_pt(img, 1.0, True, False, torch.ones(2, device=dev), generator=D.PhiloxState(3, 0))
    assert torch.equal(p1, p2)


def test_poisson_noise_statistics_on_flat_image(dev):
    """Flat image -> one distinct level -> vals = 1 -> lambda = q; noise = Poisson(q) - q scaled."""
    q = 128 / 255
    img = torch.full((2, 3, 256, 256), q, device=dev)
    export = {}
    out = D.add_poisson_noise_pt(img, 0.1, False, False, 0, generator=D.PhiloxState(21), _export=export)
    assert export["vals"][:, 0].tolist() == [1.0, 1.0]
    noise = (out - img) / 0.1
    assert abs(noise.mean().item()) < 0.01 and abs(noise.var().item() - q) < 0.02



def _counts_through_noise_api(levels_img, dev, gray_flag, seed):
    """Poisson counts recovered from the production path (generate_poisson_noise_pt: presence masks, CDF tables,
    table inversion): noise = cnt / vals - q with scale 1."""
    from trainner_redux_b200 import degradations as D

    gen = D.PhiloxState(seed)
    b = levels_img.size(0)
    gray = torch.full((b,), float(gray_flag), device=dev)
    noise = D.generate_poisson_noise_pt(levels_img, 1.0, gray, generator=gen)
    return noise


@pytest.mark.parametrize("gray_flag", [0, 1])
def test_poisson_table_inversion_distribution(dev, gray_flag):
    """The production sampler (exact table inversion, one uniform per sample) through the public noise API: an image
    that holds all 256 levels (so vals = 256 and lambda_L = L/255*256), chi-square of the counts at several levels."""
    import math

    n_per = 4096  # pixels per level and channel
    lv = torch.arange(256, dtype=torch.float32).repeat_interleave(n_per)  # (256 * n_per,)
    h, w = 1024, 256 * n_per // 1024
    plane = (lv / 255.0).view(1, 1, h, w)
    img = plane.repeat(2, 3, 1, 1).to(dev)  # gray of an r=g=b image is the same level (0.2989+0.587+0.114 = 1.0009 -> rounds back)
    noise = _counts_through_noise_api(img, dev, gray_flag, seed=77)
    q = torch.round(plane * 255) / 255
    cnt = torch.round((noise[:, 0:1].cpu() + q) * 256).view(2, 256, n_per)  # vals = 256
    if gray_flag:  # one field for the three channels
        assert torch.equal(noise[:, 0], noise[:, 1]) and torch.equal(noise[:, 1], noise[:, 2])
    else:
        assert not torch.equal(noise[:, 0], noise[:, 1])
    for level in (1, 4, 10, 40, 128, 200, 255):
        lam = float(torch.tensor(level / 255.0, dtype=torch.float32) * 256)
        x = cnt[:, level].flatten().numpy()
        n = x.size
        lo, hi = max(0, int(lam - 5 * math.sqrt(lam) - 2)), int(lam + 5 * math.sqrt(lam) + 3)
        # chi-square over the integer bins [lo, hi) + two tail bins
        ks = np.arange(lo, hi)
        logp = -lam + ks * math.log(lam) - np.array([math.lgamma(k + 1) for k in ks])
        p = np.exp(logp)
        obs = np.array([(x == k).sum() for k in ks], dtype=np.float64)
        p_tail = max(1.0 - p.sum(), 1e-12)
        obs_tail = n - obs.sum()
        keep = p * n >= 5
        chi2 = (((obs[keep] - p[keep] * n) ** 2) / (p[keep] * n)).sum()
        rest_p, rest_o = p[~keep].sum() + p_tail, obs[~keep].sum() + obs_tail
        if rest_p * n >= 5:
            chi2 += (rest_o - rest_p * n) ** 2 / (rest_p * n)
        dof = int(keep.sum())
        assert chi2 < dof + 5 * math.sqrt(2 * dof) + 10, (level, lam, chi2, dof)
        assert abs(x.mean() - lam) < 5 * math.sqrt(lam / n) + 1e-3, (level, x.mean(), lam)
    assert torch.all(cnt[:, 0] == 0)  # lambda 0


def test_poisson_table_and_rejection_paths_agree_in_law(dev, monkeypatch):
    """Same image through both samplers: different streams, same distribution (mean / variance per level band)."""
    from trainner_redux_b200 import degradations as D

    img = torch.rand(4, 3, 256, 256, device=dev)
    a = D.generate_poisson_noise_pt(img, 1.0, 0, generator=D.PhiloxState(5))
    small = D._poisson(img, 1.0, 0, False, False, False, generator=D.PhiloxState(5), export={})  # export -> rejection sampler
    for lo in (0.0, 0.25, 0.5, 0.75):
        m = (img >= lo) & (img < lo + 0.25)
        assert abs(a[m].mean().item() - small[m].mean().item()) < 2e-3
        assert abs(a[m].var().item() / small[m].var().item() - 1) < 0.03
    again = D.generate_poisson_noise_pt(img, 1.0, 0, generator=D.PhiloxState(5))
    assert torch.equal(a, again)  # reproducible from (seed, offset)
