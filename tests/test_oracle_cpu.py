"""CPU tests of the oracle itself: it must reproduce the golden vectors that
oracle/make_goldens.py produced by running the REFERENCE's Python (bit for bit on the machine
that generated them; a few ulp elsewhere because BLAS/oneDNN kernels differ by CPU)."""

import os

import numpy as np
import pytest
import torch

from oracle import otf_oracle as O

FP = 2e-6  # cross-machine slack for fp32 stages
LSB = 1 / 255


def close(a, b, tol=FP):
    assert a.shape == b.shape
    d = (a.contiguous() - b).abs().max().item()
    assert d <= tol, f"max-abs {d:.3e}"


def lsb(a, b, frac=0.999):
    d = (a.contiguous() - b).abs()
    assert (d <= LSB + 1e-6).float().mean().item() >= frac, f"max {d.max().item():.4f}"


def test_filter2d(golden):
    g = golden
    close(O.filter2d(g["img"], g["f2d_k21"]), g["f2d_out21"])
    close(O.filter2d(g["img_u"], g["f2d_k5"]), g["f2d_out5"])
    close(O.filter2d(g["img_u"], g["f2d_sinc"]), g["f2d_outsinc"])
    with pytest.raises(ValueError, match="Wrong kernel size"):
        O.filter2d(g["img"], torch.ones(2, 4, 4))


def test_usm_and_cv_kernel(golden):
    g = golden
    for r in (50, 7):
        close(O.usm_sharp(g["usm_img"], O.usm_kernel(r)), g[f"usm_out_r{r}"], 1e-4)
    close(O.usm_sharp(g["usm_img"], O.usm_kernel(50), 0.8, 4), g["usm_out_w08_t4"], 1e-4)
    cv2 = pytest.importorskip("cv2")
    for k in (1, 3, 5, 7, 9, 11, 25, 51):
        # OpenCV evaluates exp() with its own bit-exact soft-float: <= 1 ulp of float64 apart
        assert np.abs(O.cv_gaussian_kernel_1d(k) - cv2.getGaussianKernel(k, 0)).max() < 1e-15, k
        assert torch.equal(O.usm_kernel(k), torch.FloatTensor(cv2.getGaussianKernel(k, 0) @ cv2.getGaussianKernel(k, 0).T)[None]) or k > 9
    assert abs(0.3 * ((51 - 1) * 0.5 - 1) + 0.8 - 8.0) < 1e-12


@pytest.mark.parametrize("mode", O.RESIZE_MODES)
def test_resize(golden, mode):
    x = golden["img_u"]
    for s in (0.4, 0.75, 1.25, 1.5):
        close(O.resize_pt(x, mode, scale_factor=s), golden[f"rs_{mode}_s{s}"])
    for size in ((10, 9), (40, 36), (17, 50)):
        close(O.resize_pt(x, mode, size=size), golden[f"rs_{mode}_{size[0]}x{size[1]}"])
    with pytest.raises(ValueError, match="scale_factor or size is required"):
        O.resize_pt(x, mode)


@pytest.mark.parametrize("tag", ["mixed", "color", "allgray"])
def test_gaussian(golden, tag):
    g = golden
    got = O.add_gaussian_noise(g["img"], g[f"gn_{tag}_sigma"], g[f"gn_{tag}_gray"], g[f"gn_{tag}_ncol"], g.get(f"gn_{tag}_ngray"))
    close(got, g[f"gn_{tag}_out"], 1e-7)


def test_gaussian_gray_field_is_shared_by_the_batch():
    img = torch.zeros(3, 3, 8, 8) + 0.5
    ng, nc = torch.randn(8, 8), torch.randn(3, 3, 8, 8)
    f = O.gaussian_noise_field(img, torch.tensor([10.0, 10.0, 20.0]), torch.ones(3), nc, ng)
    assert torch.equal(f[0], f[1]) and torch.allclose(f[2], 2 * f[0]) and torch.equal(f[0, 0], f[0, 2])


@pytest.mark.parametrize("tag", ["mixed", "color", "allgray", "twolevel", "flat"])
def test_poisson(golden, tag):
    g = golden
    img = g[f"pn_{tag}_img"]
    qc, vc, lc, qg, vg, lg = O.poisson_lambda(img)
    assert torch.equal(vc.view(-1), g[f"pn_{tag}_vals_color"]) and torch.equal(vg.view(-1), g[f"pn_{tag}_vals_gray"])
    got = O.add_poisson_noise(img, g[f"pn_{tag}_scale"], g[f"pn_{tag}_gray"], counts_color=g[f"pn_{tag}_counts_color"],
                              counts_gray=g.get(f"pn_{tag}_counts_gray"))
    close(got, g[f"pn_{tag}_out"], 1e-7)


def test_poisson_vals_edge_cases():
    flat = O.synth_gt(1, 8, 8, "flat")
    two = O.synth_gt(1, 16, 16, "twolevel")
    assert O.poisson_vals(O.quantise8(flat)).item() == 1 and O.poisson_vals(O.quantise8(two)).item() == 2
    ramp = (torch.arange(256).float() / 255).view(1, 1, 16, 16).repeat(1, 3, 1, 1)
    assert O.poisson_vals(O.quantise8(ramp)).item() == 256
    assert O.poisson_vals(O.quantise8(ramp[..., :9, :])).item() == 256  # 144 distinct -> 256


@pytest.mark.parametrize("diff", [False, True])
def test_diffjpeg(golden, diff):
    g = golden
    lsb(O.diffjpeg(g["jpg_img"], g["jpg_t_q"].clone(), diff), g[f"jpg_t_out_d{int(diff)}"])
    lsb(O.diffjpeg(g["jpg_img2"], g["jpg_u_q"].clone(), diff), g[f"jpg_u_out_d{int(diff)}"])
    lsb(O.diffjpeg(g["jpg_img"], 50, diff), g[f"jpg_s50_out_d{int(diff)}"])
    q = g["jpg_t_q"].clone()
    O.diffjpeg(g["jpg_img"], q, diff)
    assert torch.equal(q, g["jpg_t_factor"])  # quirk Q1: overwritten with factors


def test_diffjpeg_tables_and_separable_dct():
    y, c = O.jpeg_tables()
    assert y[0, :4].tolist() == [16, 12, 14, 14] and torch.equal(c, c.T) and c[4, 4] == 99
    # the 4-D tensordot DCT of the reference equals the orthonormal separable DCT-II the kernel uses
    blk = torch.rand(5, 8, 8, generator=torch.Generator().manual_seed(0)) * 255 - 128
    ref = O._DCT_SCALE * torch.tensordot(blk, O._DCT_T, dims=2)
    n = torch.arange(8).float()
    cm = 0.5 * torch.cos((2 * n[None, :] + 1) * n[:, None] * torch.pi / 16)
    cm[0] *= 2 ** -0.5
    assert (cm @ blk @ cm.T - ref).abs().max() < 5e-4  # coefficients reach ~1e3: a few fp32 ulp
    assert O.quality_to_factor(30) == pytest.approx(5000 / 30 / 100) and O.quality_to_factor(80) == pytest.approx(0.4)


def test_clamp_round_crop_and_chain(golden):
    g = golden
    assert torch.equal(O.clamp_round(g["cr_in"]), g["cr_out"])
    assert torch.equal(O.clamp_round(torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255])) * 255, torch.tensor([0.0, 2.0, 2.0]))
    plan = {
        "scale": 4, "gt_size": 48, "blur1": True, "resize1": {"scale": 0.75, "mode": "bicubic"},
        "noise1": {"kind": "gaussian", "sigma": g["chain_sigma1"], "gray": g["chain_gray1"]}, "jpeg1": g["chain_q1"],
        "blur2": True, "resize2": {"scale": 1.1, "mode": "bilinear"},
        "noise2": {"kind": "poisson", "scale": g["chain_scale2"], "gray": g["chain_gray2"]},
        "final_order": "resize_first", "resize3_mode": "area", "jpeg2": g["chain_q2"],
        "crop": tuple(int(v) for v in g["chain_crop"]),
    }
    noise = {"noise1_color": g["chain_n1c"], "noise1_gray": g["chain_n1g"], "noise2_counts_color": g["chain_cc2"],
             "noise2_counts_gray": g["chain_cg2"]}
    gt_c, lq_c = O.run_chain_b(g["chain_gt"], g["chain_k1"], g["chain_k2"], g["chain_sinc"], plan, noise)
    assert torch.equal(gt_c, g["chain_gt_crop"])
    lsb(lq_c, g["chain_lq"], 0.999)
    with pytest.raises(ValueError):
        O.paired_crop(g["chain_gt"], g["chain_lq_full"][:, :, :15], 48, 4, 0, 0)


def test_pool_semantics():
    pool = O.PairPool(4)
    for step in range(4):
        lq = torch.full((2, 1, 1, 1), float(step))
        gt = torch.full((2, 1, 2, 2), float(step))
        perm = torch.tensor([3, 2, 1, 0])
        out_lq, out_gt = pool.step(lq, gt, perm)
        if step < 2:
            assert torch.equal(out_lq, lq)  # filling: pass through
        else:
            assert out_lq.flatten().tolist() != lq.flatten().tolist()
    with pytest.raises(AssertionError):
        O.PairPool(5).step(torch.zeros(2, 1, 1, 1), torch.zeros(2, 1, 1, 1))


def test_kernel_synthesis_oracle_and_draw_order(golden):
    """Row f2: the numpy restatement reproduces the reference generators bit for bit, and the host-side draw
    order of trainner_redux_b200.kernels equals the dataset's (realesrgan_dataset.py:149-206)."""
    import random

    from oracle import kernel_synth_oracle as KS
    from trainner_redux_b200.kernels import KernelOptions, draw_kernel_params

    g = golden
    got = KS.synthesize(g["ks_params"].numpy())
    assert np.abs(got - g["ks_ref"].numpy()).max() < 1e-9  # bit-identical where the numpy/scipy builds match
    assert np.allclose(got.sum((1, 2)), 1.0, atol=1e-6)
    kopt = KernelOptions(sinc_prob=0.1, sinc_prob2=0.1, final_sinc_prob=0.8, kernel_range=(7, 21), kernel_range2=(7, 21),
                         final_kernel_range=(7, 21))
    p1, p2, p3 = draw_kernel_params(kopt, 12, random.Random(78), np.random.default_rng(77))
    for name, prm in (("k1", p1), ("k2", p2), ("sinc", p3)):
        assert np.abs(KS.synthesize(prm) - g[f"ks_ds_{name}"].numpy()).max() < 1e-9, name
    assert set(p1[:, 0].astype(int)) <= set(range(7)) and (p3[:, 0] >= 6).all()


def test_config1_plumbing_batch8_256_scale4():
    """BASELINE.json configs[0] (SURVEY.md §8d Config 1): batch 8 of synthetic 256x256 RGB GT, scale 4, the reference
    order on the CPU — the host draws of the product's planner feed the oracle chain; shapes, value ranges, the 8-bit
    lattice, the GT crop window and the pair-pool semantics are what `RealESRGANModel.feed_data` promises."""
    from trainner_redux_b200.realesrgan_feed import HostRNG, OTFOptions, draw_plan

    torch.set_num_threads(min(8, os.cpu_count() or 1))
    b, h, sc, gsz = 8, 256, 4, 224
    opt = OTFOptions(scale=sc, gt_size=gsz, queue_size=16, blur_prob=1.0, blur_prob2=0.8, gaussian_noise_prob=1.0, noise_range=(1, 30),
                     gray_noise_prob=0.4, gaussian_noise_prob2=1.0, noise_range2=(1, 25), gray_noise_prob2=0.4, jpeg_range=(30, 95),
                     jpeg_range2=(30, 95), resize_mode_list=("bilinear", "bicubic", "area"), resize_mode_prob=(1 / 3,) * 3,
                     resize_mode_list2=("bilinear", "bicubic", "area"), resize_mode_prob2=(1 / 3,) * 3,
                     resize_mode_list3=("bilinear", "bicubic", "area"), resize_mode_prob3=(1 / 3,) * 3)
    rng = HostRNG(0)
    pool = O.PairPool(opt.queue_size)
    k1, k2, sk = O.synth_blur_kernels(b, seed=1), O.synth_blur_kernels(b, seed=2), O.synth_sinc_or_pulse(b, seed=3)
    g = torch.Generator().manual_seed(0)
    outs = []
    for step in range(3):
        gt = O.synth_gt(b, h, h, "natural", seed=step)
        plan = draw_plan(opt, b, h, h, rng)
        assert set(plan) >= {"blur1", "resize1", "noise1", "jpeg1", "blur2", "resize2", "noise2", "final_order", "resize3_mode", "jpeg2", "crop"}
        h1, h2 = round(h * plan["resize1"]["scale"]), int(h / sc * plan["resize2"]["scale"])
        noise = {"noise1_color": torch.randn(b, 3, h1, h1, generator=g), "noise1_gray": torch.randn(h1, h1, generator=g),
                 "noise2_color": torch.randn(b, 3, h2, h2, generator=g), "noise2_gray": torch.randn(h2, h2, generator=g)}
        gt_c, lq_c = O.run_chain_b(gt, k1, k2, sk, plan, noise)
        assert tuple(gt_c.shape) == (b, 3, gsz, gsz) and tuple(lq_c.shape) == (b, 3, gsz // sc, gsz // sc)
        top, left = plan["crop"]
        assert torch.equal(gt_c, gt[:, :, top * sc : top * sc + gsz, left * sc : left * sc + gsz])  # one offset for the whole batch
        assert lq_c.min() >= 0 and lq_c.max() <= 1 and torch.equal(lq_c, torch.round(lq_c * 255) / 255)
        lq_o, gt_o = pool.step(lq_c, gt_c, torch.randperm(opt.queue_size, generator=g))
        outs.append((lq_c, lq_o))
        assert lq_o.shape == lq_c.shape and gt_o.shape == gt_c.shape
    assert torch.equal(outs[0][0], outs[0][1]) and torch.equal(outs[1][0], outs[1][1])  # queue of 16 fills in two steps: pass-through
    assert not torch.equal(outs[2][0], outs[2][1])                                      # full: the batch comes out of the shuffled pool
