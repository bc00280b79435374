"""Generate tests/golden/otf_goldens.npz by running the REFERENCE's own Python.

Run in the build container only (needs /root/reference):

    python -m oracle.make_goldens

For every hot-path primitive this script
  1. runs the imported reference function on small seeded inputs,
  2. runs the oracle restatement (oracle/otf_oracle.py) on the same inputs and
     asserts the two agree bit for bit (same ATen build => same bits),
  3. stores inputs, injected random fields and reference outputs.
The frozen vectors then travel to the GPU box, where the reference does not
exist, and pin both the oracle (tests/test_oracle_cpu.py) and the CUDA kernels
(tests/test_parity_gpu.py).
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import otf_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "otf_goldens.npz")


def same(a: torch.Tensor, b: torch.Tensor, what: str) -> None:
    a = a.contiguous()
    b = b.contiguous()
    if a.shape != b.shape or not torch.equal(a, b):
        d = (a - b).abs().max().item() if a.shape == b.shape else float("nan")
        raise SystemExit(f"oracle != reference for {what}: max-abs {d}")


def main() -> None:
    torch.set_num_threads(1)  # fixed reduction order for the frozen vectors
    R = load_reference()
    G: dict[str, np.ndarray] = {}

    def put(name: str, t) -> None:
        if isinstance(t, torch.Tensor):
            t = t.detach().contiguous().numpy()
        G[name] = np.asarray(t)

    img = O.synth_gt(2, 40, 36, "natural", seed=11)
    img_u = O.synth_gt(2, 40, 36, "uniform", seed=12)
    put("img", img)
    put("img_u", img_u)

    # ---- a1 filter2d: per-sample 21x21 (true sizes 7..21), shared 5x5 ----
    k21 = O.synth_blur_kernels(2, seed=3)
    put("f2d_k21", k21)
    ref = R.ipu.filter2d(img, k21)
    same(O.filter2d(img, k21), ref, "filter2d per-sample")
    put("f2d_out21", ref)
    k5 = torch.rand(1, 5, 5, generator=torch.Generator().manual_seed(5))
    k5 = k5 / k5.sum()
    put("f2d_k5", k5)
    ref = R.ipu.filter2d(img_u, k5)
    same(O.filter2d(img_u, k5), ref, "filter2d shared")
    put("f2d_out5", ref)
    sinc = O.synth_sinc_or_pulse(2, seed=4, sinc_prob=1.0)
    put("f2d_sinc", sinc)
    ref = R.ipu.filter2d(img_u, sinc)
    same(O.filter2d(img_u, sinc), ref, "filter2d sinc")
    put("f2d_outsinc", ref)

    # ---- a2 USM: radius 50 (51 taps, sigma 8) and radius 7 (cv2 table) ----
    img_usm = O.synth_gt(2, 64, 60, "natural", seed=13)
    put("usm_img", img_usm)
    for radius in (50, 7):
        mod = R.ipu.USMSharp(radius=radius)
        same(O.usm_kernel(radius), mod.kernel, f"usm kernel r={radius}")
        ref = mod(img_usm)
        same(O.usm_sharp(img_usm, O.usm_kernel(radius)), ref, f"usm r={radius}")
        put(f"usm_out_r{radius}", ref)
    mod = R.ipu.USMSharp(radius=50)
    ref = mod(img_usm, weight=0.8, threshold=4)
    same(O.usm_sharp(img_usm, O.usm_kernel(50), 0.8, 4), ref, "usm w/t")
    put("usm_out_w08_t4", ref)

    # ---- a3 resize_pt: 5 modes x {scale factors, explicit size} ----
    for mode in O.RESIZE_MODES:
        for s in (0.4, 0.75, 1.25, 1.5):
            ref = R.deg.resize_pt(img_u, mode, scale_factor=s)
            same(O.resize_pt(img_u, mode, scale_factor=s), ref, f"resize {mode} s={s}")
            put(f"rs_{mode}_s{s}", ref)
        for size in ((10, 9), (40, 36), (17, 50)):
            ref = R.deg.resize_pt(img_u, mode, size=size)
            same(O.resize_pt(img_u, mode, size=size), ref, f"resize {mode} size={size}")
            put(f"rs_{mode}_{size[0]}x{size[1]}", ref)

    # ---- a4 Gaussian noise: replay the reference's torch draws ----
    for tag, gray_prob, seed in (("mixed", 0.5, 21), ("color", 0.0, 22), ("allgray", 1.0, 23)):
        torch.manual_seed(seed)
        ref = R.deg.random_add_gaussian_noise_pt(img, sigma_range=(1, 30), gray_prob=gray_prob, clip=True, rounds=False)
        torch.manual_seed(seed)
        sigma = torch.rand(2) * (30 - 1) + 1
        gray = (torch.rand(2) < gray_prob).float()
        n_gray = torch.randn(40, 36) if gray.sum() > 0 else None
        n_col = torch.randn(2, 3, 40, 36)
        same(O.add_gaussian_noise(img, sigma, gray, n_col, n_gray), ref, f"gaussian {tag}")
        put(f"gn_{tag}_sigma", sigma)
        put(f"gn_{tag}_gray", gray)
        put(f"gn_{tag}_ncol", n_col)
        if n_gray is not None:
            put(f"gn_{tag}_ngray", n_gray)
        put(f"gn_{tag}_out", ref)
    torch.manual_seed(24)
    ref = R.deg.add_gaussian_noise_pt(img, sigma=12.5, gray_noise=0, clip=True, rounds=True)
    torch.manual_seed(24)
    n_col = torch.randn(2, 3, 40, 36)
    same(O.add_gaussian_noise(img, torch.full((2,), 12.5), torch.zeros(2), n_col, None, True, True), ref, "gaussian rounds")
    put("gn_rounds_ncol", n_col)
    put("gn_rounds_out", ref)

    # ---- a5 Poisson noise: same generator state => same torch.poisson counts ----
    for tag, src, gray_prob, seed in (
        ("mixed", img, 0.5, 31),
        ("color", img_u, 0.0, 32),
        ("allgray", img, 1.0, 33),
        ("twolevel", O.synth_gt(2, 40, 36, "twolevel", seed=14), 0.5, 34),
        ("flat", O.synth_gt(2, 40, 36, "flat"), 0.5, 35),
    ):
        torch.manual_seed(seed)
        ref = R.deg.random_add_poisson_noise_pt(src, scale_range=(0.05, 3), gray_prob=gray_prob, clip=True, rounds=False)
        torch.manual_seed(seed)
        scale = torch.rand(2) * (3 - 0.05) + 0.05
        gray = (torch.rand(2) < gray_prob).float()
        qc, vc, lc, qg, vg, lg = O.poisson_lambda(src)
        cg = torch.poisson(lg) if gray.sum() > 0 else None
        cc = torch.poisson(lc)
        got = O.add_poisson_noise(src, scale, gray, counts_color=cc, counts_gray=cg)
        same(got, ref, f"poisson {tag}")
        put(f"pn_{tag}_img", src)
        put(f"pn_{tag}_scale", scale)
        put(f"pn_{tag}_gray", gray)
        put(f"pn_{tag}_vals_color", vc.view(-1))
        put(f"pn_{tag}_vals_gray", vg.view(-1))
        put(f"pn_{tag}_counts_color", cc)
        if cg is not None:
            put(f"pn_{tag}_counts_gray", cg)
        put(f"pn_{tag}_out", ref)

    # ---- a6 DiffJPEG ----
    jimg = O.synth_gt(2, 40, 36, "natural", seed=15)  # pads to 48x48
    jimg2 = O.synth_gt(2, 48, 32, "uniform", seed=16)
    put("jpg_img", jimg)
    put("jpg_img2", jimg2)
    for diff in (False, True):
        mod = R.dj.DiffJPEG(differentiable=diff)
        for tag, src, q in (("t", jimg, torch.tensor([30.0, 80.0])), ("u", jimg2, torch.tensor([55.0, 95.0]))):
            ref = mod(src, quality=q.clone())
            same(O.diffjpeg(src, q.clone(), diff), ref, f"diffjpeg {tag} diff={diff}")
            put(f"jpg_{tag}_q", q)
            put(f"jpg_{tag}_out_d{int(diff)}", ref)
        ref = mod(jimg, quality=50)
        same(O.diffjpeg(jimg, 50, diff), ref, "diffjpeg scalar q")
        put(f"jpg_s50_out_d{int(diff)}", ref)
    # quirk Q1: the quality tensor is overwritten with factors
    q = torch.tensor([30.0, 80.0])
    R.dj.DiffJPEG(differentiable=False)(jimg, quality=q)
    put("jpg_t_factor", q)

    # ---- a7/a8 clamp-round and paired crop ----
    x = img_u * 1.2 - 0.1
    ref = torch.clamp((x * 255.0).round(), 0, 255) / 255.0  # realesrgan_model.py:616
    same(O.clamp_round(x), ref, "clamp_round")
    put("cr_in", x)
    put("cr_out", ref)

    # ---- chain, order (B): composed here from REFERENCE primitives ----
    gt = O.synth_gt(2, 64, 64, "natural", seed=17)
    k1 = O.synth_blur_kernels(2, seed=6)
    k2 = O.synth_blur_kernels(2, seed=7)
    sk = O.synth_sinc_or_pulse(2, seed=8, sinc_prob=1.0)
    jpeger = R.dj.DiffJPEG(differentiable=False)
    g = torch.Generator().manual_seed(41)
    sigma1 = torch.rand(2, generator=g) * 29 + 1
    gray1 = torch.tensor([1.0, 0.0])
    scale2 = torch.rand(2, generator=g) * 2.95 + 0.05
    gray2 = torch.tensor([0.0, 1.0])
    q1 = torch.tensor([35.0, 90.0])
    q2 = torch.tensor([60.0, 45.0])
    out = R.ipu.filter2d(gt, k1)
    out = R.deg.resize_pt(out, "bicubic", scale_factor=0.75)
    h1, w1 = out.shape[2:]
    n1c = torch.randn(2, 3, h1, w1, generator=g)
    n1g = torch.randn(h1, w1, generator=g)
    out = out + O.gaussian_noise_field(out, sigma1, gray1, n1c, n1g)  # explicit-noise form of degradations.py:569-605
    out = torch.clamp(out, 0, 1)
    out = jpeger(torch.clamp(out, 0, 1), quality=q1.clone()).contiguous()
    out = R.ipu.filter2d(out, k2)
    out = R.deg.resize_pt(out, "bilinear", size=(int(64 / 4 * 1.1), int(64 / 4 * 1.1)))
    # Poisson stage through the reference with torch's generator replayed
    torch.manual_seed(42)
    qc, vc, lc, qg, vg, lg = O.poisson_lambda(out)
    cg2 = torch.poisson(lg)
    cc2 = torch.poisson(lc)
    torch.manual_seed(42)
    out_ref = R.deg.add_poisson_noise_pt(out, scale=scale2, gray_noise=gray2, clip=True, rounds=False)
    same(O.add_poisson_noise(out, scale2, gray2, counts_color=cc2, counts_gray=cg2), out_ref, "chain poisson")
    out = out_ref
    out = R.deg.resize_pt(out, "area", size=(16, 16))
    out = R.ipu.filter2d(out, sk)
    out = jpeger(torch.clamp(out, 0, 1), quality=q2.clone()).contiguous()
    lq = torch.clamp((out * 255.0).round(), 0, 255) / 255.0
    import random

    random.seed(9)
    gt_c, lq_c = R.tfm.paired_random_crop(gt, lq, 48, 4)
    random.seed(9)
    top, left = random.randint(0, 16 - 12), random.randint(0, 16 - 12)
    plan = {
        "scale": 4,
        "gt_size": 48,
        "blur1": True,
        "resize1": {"scale": 0.75, "mode": "bicubic"},
        "noise1": {"kind": "gaussian", "sigma": sigma1, "gray": gray1},
        "jpeg1": q1,
        "blur2": True,
        "resize2": {"scale": 1.1, "mode": "bilinear"},
        "noise2": {"kind": "poisson", "scale": scale2, "gray": gray2},
        "final_order": "resize_first",
        "resize3_mode": "area",
        "jpeg2": q2,
        "crop": (top, left),
    }
    noise = {"noise1_color": n1c, "noise1_gray": n1g, "noise2_counts_color": cc2, "noise2_counts_gray": cg2}
    o_gt, o_lq = O.run_chain_b(gt, k1, k2, sk, plan, noise)
    same(o_lq, lq_c, "chain B lq")
    same(o_gt, gt_c, "chain B gt")
    for name, t in (
        ("gt", gt), ("k1", k1), ("k2", k2), ("sinc", sk), ("sigma1", sigma1), ("gray1", gray1), ("scale2", scale2),
        ("gray2", gray2), ("q1", q1), ("q2", q2), ("n1c", n1c), ("n1g", n1g), ("cc2", cc2), ("cg2", cg2),
        ("crop", torch.tensor([top, left])), ("lq_full", lq), ("lq", lq_c), ("gt_crop", gt_c),
    ):
        put(f"chain_{name}", t)


    # ---- f2 kernel synthesis: reference generators with explicit parameters + a seeded dataset-style draw ----
    from oracle import kernel_synth_oracle as KS

    prm = []
    g2 = np.random.default_rng(51)
    for kind in range(6):
        for k in (5, 7, 13, 21):
            prm.append([kind, k, g2.uniform(0.2, 3), g2.uniform(0.2, 3), g2.uniform(-np.pi, np.pi), g2.uniform(0.5, 4), 0.0, 21])
    for k in (5, 7, 11, 13, 21):
        prm.append([6, k, 0, 0, 0, 0, g2.uniform(np.pi / 5, np.pi), 21])
    prm.append([7, 21, 0, 0, 0, 0, 0, 21])
    prm = np.asarray(prm, np.float64)
    ref_k = np.zeros((len(prm), 21, 21), np.float32)
    for i, (kind, k, sx, sy, th, beta, wc, _pad) in enumerate(prm):
        kind, k = int(kind), int(k)
        iso = kind in (0, 2, 4)
        if kind == 7:
            ref_k[i, 10, 10] = 1
            continue
        if kind == 6:
            ker = R.deg.circular_lowpass_kernel(wc, k, pad_to=False)
        else:
            if kind < 2:
                ker = R.deg.bivariate_gaussian(k, sx, sy, th, isotropic=iso)
            elif kind < 4:
                ker = R.deg.bivariate_generalized_gaussian(k, sx, sy, th, beta, isotropic=iso)
            else:
                ker = R.deg.bivariate_plateau(k, sx, sy, th, beta, isotropic=iso)
            ker = ker / np.sum(ker)  # the random_* wrappers normalise once more (degradations.py:258/313/369)
        pd = (21 - k) // 2
        ref_k[i] = np.pad(ker, ((pd, pd), (pd, pd))).astype(np.float32)
    assert np.array_equal(KS.synthesize(prm), ref_k), "kernel_synth oracle != reference"
    put("ks_params", prm)
    put("ks_ref", ref_k)
    # dataset-style draws (realesrgan_dataset.py:149-206) through the reference's own random_* functions
    import math as _math
    import random as _random

    from trainner_redux_b200.kernels import KernelOptions, draw_kernel_params

    kopt = KernelOptions(sinc_prob=0.1, sinc_prob2=0.1, final_sinc_prob=0.8, kernel_range=(7, 21), kernel_range2=(7, 21),
                         final_kernel_range=(7, 21))
    R.RNG._rng = np.random.default_rng(77)
    _random.seed(78)
    ds_k1, ds_k2, ds_sk = [], [], []

    def ref_blur(sizes, sinc_prob, klist, kprob, sigma, betag, betap):
        ksz = _random.choice(sizes)
        if R.RNG.get_rng().uniform() < sinc_prob:
            wc = R.RNG.get_rng().uniform(np.pi / 3, np.pi) if ksz < 13 else R.RNG.get_rng().uniform(np.pi / 5, np.pi)
            ker = R.deg.circular_lowpass_kernel(wc, ksz, pad_to=False)
        else:
            ker = R.deg.random_mixed_kernels(klist, kprob, ksz, sigma, sigma, (-_math.pi, _math.pi), betag, betap, noise_range=None)
        pd = (21 - ksz) // 2
        return np.pad(ker, ((pd, pd), (pd, pd)))

    for _ in range(12):
        ds_k1.append(ref_blur(list(range(7, 22, 2)), kopt.sinc_prob, kopt.kernel_list, kopt.kernel_prob, kopt.blur_sigma, kopt.betag_range, kopt.betap_range))
        ds_k2.append(ref_blur(list(range(7, 22, 2)), kopt.sinc_prob2, kopt.kernel_list2, kopt.kernel_prob2, kopt.blur_sigma2, kopt.betag_range2, kopt.betap_range2))
        if R.RNG.get_rng().uniform() < kopt.final_sinc_prob:
            ksz = _random.choice(list(range(7, 22, 2)))
            wc = R.RNG.get_rng().uniform(np.pi / 3, np.pi)
            ds_sk.append(R.deg.circular_lowpass_kernel(wc, ksz, pad_to=21))
        else:
            pulse = np.zeros((21, 21))
            pulse[10, 10] = 1
            ds_sk.append(pulse)
    ds = [np.stack(x).astype(np.float32) for x in (ds_k1, ds_k2, ds_sk)]
    mine = draw_kernel_params(kopt, 12, _random.Random(78), np.random.default_rng(77))
    for name, want, prm_i in zip(("k1", "k2", "sinc"), ds, mine):
        assert np.array_equal(KS.synthesize(prm_i), want), f"dataset-order draw mismatch for {name}"
        put(f"ks_ds_{name}", want)

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT)/1e6:.2f} MB; oracle == reference on every case")


if __name__ == "__main__":
    main()
