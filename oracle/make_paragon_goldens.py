"""Generate tests/golden/paragon_goldens.npz by running the REFERENCE's own fork extras.

Build container only (needs /root/reference):

    python -m oracle.make_paragon_goldens

`traiNNer/models/paragon_otf_degradations.py` is loaded by file path (``traiNNer.models`` itself does not import
here: spandrel / ema_pytorch are missing).  For every stage of SURVEY.md §8 row f3 the script
  1. seeds the reference's generators, calls the reference function with the stage forced on,
  2. replays the draws with ``oracle.paragon_oracle.draw_extras`` / explicit replay and calls the oracle restatement,
  3. asserts oracle == reference bit for bit and stores inputs, parameters and outputs.
It then composes order (A) of ``RealESRGANModel.feed_data`` (realesrgan_model.py:512-616) from the reference's
functions for several seeds, checks that ``draw_extras`` reproduces the draw order (identical outputs) and that the
product's ``draw_plan(order="fork")`` draws the same plan, and stores one chain with per-stage taps.
"""

from __future__ import annotations

import importlib.util
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import otf_oracle as O  # noqa: E402
from oracle import paragon_oracle as P  # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "paragon_goldens.npz")


def same(a: torch.Tensor, b: torch.Tensor, what: str) -> None:
    if a.shape != b.shape or not torch.equal(a.contiguous(), b.contiguous()):
        d = (a - b).abs().max().item() if a.shape == b.shape else float("nan")
        raise SystemExit(f"oracle != reference for {what}: max-abs {d}")


def base_opt(**kw) -> SimpleNamespace:
    """The reference's option fields for this path with the schema defaults (redux_options.py:525-700, :854-894)."""
    o = SimpleNamespace(
        p_clean=0.0,  # ReduxOptions always defines it (redux_options.py), so the gate at realesrgan_model.py:487-489 always draws
        scale=4, blur_prob=0.0, lens_distort_prob=0.0, lens_distort_strength_range=(-0.3, 0.3), chromatic_aberration_prob=0.0,
        motion_blur_prob=0.0, motion_blur_kernel_size=(5, 15), motion_blur_angle_range=(0, 360), demosaic_prob=0.0,
        sensor_noise_prob=0.0, sensor_noise_std_range=(0.01, 0.1), rolling_shutter_prob=0.0, rolling_shutter_strength_range=(-0.1, 0.1),
        exposure_prob=0.0, exposure_factor_range=(0.5, 2.0), color_temp_prob=0.0, color_temp_shift_range=(-0.2, 0.2),
        oversharpen_prob=0.0, oversharpen_strength=(1.0, 2.0), aliasing_prob=0.0, aliasing_scale_range=(0.6, 0.9),
        resize_mode_list3=["bilinear", "bicubic", "nearest-exact", "lanczos"], resize_mode_prob3=[0.25, 0.25, 0.25, 0.25],
        compression_formats=["jpeg", "webp", "avif", "heif"], compression_weights=[0.60, 0.25, 0.10, 0.05],
        compression_jpeg_range=(45, 95), compression_webp_range=(45, 95), compression_avif_range=(35, 90), compression_heif_range=(40, 90),
        recompression_prob=0.0, recompression_formats=["jpeg", "webp", "avif", "heif"], recompression_weights=[0.50, 0.35, 0.10, 0.05],
        editing_prob=0.0, editing_exposure_prob=0.0, editing_exposure_range=(0.9, 1.1), editing_oversharpen_prob=0.0,
        editing_oversharpen_strength=(1.0, 1.3),
    )
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def main() -> None:
    torch.set_num_threads(1)
    R = load_reference()
    spec = importlib.util.spec_from_file_location("paragon_otf_degradations", os.path.join(REFERENCE_ROOT, "traiNNer", "models", "paragon_otf_degradations.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Ref = mod.ParagonOTF
    G: dict[str, np.ndarray] = {}

    def put(name: str, t) -> None:
        if isinstance(t, torch.Tensor):
            t = t.detach().contiguous().numpy()
        G[name] = np.asarray(t)

    def seed(s: int):
        R.RNG._rng = np.random.default_rng(s)
        random.seed(s + 1000)
        torch.manual_seed(s + 2000)
        return np.random.default_rng(s), random.Random(s + 1000)

    imgs = {"nat": O.synth_gt(2, 40, 36, "natural", seed=21), "uni": O.synth_gt(2, 37, 52, "uniform", seed=22),
            "sq": O.synth_gt(1, 64, 64, "uniform", seed=23)}
    for k, v in imgs.items():
        put(f"img_{k}", v)

    # ---- per-stage cases: (name, option overrides, oracle call given the replayed draws) ----
    n_cases = 0
    for key, img in imgs.items():
        for s in range(3):
            # lens distortion
            opt = base_opt(lens_distort_prob=1.0)
            nr, _ = seed(100 + s)
            ref = Ref.apply_lens_distortion(img, opt)
            nr.uniform()
            st = nr.uniform(*opt.lens_distort_strength_range)
            same(P.lens_distortion(img, st), ref, "lens")
            put(f"lens_{key}_{s}_p", np.float64(st)); put(f"lens_{key}_{s}", ref)
            put(f"lens_{key}_{s}_ieee", P.lens_distortion(img, st, sqrt=P.ieee_sqrt))
            # rolling shutter
            opt = base_opt(rolling_shutter_prob=1.0, rolling_shutter_strength_range=(-0.3, 0.3))
            nr, _ = seed(200 + s)
            ref = Ref.apply_rolling_shutter(img, opt)
            nr.uniform()
            st = nr.uniform(*opt.rolling_shutter_strength_range)
            same(P.rolling_shutter(img, st), ref, "shutter")
            put(f"shutter_{key}_{s}_p", np.float64(st)); put(f"shutter_{key}_{s}", ref)
            # motion blur (sizes include even K: the reference's output grows by one row/column)
            opt = base_opt(motion_blur_prob=1.0, motion_blur_kernel_size=(3, 21))
            nr, pr = seed(300 + s)
            ref = Ref.apply_motion_blur(img, opt)
            nr.uniform()
            ks = pr.randint(3, 21)
            ang = nr.uniform(0, 360)
            same(P.motion_blur(img, ks, ang), ref, "motion")
            put(f"motion_{key}_{s}_p", np.array([ks, ang], dtype=np.float64)); put(f"motion_{key}_{s}", ref)
            put(f"motion_{key}_{s}_k", P.motion_blur_kernel(ks, ang))
            # exposure / colour temperature / oversharpen / aliasing / sensor noise
            opt = base_opt(exposure_prob=1.0)
            nr, _ = seed(400 + s)
            ref = Ref.apply_exposure_errors(img, opt)
            nr.uniform()
            f = nr.uniform(*opt.exposure_factor_range)
            same(P.exposure(img, f), ref, "exposure")
            put(f"exposure_{key}_{s}_p", np.float64(f)); put(f"exposure_{key}_{s}", ref)
            opt = base_opt(color_temp_prob=1.0)
            nr, _ = seed(500 + s)
            ref = Ref.apply_color_temperature_shift(img, opt)
            nr.uniform()
            sh = nr.uniform(*opt.color_temp_shift_range)
            same(P.color_temperature(img, sh), ref, "color_temp")
            put(f"ctemp_{key}_{s}_p", np.float64(sh)); put(f"ctemp_{key}_{s}", ref)
            opt = base_opt(oversharpen_prob=1.0)
            nr, _ = seed(600 + s)
            ref = Ref.apply_oversharpening(img, opt)
            nr.uniform()
            st = nr.uniform(*opt.oversharpen_strength)
            same(P.oversharpen(img, st), ref, "oversharpen")
            put(f"oversharp_{key}_{s}_p", np.float64(st)); put(f"oversharp_{key}_{s}", ref)
            opt = base_opt(aliasing_prob=1.0, aliasing_scale_range=(0.3, 0.95))
            nr, _ = seed(700 + s)
            ref = Ref.apply_aliasing_artifacts(img, opt)
            nr.uniform()
            sc = nr.uniform(*opt.aliasing_scale_range)
            same(P.aliasing(img, sc), ref, "aliasing")
            put(f"alias_{key}_{s}_p", np.float64(sc)); put(f"alias_{key}_{s}", ref)
            opt = base_opt(sensor_noise_prob=1.0)
            nr, _ = seed(800 + s)
            ref = Ref.apply_sensor_noise(img, opt)
            nr.uniform()
            sd = nr.uniform(*opt.sensor_noise_std_range)
            torch.manual_seed(800 + s + 2000)
            nz = torch.randn_like(img)
            same(P.sensor_noise(img, sd, nz), ref, "sensor noise")
            put(f"sensor_{key}_{s}_p", np.float64(sd)); put(f"sensor_{key}_{s}_noise", nz); put(f"sensor_{key}_{s}", ref)
            n_cases += 9
        # Bayer demosaic (cv2 on the host in the reference): reference == cv2 restatement == numpy integer restatement
        opt = base_opt(demosaic_prob=1.0)
        seed(850)
        ref = Ref.apply_demosaicing_artifacts(img, opt)
        same(P.demosaic_cv2(img), ref, "demosaic (cv2)")
        same(P.demosaic(img), ref, "demosaic (numpy restatement of OpenCV's bilinear Bayer interpolation)")
        put(f"demosaic_{key}", ref)
        n_cases += 1
        # chromatic aberration (no parameters) and aliasing at the identity / halving shortcuts of legacy nearest
        opt = base_opt(chromatic_aberration_prob=1.0)
        seed(900)
        ref = Ref.apply_chromatic_aberration(img, opt)
        same(P.chromatic_aberration(img), ref, "chroma")
        put(f"chroma_{key}", ref)
        put(f"alias_{key}_half", P.aliasing(img, 0.5))
        n_cases += 2
    for hh, ww in ((3, 3), (4, 7), (9, 4), (2, 6), (31, 33)):  # odd sizes, the 3x3 minimum, and the degenerate < 3 case (zeros)
        small = O.synth_gt(1, hh, ww, "uniform", seed=hh * 10 + ww)
        ref = Ref.apply_demosaicing_artifacts(small, base_opt(demosaic_prob=1.0)) if seed(851) else None
        same(P.demosaic(small), ref, f"demosaic {hh}x{ww}")
        put(f"demosaic_small_{hh}x{ww}_in", small); put(f"demosaic_small_{hh}x{ww}", ref)
    # the PIL JPEG round the product replaces with DiffJPEG: oracle restatement == reference
    opt = base_opt()
    nr, _ = seed(950)
    ref = Ref._compress_with_format(imgs["nat"], "jpeg", opt, round=1)
    q = nr.uniform(*opt.compression_jpeg_range)
    same(P.pil_jpeg(imgs["nat"], q), ref, "PIL jpeg")

    # ---- order (A): realesrgan_model.py:512-616 composed from the reference's own functions ----
    from trainner_redux_b200.realesrgan_feed import HostRNG, OTFOptions, draw_plan  # product-side host draws (checked here)

    def reference_chain(gt, k1, sk, opt):
        if hasattr(opt, "p_clean") and R.RNG.get_rng().uniform() < opt.p_clean:   # :487-489 (drawn even at probability 0)
            return torch.clamp((gt.clone() * 255.0).round(), 0, 255) / 255.0      # :491-493
        out = Ref.apply_lens_distortion(gt, opt)                       # :516
        out = Ref.apply_chromatic_aberration(out, opt)                 # :519 (the model's copy, :244-310, is the same code)
        out = Ref.apply_motion_blur(out, opt)                          # :522
        if R.RNG.get_rng().uniform() < opt.blur_prob:                  # :525
            out = R.ipu.filter2d(out, k1)
        out = Ref.apply_demosaicing_artifacts(out, opt)                # :534 (the model's copy, :312-363, is the same code)
        out = Ref.apply_sensor_noise(out, opt)                         # :537
        out = Ref.apply_rolling_shutter(out, opt)                      # :540
        out = Ref.apply_exposure_errors(out, opt)                      # :548
        out = Ref.apply_color_temperature_shift(out, opt)              # :551
        out = Ref.apply_oversharpening(out, opt)                       # :554
        out = Ref.apply_aliasing_artifacts(out, opt)                   # :557
        mode = random.choices(opt.resize_mode_list3, weights=opt.resize_mode_prob3)[0]  # :564
        h, w = gt.shape[2:4]
        out = R.deg.resize_pt(out, size=(h // opt.scale, w // opt.scale), mode=mode)   # :567
        out = R.ipu.filter2d(out, sk)                                  # :574
        out = Ref.apply_realistic_compression_pipeline(out, opt)       # :581
        if R.RNG.get_rng().uniform() < opt.editing_prob:               # :590
            if R.RNG.get_rng().uniform() < opt.editing_exposure_prob:  # :594
                f = R.RNG.get_rng().uniform(*opt.editing_exposure_range)
                out = torch.clamp(out * f, 0, 1)
            R.RNG.get_rng().uniform()                                  # :606 (gate drawn, nothing applied)
        return torch.clamp((out * 255.0).round(), 0, 255) / 255.0     # :616

    gt = O.synth_gt(2, 64, 48, "natural", seed=31)
    k1 = O.synth_blur_kernels(2, seed=5)
    sk = O.synth_sinc_or_pulse(2, seed=6)
    put("chain_gt", gt); put("chain_k1", k1); put("chain_sk", sk)
    probs = dict(blur_prob=0.6, lens_distort_prob=0.6, chromatic_aberration_prob=0.6, motion_blur_prob=0.6, sensor_noise_prob=0.6,
                 rolling_shutter_prob=0.6, demosaic_prob=0.5, exposure_prob=0.6, color_temp_prob=0.6, oversharpen_prob=0.6, aliasing_prob=0.6,
                 recompression_prob=0.5, editing_prob=0.6, editing_exposure_prob=0.6, editing_oversharpen_prob=0.5,
                 motion_blur_kernel_size=(5, 15), compression_formats=["jpeg", "avif"], compression_weights=[0.8, 0.2],
                 recompression_formats=["jpeg", "heif"], recompression_weights=[0.7, 0.3],
                 resize_mode_list3=["bilinear", "bicubic", "area", "nearest-exact", "lanczos"], resize_mode_prob3=[0.2] * 5)
    stored = None
    for s in range(12):
        opt = base_opt(**probs)
        nr, pr = seed(1000 + s)
        ref = reference_chain(gt, k1, sk, opt)
        assert not (nr.uniform() < opt.p_clean)  # the oracle's replay of the p_clean gate (:489)
        plan = P.draw_extras(opt, nr, pr)
        plan["scale"] = opt.scale
        torch.manual_seed(1000 + s + 2000)
        grow = 1 if ("motion" in plan and plan["motion"][0] % 2 == 0) else 0  # an even motion kernel grows the image (conv2d padding K//2)
        inject = {"sensor_noise": torch.randn(gt.size(0), 3, gt.size(2) + grow, gt.size(3) + grow)} if "sensor" in plan else {}
        taps: dict = {}
        mine = P.apply_extras_a(gt, k1, sk, plan, inject, taps)
        same(mine, ref, f"order (A) chain, seed {s}")
        # the product's host draws: same generators, same plan
        popt = OTFOptions(order="fork", gt_size=32, **{k: v for k, v in probs.items()})
        rng = HostRNG(0)
        rng.np, rng.py = np.random.default_rng(1000 + s), random.Random(1000 + s + 1000)
        pplan = draw_plan(popt, 2, 64, 48, rng)
        for k, v in plan.items():
            assert pplan.get(k) == v, f"product draw_plan differs at {k!r}: {pplan.get(k)!r} vs {v!r} (seed {s})"
        n_on = sum(k in plan for k in ("lens", "chroma", "motion", "demosaic", "sensor", "shutter", "exposure", "color_temp", "oversharpen", "aliasing"))
        if stored is None or n_on > stored[0]:
            stored = (n_on, s, plan, inject, taps)
    n_on, s, plan, inject, taps = stored
    put("chain_seed", np.int64(1000 + s))
    import json

    put("chain_plan_json", np.frombuffer(json.dumps(plan).encode(), dtype=np.uint8))
    if inject:
        put("chain_sensor_noise", inject["sensor_noise"])
    for k, v in taps.items():
        put(f"chain_tap_{k}", v)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {len(G)} arrays ({n_cases} stage cases, chain with {n_on} extras on), {os.path.getsize(OUT)/1e6:.2f} MB; "
          "oracle == reference on every case")


if __name__ == "__main__":
    main()
