"""The native stage executor (otf_run_stages_f32, stages.py) against the per-stage Python path: the same
plan, the same generators -> bit-identical pairs, for both stage orders and every stage kind (USM, Poisson,
lanczos prefilters, both final orders, injected fields); plus the executor's own argument checks."""

from __future__ import annotations

import ctypes as C

import pytest
import torch

from oracle import otf_oracle as O
from trainner_redux_b200 import _lib
from trainner_redux_b200 import degradations as D
from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed, clamp_round
from trainner_redux_b200.stages import StageList

pytestmark = pytest.mark.gpu


def _opts(order: str, seed: int) -> OTFOptions:
    return OTFOptions(
        scale=(4, 1, 2, 2)[seed % 4], gt_size=64, order=order, lq_usm=(seed % 3 == 0), lq_usm_radius_range=(3, 21),
        blur_prob=0.8, blur_prob2=0.7, gaussian_noise_prob=0.5, gaussian_noise_prob2=0.5, noise_range=(1, 30), noise_range2=(1, 25),
        poisson_scale_range=(0.05, 3), poisson_scale_range2=(0.05, 2.5), gray_noise_prob=0.4, gray_noise_prob2=0.4,
        jpeg_prob=0.8, jpeg_range=(30, 95), jpeg_prob2=0.8, jpeg_range2=(30, 95), resize_range=(0.3, 1.5), resize_range2=(0.4, 1.2),
        resize_mode_list=("bilinear", "bicubic", "nearest-exact", "lanczos", "area"), resize_mode_prob=(0.2,) * 5,
        resize_mode_list2=("bilinear", "bicubic", "nearest-exact", "lanczos", "area"), resize_mode_prob2=(0.2,) * 5,
        queue_size=12, p_clean=0.3 if seed % 4 == 1 else 0)  # the clean pass-through only works at scale 1


def _data(b: int, size: int, seed: int) -> dict:
    return {"gt": O.synth_gt(b, size, size, "natural", seed=seed), "kernel1": O.synth_blur_kernels(b, seed=seed + 1),
            "kernel2": O.synth_blur_kernels(b, seed=seed + 2), "sinc_kernel": O.synth_sinc_or_pulse(b, seed=seed + 3)}


@pytest.mark.parametrize("order", ["classic", "fork"])
@pytest.mark.parametrize("seed", list(range(8)))
def test_native_chain_is_bit_identical_to_the_per_stage_path(dev, order, seed):
    b, size = 4, 96 + 4 * seed
    eager = RealESRGANFeed(_opts(order, seed), device=dev, manual_seed=seed, use_pool=True)
    native = RealESRGANFeed(_opts(order, seed), device=dev, manual_seed=seed, use_pool=True)
    eager.native_chain = False
    assert native.native_chain
    for it in range(5):  # several iterations: the pool fills (12 / 4) and starts to shuffle
        d = _data(b, size, 10 * seed + it)
        D._TABLE_CACHE.clear()  # both paths start from the same resize-table cache state
        l0 = _lib.launch_count
        eager.feed_data(dict(d))
        l1 = _lib.launch_count
        D._TABLE_CACHE.clear()
        native.feed_data(dict(d))
        l2 = _lib.launch_count
        assert eager.last_plan.keys() == native.last_plan.keys()
        assert torch.equal(eager.lq, native.lq), (order, seed, it, eager.last_plan)
        assert torch.equal(eager.gt, native.gt)
        # the executor fuses resize + Gaussian noise, the last JPEG / clamp-round + crop, and skips identity resizes
        assert l2 - l1 <= l1 - l0, "the native chain must never launch more kernels than the per-stage path"


@pytest.mark.parametrize("seed", list(range(6)))
def test_native_fork_chain_with_extras_is_bit_identical_to_the_per_stage_path(dev, seed):
    """Order (A) with every extra stage of the fork likely (lens / chroma / motion incl. even K / demosaic / sensor noise /
    shutter / exposure / colour temperature / oversharpen / aliasing / JPEG rounds / editing exposure): one library call
    per chain (stage ops OTF_OP_WARP .. OTF_OP_TRUNC8) against one Python call per stage."""
    opt = OTFOptions(order="fork", scale=(4, 2)[seed % 2], gt_size=64, blur_prob=0.7, lens_distort_prob=0.6, chromatic_aberration_prob=0.5,
                     motion_blur_prob=0.6, sensor_noise_prob=0.6, rolling_shutter_prob=0.5, exposure_prob=0.5, color_temp_prob=0.5,
                     oversharpen_prob=0.5, aliasing_prob=0.5, demosaic_prob=0.4, recompression_prob=0.5, editing_prob=0.5,
                     editing_exposure_prob=0.5, compression_formats=("jpeg", "webp"), compression_weights=(0.8, 0.2),
                     recompression_formats=("jpeg",), recompression_weights=(1.0,), motion_blur_kernel_size=(4, 15), queue_size=12)
    eager = RealESRGANFeed(opt, device=dev, manual_seed=seed, use_pool=True)
    native = RealESRGANFeed(opt, device=dev, manual_seed=seed, use_pool=True)
    eager.native_chain = False
    seen: set = set()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # (the webp rounds pass through with a one-time warning)
        for it in range(6):
            d = _data(4, 96 + 4 * seed, 10 * seed + it)
            D._TABLE_CACHE.clear()
            l0 = _lib.launch_count
            eager.feed_data(dict(d))
            l1 = _lib.launch_count
            D._TABLE_CACHE.clear()
            native.feed_data(dict(d))
            l2 = _lib.launch_count
            assert eager.last_plan.keys() == native.last_plan.keys()
            seen.update(eager.last_plan.keys())
            assert torch.equal(eager.lq, native.lq), (seed, it, {k: v for k, v in eager.last_plan.items() if not torch.is_tensor(v)})
            assert torch.equal(eager.gt, native.gt)
            assert l2 - l1 <= l1 - l0
    assert len(seen & {"lens", "chroma", "motion", "demosaic", "sensor", "shutter", "exposure", "color_temp", "oversharpen", "aliasing"}) >= 6


@pytest.mark.parametrize("noise", ["gaussian", "poisson"])
@pytest.mark.parametrize("final_order", ["resize_first", "jpeg_first"])
def test_fused_launches_are_bit_identical_to_the_unfused_executor(dev, monkeypatch, noise, final_order):
    """Row g1: the executor's fused launches (resize + Gaussian noise, DiffJPEG + lattice + both crops, clamp/round +
    both crops) against the same executor with OTF_FUSE=0 — same plan, same Philox positions -> identical bits, fewer
    launches.  Odd extents on purpose (partial quads at row ends, ragged tiles)."""
    b, size = 3, 132
    d = {k: v.to(dev) for k, v in _data(b, size, 5).items()}
    g = torch.Generator().manual_seed(3)
    key = "sigma" if noise == "gaussian" else "scale"
    plan = {"scale": 4, "gt_size": 96, "order": "classic", "blur1": True, "resize1": {"scale": 0.83, "mode": "bicubic"},
            "noise1": {"kind": noise, key: torch.rand(b, generator=g) * 20 + 1, "gray": torch.tensor([0.0, 1.0, 0.0])},
            "jpeg1": torch.tensor([40.0, 90.0, 60.0]), "blur2": True, "resize2": {"scale": 1.0, "mode": "bilinear"},
            "noise2": {"kind": noise, key: torch.rand(b, generator=g) * 10 + 1, "gray": torch.tensor([1.0, 0.0, 0.0])},
            "final_order": final_order, "resize3_mode": "area", "jpeg2": torch.tensor([55.0, 75.0, 35.0]), "crop": (3, 6)}
    outs, launches = [], []
    for fuse in ("0", "1"):
        monkeypatch.setenv("OTF_FUSE", fuse)
        feed = RealESRGANFeed(OTFOptions(scale=4, gt_size=96), device=dev, manual_seed=1, use_pool=False)
        feed.use_graphs = False
        l0 = _lib.launch_count
        feed.feed_data(dict(d), plan=plan)
        launches.append(_lib.launch_count - l0)
        outs.append((feed.gt.clone(), feed.lq.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # gaussian: two resize+noise pairs and the tail fuse (3 fewer); poisson: only the tail (1 fewer)
    assert launches[0] - launches[1] == (3 if noise == "gaussian" else 1), launches


@pytest.mark.parametrize("mode", ["bilinear", "bicubic", "area", "nearest-exact", "lanczos"])
@pytest.mark.parametrize("shape", [(3, 3, 96, 96, 72, 72), (2, 3, 70, 90, 101, 67), (2, 3, 64, 64, 21, 23), (1, 3, 300, 40, 7, 40)])
def test_resize_gauss_entry_point_matches_two_launches(dev, mode, shape):
    """otf_resize_gauss_f32 against otf_resize_f32 + otf_gaussian_noise_f32, through the raw C ABI: every mode, up and
    down, ragged widths (OW % 4 != 0), gray and colour samples, the device-side offset word."""
    b, c, h, w, oh, ow = shape
    lib = _lib.load()
    x = torch.rand(b, c, h, w, device=dev)
    mode_id = _lib.RESIZE_LANCZOS if mode == "lanczos" else D._MODE_ID[mode]
    nbytes = lib.otf_resize_workspace_bytes(h, w, oh, ow, mode_id)
    if nbytes <= 0:  # the one-pass lanczos tables do not cover extreme down-scales (stages.py splits those)
        pytest.skip("no single-launch tables for this mode / extent")
    ws = torch.empty(nbytes // 4, dtype=torch.int32, device=dev)
    sigma = torch.linspace(2, 25, b, device=dev)
    gray = (torch.arange(b, device=dev) % 2).float()
    off = torch.tensor([5], dtype=torch.int64, device=dev)
    for gp in (None, gray):
        mid = torch.empty(b, c, oh, ow, device=dev)
        want = torch.empty_like(mid)
        got = torch.empty_like(mid)
        _lib.call("otf_resize_f32", _lib.ptr(x), b * c, h, w, _lib.ptr(mid), oh, ow, mode_id, 1, _lib.ptr(ws), nbytes, 0, _lib.stream())
        _lib.call("otf_gaussian_noise_f32", _lib.ptr(mid), b, c, oh, ow, _lib.ptr(sigma), _lib.ptr(gp), None, None, 77, 3, _lib.ptr(off),
                  _lib.NOISE_CLIP, _lib.ptr(want), _lib.stream())
        _lib.call("otf_resize_gauss_f32", _lib.ptr(x), b, c, h, w, _lib.ptr(got), oh, ow, mode_id, 1, _lib.ptr(ws), nbytes, 1,
                  _lib.ptr(sigma), _lib.ptr(gp), 77, 3, _lib.ptr(off), _lib.NOISE_CLIP, _lib.stream())
        assert torch.equal(got, want), (mode, shape, gp is not None)
        assert not torch.equal(want, mid)


def test_degrade_with_injected_fields_matches(dev):
    """degrade() (no crop) through both paths with injected Gaussian fields and Poisson counts."""
    b, size = 2, 80
    d = {k: v.to(dev) for k, v in _data(b, size, 3).items()}
    g = torch.Generator().manual_seed(0)
    h1 = round(size * 0.7)
    plan = {"scale": 4, "gt_size": 64, "order": "classic", "blur1": True, "resize1": {"scale": 0.7, "mode": "bicubic"},
            "noise1": {"kind": "gaussian", "sigma": torch.tensor([5.0, 20.0]), "gray": torch.tensor([0.0, 1.0])},
            "jpeg1": torch.tensor([40.0, 90.0]), "blur2": True, "resize2": {"scale": 1.1, "mode": "lanczos"},
            "noise2": {"kind": "poisson", "scale": torch.tensor([0.5, 2.0]), "gray": torch.tensor([1.0, 0.0])},
            "final_order": "jpeg_first", "resize3_mode": "lanczos", "jpeg2": torch.tensor([55.0, 75.0]), "crop": (0, 0)}
    h2 = int(size / 4 * 1.1)
    inject = {"noise1_color": torch.randn(b, 3, h1, h1, generator=g).to(dev), "noise1_gray": torch.randn(h1, h1, generator=g).to(dev),
              "noise2_counts_color": torch.poisson(torch.full((b, 3, h2, h2), 30.0), generator=g).to(dev),
              "noise2_counts_gray": torch.poisson(torch.full((b, 1, h2, h2), 30.0), generator=g).to(dev)}
    outs = []
    for native in (False, True):
        feed = RealESRGANFeed(OTFOptions(scale=4, gt_size=64), device=dev, manual_seed=1, use_pool=False)
        feed.native_chain = native
        outs.append(feed.degrade(d["gt"], d["kernel1"], d["kernel2"], d["sinc_kernel"], plan, inject))
    assert outs[0].shape == (b, 3, size // 4, size // 4)
    assert torch.equal(outs[0], outs[1])
    # and against the oracle, at the chain's end-to-end bar (within 1 LSB on >= 99.9 % of pixels)
    _, want = O.run_chain_b(d["gt"].cpu(), d["kernel1"].cpu(), d["kernel2"].cpu(), d["sinc_kernel"].cpu(), plan,
                            {k: v.cpu() for k, v in inject.items()})
    diff = (outs[1].cpu()[:, :, :16, :16] - want).abs()
    frac_ok = (diff <= 1 / 255 + 1e-6).float().mean().item()
    print(f"[lsb] native degrade vs oracle: {frac_ok*100:.3f}%")
    assert frac_ok >= 0.999


def test_clean_pass_through_raises_like_the_reference_above_scale_1(dev):
    feed = RealESRGANFeed(OTFOptions(scale=2, gt_size=32, p_clean=1.0), device=dev, manual_seed=0, use_pool=False)
    with pytest.raises(ValueError, match="Scale mismatches"):
        feed.feed_data(_data(2, 64, 0))


def test_stage_list_direct_and_errors(dev):
    x = torch.rand(3, 3, 40, 52, device=dev)
    sl = StageList(x)
    sl.resize("bilinear", size=(20, 30))
    sl.clamp_round()
    got = sl.run()
    want = clamp_round(D.resize_pt(x, "bilinear", size=(20, 30)))
    assert torch.equal(got, want)
    with pytest.raises(ValueError, match="Wrong kernel size"):
        StageList(x).filter2d(torch.rand(3, 4, 4, device=dev))
    with pytest.raises(ValueError, match="scale_factor or size"):
        StageList(x).resize("bilinear")
    with pytest.raises(ValueError, match="empty"):
        StageList(x).run()
    # the C entry point validates what the builder cannot express
    bad = _lib.Stage()
    bad.op = 77
    arr = (_lib.Stage * 1)(bad)
    assert _lib.load().otf_run_stages_workspace_bytes(3, 3, 40, 52, arr, 1) < 0
    with pytest.raises(_lib.OtfError, match="unknown op"):
        _lib.call("otf_run_stages_f32", _lib.ptr(x), 3, 3, 40, 52, arr, 1, None, 0, None, None, _lib.stream())
    ok = _lib.Stage()
    ok.op = _lib.OP_CLAMP_ROUND
    arr = (_lib.Stage * 1)(ok)
    tiny = torch.empty(16, dtype=torch.int32, device=dev)
    with pytest.raises(_lib.OtfError, match="too small"):
        _lib.call("otf_run_stages_f32", _lib.ptr(x), 3, 3, 40, 52, arr, 1, _lib.ptr(tiny), 64, None, None, _lib.stream())
    crop = _lib.Stage()
    crop.op = _lib.OP_CROP_PAIR
    arr = (_lib.Stage * 2)(crop, ok)
    with pytest.raises(_lib.OtfError, match="last stage"):
        _lib.call("otf_run_stages_f32", _lib.ptr(x), 3, 3, 40, 52, arr, 2, _lib.ptr(tiny), 64, None, None, _lib.stream())
    fh, fw = C.c_int(0), C.c_int(0)
    sl = StageList(x)
    sl.resize("area", size=(10, 13))
    out = sl.run()
    assert out.shape == (3, 3, 10, 13)
    del fh, fw
