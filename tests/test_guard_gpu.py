"""Out-of-bounds guard tests (compute-sanitizer is not available on the GPU pool): every kernel is run on an input that
sits inside a larger buffer filled with NaN.  A read outside the tensor — even one that is multiplied by a zero weight —
turns the output into NaN, and a result that depends on the guard band differs from the result on a plain tensor.  For
the kernels called through raw pointers the OUTPUT also sits inside a sentinel-filled buffer that must stay intact."""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest
import torch

import trainner_redux_b200 as T
from trainner_redux_b200 import _lib
from trainner_redux_b200 import degradations as D
from trainner_redux_b200 import paragon_otf as PO
from trainner_redux_b200.transforms import crop_pair

pytestmark = pytest.mark.gpu
SENT = 12345.0


def guarded(x: torch.Tensor, dev, pad: int) -> torch.Tensor:
    """A contiguous copy of x that lives in the middle of a NaN-filled buffer (pad floats on either side)."""
    big = torch.full((x.numel() + 2 * pad,), float("nan"), device=dev)
    big[pad : pad + x.numel()] = x.flatten().to(dev)
    return big[pad : pad + x.numel()].view(x.shape)


def check(fn, x: torch.Tensor, dev, what: str, pads=(64, 3)) -> None:
    want = fn(x.to(dev).clone())
    assert torch.isfinite(want).all(), what
    for pad in pads:  # 64: keeps 16-byte alignment (vector / TMA paths); 3: breaks it (scalar paths)
        got = fn(guarded(x, dev, pad))
        assert torch.isfinite(got).all(), f"{what}: NaN from the guard band (pad {pad})"
        if pad % 4 == 0:
            assert torch.equal(got, want), f"{what}: result depends on memory outside the tensor (pad {pad})"
        else:  # a different code path may round differently; it must still agree closely
            assert (got - want).abs().max().item() <= 1e-5, f"{what}: pad {pad}"


@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (1, 3, 50, 77), (3, 3, 128, 96), (2, 3, 33, 40)])
def test_guard_primitives(dev, shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(shape, generator=g)
    b, _c, h, w = shape
    k = torch.rand(b, 21, 21, generator=g)
    k = (k / k.sum((1, 2), keepdim=True)).to(dev)
    k7 = torch.zeros(b, 21, 21)
    k7[:, 7:14, 7:14] = torch.rand(b, 7, 7, generator=g)
    k7 = (k7 / k7.sum((1, 2), keepdim=True)).to(dev)
    check(lambda t: T.filter2d(t, k), x, dev, "filter2d K=21")
    check(lambda t: T.filter2d(t, k7), x, dev, "filter2d K=7")
    check(lambda t: T.filter2d(t, k[:1]), x, dev, "filter2d shared kernel")
    for mode in ("bilinear", "bicubic", "area", "nearest-exact", "lanczos"):
        for s in (0.37, 0.75, 1.0, 1.3, 2.1):
            check(lambda t, mode=mode, s=s: T.resize_pt(t, mode, scale_factor=s), x, dev, f"resize {mode} x{s}")
        check(lambda t, mode=mode: T.resize_pt(t, mode, size=(7, 5)), x, dev, f"resize {mode} -> 7x5")
    sigma = torch.full((b,), 10.0, device=dev)
    gray = torch.tensor([1.0, 0.0, 1.0][:b], device=dev)
    nz = torch.randn(shape, generator=g).to(dev)
    ng = torch.randn(h, w, generator=g).to(dev)
    check(lambda t: D.add_gaussian_noise_pt(t, sigma, gray, noise=nz, noise_gray=ng), x, dev, "gaussian")
    check(lambda t: D.add_poisson_noise_pt(t, 1.0, True, False, gray, generator=D.PhiloxState(3)), x, dev, "poisson")
    jp = T.DiffJPEG(differentiable=False)
    check(lambda t: jp(t, quality=55.0), x, dev, "diffjpeg")
    check(lambda t: T.USMSharp(radius=9).to(dev)(t), x, dev, "usm") if min(h, w) > 10 else None
    check(lambda t: PO.lens_distortion(t, 0.25), x, dev, "lens")
    check(lambda t: PO.rolling_shutter(t, -0.1), x, dev, "shutter")
    check(lambda t: PO.chromatic_aberration(t), x, dev, "chroma")
    check(lambda t: PO.motion_blur(t, 9, 33.0), x, dev, "motion 9")
    check(lambda t: PO.motion_blur(t, 12, 120.0), x, dev, "motion 12")
    check(lambda t: PO.oversharpen(t, 1.4), x, dev, "oversharpen")
    check(lambda t: PO.exposure(t, 1.2), x, dev, "exposure")
    check(lambda t: PO.sensor_noise(t, 0.03, noise=nz), x, dev, "sensor")
    check(lambda t: PO.aliasing(t, 0.7), x, dev, "aliasing")
    check(lambda t: PO.trunc8(t), x, dev, "trunc8")


def test_guard_crop_pair(dev):
    gt = torch.rand(2, 3, 96, 80)
    lq = torch.rand(2, 3, 24, 20)
    lq_d = lq.to(dev)
    want = crop_pair(gt.to(dev), lq_d, 64, 4, 3, 2)
    for pad in (64, 3):
        got = crop_pair(guarded(gt, dev, pad), guarded(lq, dev, pad), 64, 4, 3, 2)
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


@pytest.mark.parametrize("case", [("bicubic", 64, 64, 48, 48), ("bilinear", 50, 77, 81, 40), ("area", 96, 64, 31, 33), ("nearest", 40, 40, 28, 28)])
def test_guard_output_band_raw_abi(dev, case):
    """Raw C-ABI call with the output inside a sentinel band: nothing outside the output extent may be written."""
    mode, h, w, oh, ow = case
    mode_id = {"bicubic": _lib.RESIZE_BICUBIC_AA, "bilinear": _lib.RESIZE_BILINEAR_AA, "area": _lib.RESIZE_AREA, "nearest": _lib.RESIZE_NEAREST}[mode]
    planes = 6
    x = guarded(torch.rand(planes, h, w), dev, 64)
    n_out = planes * oh * ow
    big = torch.full((n_out + 256,), SENT, device=dev)
    out = big[128 : 128 + n_out]
    ws = torch.empty(_lib.load().otf_resize_workspace_bytes(h, w, oh, ow, mode_id) // 4, dtype=torch.int32, device=dev)
    _lib.call("otf_resize_f32", _lib.ptr(x), planes, h, w, C.c_void_p(out.data_ptr()), oh, ow, mode_id, 1, _lib.ptr(ws), ws.numel() * 4, 0,
              _lib.stream())
    torch.cuda.synchronize()
    assert torch.all(big[:128] == SENT) and torch.all(big[128 + n_out :] == SENT), "resize wrote outside its output"
    assert torch.isfinite(out).all() and out.min() >= 0 and out.max() <= 1
    # the f3 raw entry points
    n2 = planes * h * w
    big2 = torch.full((n2 + 256,), SENT, device=dev)
    o2 = big2[128 : 128 + n2]
    _lib.call("otf_warp_f32", _lib.ptr(x), 2, 3, h, w, _lib.WARP_LENS, 0.2, C.c_void_p(o2.data_ptr()), _lib.stream())
    torch.cuda.synchronize()
    assert torch.all(big2[:128] == SENT) and torch.all(big2[128 + n2 :] == SENT) and torch.isfinite(o2).all(), "warp wrote outside its output"
    big2.fill_(SENT)
    kern = np.ascontiguousarray(PO.motion_blur_kernel(7, 45.0))
    _lib.call("otf_taps_zero_f32", _lib.ptr(x), planes, h, w, 7, kern.ctypes.data_as(C.c_void_p), 0, 0.0, C.c_void_p(o2.data_ptr()), _lib.stream())
    torch.cuda.synchronize()
    assert torch.all(big2[:128] == SENT) and torch.all(big2[128 + n2 :] == SENT) and torch.isfinite(o2).all(), "taps_zero wrote outside its output"
    big2.fill_(SENT)
    kd = torch.rand(2, 21, 21, device=dev)
    kd = kd / kd.sum((1, 2), keepdim=True)
    scratch = torch.empty(_lib.load().otf_filter2d_scratch_words(2), dtype=torch.int32, device=dev)
    _lib.call("otf_filter2d_f32", _lib.ptr(x), 2, 3, h, w, _lib.ptr(kd), 2, 21, _lib.ptr(scratch), 0, C.c_void_p(o2.data_ptr()), _lib.stream(), launches=3)
    torch.cuda.synchronize()
    assert torch.all(big2[:128] == SENT) and torch.all(big2[128 + n2 :] == SENT) and torch.isfinite(o2).all(), "filter2d wrote outside its output"


def test_diffjpeg_refuses_to_drop_a_gradient(dev):
    """The reference DiffJPEG is an autograd graph (diffjpeg.py:485-527); the kernel is forward-only and says so."""
    import trainner_redux_b200 as T

    jp = T.DiffJPEG(differentiable=True)
    x = torch.rand(1, 3, 32, 32, device=dev, requires_grad=True)
    with pytest.raises(RuntimeError, match="forward-only"):
        jp(x, quality=80.0)
    with torch.no_grad():
        y = jp(x, quality=80.0)
    assert y.shape == x.shape and not y.requires_grad
    assert jp(x.detach(), quality=80.0).shape == x.shape


def test_feed_accepts_a_frozen_redux_options_shaped_object(dev):
    """ReduxOptions is a frozen struct without `order` / top-level `gt_size`; MoA lives under opt.train."""
    import dataclasses

    from test_host_cpu import _Frozen
    from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed

    fields = {f.name: getattr(OTFOptions(), f.name) for f in dataclasses.fields(OTFOptions)}
    for k in ("gt_size", "order", "use_moa", "moa_augs", "moa_probs"):
        fields.pop(k)
    fields.update(blur_prob=1.0, queue_size=8, compression_formats=["jpeg"], compression_weights=[1.0])
    opt = _Frozen(**fields, datasets={"train": _Frozen(gt_size=64)},
                  train=_Frozen(use_moa=True, moa_augs=["none", "mixup"], moa_probs=[0.5, 0.5], moa_debug=False, moa_debug_limit=0))
    feed = RealESRGANFeed(opt, device=dev, manual_seed=1)
    assert feed.order == "fork" and feed.gt_size == 64 and feed.batch_augment is not None
    g = torch.Generator().manual_seed(0)
    k = torch.zeros(4, 21, 21)
    k[:, 8:13, 8:13] = 1 / 25
    data = {"gt": torch.rand(4, 3, 96, 96, generator=g), "kernel1": k, "kernel2": k, "sinc_kernel": k}
    for _ in range(3):
        feed.feed_data(data)
    assert tuple(feed.gt.shape) == (4, 3, 64, 64) and tuple(feed.lq.shape) == (4, 3, 16, 16)
    assert feed.last_plan["order"] == "fork" and feed.last_plan["gt_size"] == 64
