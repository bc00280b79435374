"""CPU oracle for the fork's extra OTF stages — TEST INFRASTRUCTURE ONLY.

A restatement, with explicit parameters instead of global RNG draws, of the tensor-math stages that
this fork's ``RealESRGANModel.feed_data`` runs around the classical primitives:
``traiNNer/models/paragon_otf_degradations.py:251-572`` and the model's own copies at
``traiNNer/models/realesrgan_model.py:193-402``.  Every function calls the same ATen entry points in
the same order as the reference, so on one machine the results are bit-identical to it;
``oracle/make_paragon_goldens.py`` asserts exactly that against the imported reference and freezes the
vectors in ``tests/golden/paragon_goldens.npz`` (parity pinned).  ``draw_extras`` restates the ORDER of
the host-side draws of ``feed_data`` (``realesrgan_model.py:512-604``), pinned the same way.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""

from __future__ import annotations

import io
import math
from typing import Any

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


# ------------------------------------------------------------------ stages ----
def motion_blur_kernel(kernel_size: int, angle: float) -> Tensor:
    """paragon_otf_degradations.py:276-294 — pixels within half a pixel of the line through the centre."""
    center = kernel_size // 2
    c, s = math.cos(math.radians(angle)), math.sin(math.radians(angle))
    k = torch.zeros((kernel_size, kernel_size))
    for i in range(kernel_size):
        for j in range(kernel_size):
            if abs((i - center) * c + (j - center) * s) < 0.5:
                k[i, j] = 1.0
    return k / k.sum()


def motion_blur(img: Tensor, kernel_size: int, angle: float) -> Tensor:
    """:251-273 — F.conv2d with ZERO padding K//2 (an even K grows the image by one row and column)."""
    k = motion_blur_kernel(kernel_size, angle).unsqueeze(0).unsqueeze(0).repeat(img.size(1), 1, 1, 1)
    return F.conv2d(img, k, padding=kernel_size // 2, groups=img.size(1))


def lens_distortion(img: Tensor, strength: float, sqrt=torch.sqrt) -> Tensor:
    """:297-342.  ``sqrt`` lets a test swap ATen's vectorised CPU square root (not correctly rounded: 0.5 % of
    results are one ulp off) for an IEEE one, which is what both CUDA builds (the reference's and ours) use."""
    b, _c, h, w = img.shape
    gx, gy = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    r = sqrt(gx**2 + gy**2)
    rd = r * (1 + strength * r**2)
    rd[r == 0] = 0
    r[r == 0] = 1e-6
    grid = torch.stack([gx * (rd / r), gy * (rd / r)], dim=-1).unsqueeze(0).repeat(b, 1, 1, 1)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="reflection", align_corners=False)


def ieee_sqrt(t: Tensor) -> Tensor:
    return torch.from_numpy(np.sqrt(t.numpy()))


def rolling_shutter(img: Tensor, strength: float) -> Tensor:
    """:417-455."""
    b, _c, h, w = img.shape
    slant = strength * h / w
    gy, gx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    grid = torch.stack([gx + slant * gy, gy], dim=-1).unsqueeze(0).repeat(b, 1, 1, 1)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="reflection", align_corners=False)


def chromatic_aberration(img: Tensor) -> Tensor:
    """:485-523 / realesrgan_model.py:244-310 — R scaled 1.001, B 0.999 (affine_grid + grid_sample, zeros padding)."""
    b, c = img.shape[:2]
    if c != 3:
        return img
    out = []
    for ch, s in ((0, 1.001), (1, None), (2, 0.999)):
        x = img[:, ch : ch + 1]
        if s is not None:
            theta = torch.tensor([[[s, 0, 0], [0, s, 0]]], dtype=torch.float32).repeat(b, 1, 1)
            x = F.grid_sample(x, F.affine_grid(theta, x.size(), align_corners=False), align_corners=False, mode="bilinear")
        out.append(x)
    return torch.clamp(torch.cat(out, dim=1), 0, 1)


def exposure(img: Tensor, factor: float) -> Tensor:
    """:345-362 (and the editing stage, realesrgan_model.py:596-603)."""
    return torch.clamp(img * factor, 0, 1)


def color_temperature(img: Tensor, shift: float) -> Tensor:
    """:365-394."""
    if img.size(1) != 3:
        return img
    r, g, b = img[:, 0:1], img[:, 1:2], img[:, 2:3]
    if shift > 0:
        r = r * (1 + shift * 0.3)
        g = g * (1 + shift * 0.1)
    else:
        b = b * (1 - shift * 0.3)
        g = g * (1 - shift * 0.1)
    return torch.clamp(torch.cat([r, g, b], dim=1), 0, 1)


def color_temperature_gains(shift: float) -> tuple[float, float, float]:
    if shift > 0:
        return 1 + shift * 0.3, 1 + shift * 0.1, 1.0
    return 1.0, 1 - shift * 0.1, 1 - shift * 0.3


def sensor_noise(img: Tensor, std: float, noise: Tensor) -> Tensor:
    """:397-414 with the ``randn_like`` field injected."""
    return torch.clamp(img + noise * std, 0, 1)


def oversharpen(img: Tensor, strength: float) -> Tensor:
    """:458-482 / realesrgan_model.py:193-242 — 5x5 box (zero padding), img + (img - blur) * strength, clamp."""
    c = img.size(1)
    weight = (torch.ones(1, 1, 5, 5) / 25).repeat(c, 1, 1, 1)
    blurred = F.conv2d(img, weight, padding=2, groups=c)
    return torch.clamp(img + (img - blurred) * strength, 0, 1)


def aliasing(img: Tensor, scale: float) -> Tensor:
    """:555-572 / realesrgan_model.py:365-401 — legacy `nearest` down then up."""
    h, w = img.shape[2:4]
    down = F.interpolate(img, size=(int(h * scale), int(w * scale)), mode="nearest")
    return F.interpolate(down, size=(h, w), mode="nearest")


def demosaic_cv2(img: Tensor) -> Tensor:
    """:526-552 as the reference runs it: uint8 truncation, Bayer mosaic, cv2.demosaicing on the host."""
    import cv2

    out = []
    for i in range(img.size(0)):
        a = (img[i].clamp(0, 1).numpy() * 255).astype("uint8").transpose(1, 2, 0)
        h, w, _ = a.shape
        bayer = np.zeros((h, w), dtype=np.uint8)
        bayer[0::2, 0::2] = a[0::2, 0::2, 2]
        bayer[0::2, 1::2] = a[0::2, 1::2, 1]
        bayer[1::2, 0::2] = a[1::2, 0::2, 1]
        bayer[1::2, 1::2] = a[1::2, 1::2, 0]
        d = cv2.demosaicing(bayer, cv2.COLOR_BAYER_BG2BGR)
        out.append((torch.from_numpy(d).float() / 255.0).permute(2, 0, 1))
    return torch.stack(out, dim=0)


def demosaic(img: Tensor) -> Tensor:
    """Restatement of OpenCV's bilinear Bayer demosaic (the third-party algorithm behind :546; opencv-python is unpinned
    in the reference, 4.13.0 here) in numpy integer arithmetic: missing channels are (a+b+1)>>1 / (a+b+c+d+2)>>2 of the
    nearest mosaic samples, border rows/columns copy their inner neighbours, images under 3x3 come out zero.  Pinned
    bit for bit against cv2 (oracle/make_paragon_goldens.py) so that tests need no OpenCV on the GPU box."""
    out = []
    for i in range(img.size(0)):
        a = (img[i].clamp(0, 1).numpy() * 255).astype("uint8").astype(np.int32)  # (3, H, W)
        _c, h, w = a.shape
        if h < 3 or w < 3:
            out.append(torch.zeros(3, h, w))
            continue
        yy, xx = np.mgrid[0:h, 0:w]
        ey, ex = yy % 2 == 0, xx % 2 == 0
        bay = np.where(ey & ex, a[2], np.where(~ey & ~ex, a[0], a[1]))
        p = np.pad(bay, 1, mode="edge")
        at = lambda dy, dx: p[1 + dy : 1 + dy + h, 1 + dx : 1 + dx + w]  # noqa: E731
        cross4 = (at(-1, 0) + at(1, 0) + at(0, -1) + at(0, 1) + 2) >> 2
        diag4 = (at(-1, -1) + at(-1, 1) + at(1, -1) + at(1, 1) + 2) >> 2
        hor2, ver2 = (at(0, -1) + at(0, 1) + 1) >> 1, (at(-1, 0) + at(1, 0) + 1) >> 1
        d = np.stack([np.where(~ey & ~ex, bay, np.where(~ey & ex, hor2, np.where(ey & ~ex, ver2, diag4))),
                      np.where(ey ^ ex, bay, cross4),
                      np.where(ey & ex, bay, np.where(ey & ~ex, hor2, np.where(~ey & ex, ver2, diag4)))])
        cy, cx = np.clip(np.arange(h), 1, h - 2), np.clip(np.arange(w), 1, w - 2)
        d = d[:, cy][:, :, cx].astype(np.uint8)
        out.append(torch.from_numpy(d).float() / 255.0)
    return torch.stack(out, dim=0)


def pil_jpeg(img: Tensor, quality: float) -> Tensor:
    """`_compress_with_format(..., "jpeg")` (:95-158): uint8 truncation, libjpeg through PIL, back to [0,1].
    The product runs libjpeg's round trip on the device (csrc/libjpeg.cu) and must match this bit for bit."""
    from PIL import Image

    out = []
    for i in range(img.size(0)):
        a = (img[i].clamp(0, 1).numpy() * 255).astype("uint8").transpose(1, 2, 0)
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, format="JPEG", quality=int(quality))
        buf.seek(0)
        out.append(torch.from_numpy(np.array(Image.open(buf).convert("RGB"))).float().div(255.0).permute(2, 0, 1))
    return torch.stack(out, dim=0)


# --------------------------------------------------------------- host draws ----
def draw_extras(opt: Any, np_rng: np.random.Generator, py_rng: Any) -> dict:
    """The host-side draws of the fork's ``feed_data`` between the clean-pass coin and the crop, in the reference's
    order (realesrgan_model.py:512-604 calling paragon_otf_degradations.py).  Every stage whose option fields exist
    draws its gate even at probability 0; parameters are drawn only when the gate passes.  ``py_rng`` is Python's
    ``random`` module or a ``random.Random``."""
    u = np_rng.uniform
    p: dict[str, Any] = {}
    if u() < opt.lens_distort_prob:
        p["lens"] = float(u(*opt.lens_distort_strength_range))
    if u() < opt.chromatic_aberration_prob:
        p["chroma"] = True
    if u() < opt.motion_blur_prob:
        ks = py_rng.randint(opt.motion_blur_kernel_size[0], opt.motion_blur_kernel_size[1])
        p["motion"] = (ks, float(u(*opt.motion_blur_angle_range)))
    p["blur1"] = bool(u() < opt.blur_prob)
    if u() < opt.demosaic_prob:
        p["demosaic"] = True
    if u() < opt.sensor_noise_prob:
        p["sensor"] = float(u(*opt.sensor_noise_std_range))
    if u() < opt.rolling_shutter_prob:
        p["shutter"] = float(u(*opt.rolling_shutter_strength_range))
    if u() < opt.exposure_prob:
        p["exposure"] = float(u(*opt.exposure_factor_range))
    if u() < opt.color_temp_prob:
        p["color_temp"] = float(u(*opt.color_temp_shift_range))
    if u() < opt.oversharpen_prob:
        p["oversharpen"] = float(u(*opt.oversharpen_strength))
    if u() < opt.aliasing_prob:
        p["aliasing"] = float(u(*opt.aliasing_scale_range))
    p["resize3_mode"] = py_rng.choices(list(opt.resize_mode_list3), weights=list(opt.resize_mode_prob3))[0]
    # unified compression pipeline (:39-87): format, quality, optional second round
    comp = []
    fmt = str(np_rng.choice(list(opt.compression_formats), p=list(opt.compression_weights)))
    comp.append((fmt, float(u(*getattr(opt, f"compression_{fmt}_range"))) if hasattr(opt, f"compression_{fmt}_range") else None))
    if u() < opt.recompression_prob:
        fmt = str(np_rng.choice(list(opt.recompression_formats), p=list(opt.recompression_weights)))
        comp.append((fmt, float(u(*getattr(opt, f"compression_{fmt}_range"))) if hasattr(opt, f"compression_{fmt}_range") else None))
    p["compression"] = comp
    if u() < opt.editing_prob:  # realesrgan_model.py:590-611
        if u() < opt.editing_exposure_prob:
            p["editing_exposure"] = float(u(*opt.editing_exposure_range))
        u()  # editing_oversharpen_prob gate: drawn, but the branch applies nothing (:606-611)
    return p


def apply_extras_a(gt: Tensor, kernel1: Tensor, sinc_kernel: Tensor, plan: dict, inject: dict | None = None,
                   taps: dict | None = None, jpeg=None) -> Tensor:
    """Order (A) of the fork (realesrgan_model.py:512-616) on explicit parameters; returns the full-size LQ on the
    8-bit lattice.  ``jpeg(out, quality)`` is the codec used for "jpeg" rounds (default: the reference's PIL path);
    other formats pass through (what the reference does when the plugin is missing)."""
    from . import otf_oracle as O

    inject = inject or {}
    jpeg = jpeg or pil_jpeg
    ori_h, ori_w = gt.shape[2:4]
    out = gt

    def tap(name: str, t: Tensor) -> Tensor:
        if taps is not None:
            taps[name] = t.clone()
        return t

    if "lens" in plan:
        out = tap("lens", lens_distortion(out, plan["lens"]))
    if plan.get("chroma"):
        out = tap("chroma", chromatic_aberration(out))
    if "motion" in plan:
        out = tap("motion", motion_blur(out, *plan["motion"]))
    if plan.get("blur1"):
        out = tap("blur1", O.filter2d(out, kernel1))
    if plan.get("demosaic"):
        out = tap("demosaic", demosaic(out))
    if "sensor" in plan:
        out = tap("sensor", sensor_noise(out, plan["sensor"], inject["sensor_noise"]))
    if "shutter" in plan:
        out = tap("shutter", rolling_shutter(out, plan["shutter"]))
    if "exposure" in plan:
        out = tap("exposure", exposure(out, plan["exposure"]))
    if "color_temp" in plan:
        out = tap("color_temp", color_temperature(out, plan["color_temp"]))
    if "oversharpen" in plan:
        out = tap("oversharpen", oversharpen(out, plan["oversharpen"]))
    if "aliasing" in plan:
        out = tap("aliasing", aliasing(out, plan["aliasing"]))
    out = tap("resize3", O.resize_pt(out, size=(ori_h // plan["scale"], ori_w // plan["scale"]), mode=plan["resize3_mode"]))
    out = tap("sinc", O.filter2d(out, sinc_kernel))
    for fmt, q in plan.get("compression", []):
        if fmt == "jpeg" and q is not None:
            out = jpeg(out, q)
    out = tap("compressed", out)
    if "editing_exposure" in plan:
        out = tap("editing_exposure", exposure(out, plan["editing_exposure"]))
    return tap("lq_full", O.clamp_round(out))
