"""CPU oracle for the OTF second-order degradation path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* (not a copy) of the arithmetic that
delafer/traiNNer-redux runs inside ``RealESRGANModel.feed_data``.  It exists so
that the CUDA kernels in ``trainner_redux_b200/csrc`` can be checked against the
reference's results on a box where ``/root/reference`` is absent.

Who may import this module: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``trainner_redux_b200/`` imports it; the product path has no CPU fallback.

Why torch-on-CPU and not numpy/C: the reference itself is pure Python whose
arithmetic lives in PyTorch ATen (``F.conv2d``, ``F.pad(reflect)``,
``F.interpolate(antialias=True)``, ``F.avg_pool2d``, ``torch.tensordot``,
``torch.round``) — see SURVEY.md §8c "third-party arithmetic".  ATen is a binary
dependency (pinned ``torch>=2.9.1`` in the reference's pyproject.toml:25;
2.11.0+cu128 installed here).  The oracle calls the *same ATen entry points in
the same order* as the reference call sites cited on every function below, so on
one machine it is bit-identical to the reference (checked by
``oracle/make_goldens.py`` against the imported reference and frozen in
``tests/golden/*.npz``).  An independent float64 numpy restatement of the ATen
resampling formulas lives in ``oracle/np_resample.py`` and is cross-checked in
``tests/test_oracle_cpu.py``.

Parity status: PINNED against outputs of the reference itself run in the build
container (the reference's own tests hold no vectors for this path — SURVEY.md
§4).  Generating script: ``oracle/make_goldens.py``.

All ``file:line`` citations are relative to ``/root/reference/``.
"""

from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np
import torch
from torch import Tensor
from torch.nn import functional as F  # noqa: N812

# ----------------------------------------------------------------------------
# a1  filter2d                       traiNNer/utils/img_process_util.py:8-32
# ----------------------------------------------------------------------------


def filter2d(img: Tensor, kernel: Tensor) -> Tensor:
    """Per-sample KxK cross-correlation with reflect padding.

    Follows traiNNer/utils/img_process_util.py:15-32: odd K only
    (``ValueError("Wrong kernel size")`` at :19-20), ``F.pad(..., "reflect")``
    by K//2 at :18, then a grouped ``F.conv2d`` (no kernel flip).  A kernel with
    leading dimension 1 is shared by the whole batch (:24-28); otherwise sample
    ``b`` uses ``kernel[b]`` on all of its channels (:30-32).
    """
    ksz = kernel.size(-1)
    if ksz % 2 != 1:
        raise ValueError("Wrong kernel size")
    nb, nc, hh, ww = img.size()
    r = ksz // 2
    # Quirk Q6: after DiffJPEG the tensor is NHWC-strided and the
    # reference's .view() raises on CPU.  A value-preserving contiguous() keeps
    # the arithmetic identical.
    padded = F.pad(img.contiguous(), (r, r, r, r), mode="reflect")
    ph, pw = padded.shape[-2:]
    if kernel.size(0) == 1:
        flat = padded.view(nb * nc, 1, ph, pw)
        return F.conv2d(flat, kernel.view(1, 1, ksz, ksz), padding=0).view(nb, nc, hh, ww)
    flat = padded.view(1, nb * nc, ph, pw)
    wgt = kernel.view(nb, 1, ksz, ksz).repeat(1, nc, 1, 1).view(nb * nc, 1, ksz, ksz)
    return F.conv2d(flat, wgt, groups=nb * nc).view(nb, nc, hh, ww)


# ----------------------------------------------------------------------------
# a2  USMSharp                       traiNNer/utils/img_process_util.py:35-55
# ----------------------------------------------------------------------------

_CV_SMALL_GAUSSIAN = {
    # cv2.getGaussianKernel uses these fixed taps for odd ksize<=9 when sigma<=0
    # (OpenCV imgproc/smooth.dispatch.cpp, "small_gaussian_tab").
    1: [1.0],
    3: [0.25, 0.5, 0.25],
    5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
    7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
    9: [4 / 256, 13 / 256, 30 / 256, 51 / 256, 60 / 256, 51 / 256, 30 / 256, 13 / 256, 4 / 256],
}


def cv_gaussian_kernel_1d(ksize: int, sigma: float = 0.0) -> np.ndarray:
    """Restatement of ``cv2.getGaussianKernel(ksize, sigma)`` (float64 column).

    Called by the reference at traiNNer/utils/img_process_util.py:41.  For
    sigma<=0 OpenCV substitutes ``0.3*((ksize-1)*0.5-1)+0.8`` (=8.0 at 51 taps),
    and for odd ksize<=9 it returns a fixed table.
    """
    if sigma <= 0 and ksize in _CV_SMALL_GAUSSIAN and ksize % 2 == 1:
        return np.asarray(_CV_SMALL_GAUSSIAN[ksize], dtype=np.float64).reshape(-1, 1)
    sig = sigma if sigma > 0 else 0.3 * ((ksize - 1) * 0.5 - 1.0) + 0.8
    xs = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    taps = np.exp(-(xs * xs) / (2.0 * sig * sig))
    taps /= taps.sum()
    return taps.reshape(-1, 1)


def usm_kernel(radius: int = 50, sigma: float = 0.0) -> Tensor:
    """(1,K,K) float32 buffer built as at img_process_util.py:37-43."""
    if radius % 2 == 0:
        radius += 1
    col = cv_gaussian_kernel_1d(radius, sigma)
    return torch.FloatTensor(np.dot(col, col.transpose())).unsqueeze_(0)


def usm_sharp(img: Tensor, kernel: Tensor, weight: float = 0.5, threshold: float = 10) -> Tensor:
    """Unsharp mask, img_process_util.py:45-55 (blur, residual, hard mask,
    blurred soft mask, clipped sharpen, blend)."""
    blur = filter2d(img, kernel)
    residual = img - blur
    mask = (torch.abs(residual) * 255 > threshold).float()
    soft = filter2d(mask, kernel)
    sharp = torch.clip(img + weight * residual, 0, 1)
    return soft * sharp + (1 - soft) * img


# ----------------------------------------------------------------------------
# a3  resize_pt                      traiNNer/data/degradations.py:958-1021
# ----------------------------------------------------------------------------

RESIZE_MODES = ("bilinear", "bicubic", "area", "nearest-exact", "lanczos")
_AA = {"bilinear", "bicubic"}  # degradations.py:958


def _lanczos_taps(ratio: float, a: int = 3) -> Tensor:
    """Lanczos-a prefilter taps for down-scaling by ``ratio`` (<1).

    degradations.py:971-978 (`_ramp`: positions j*ratio accumulated in float32,
    |j| <= n-2 with n=ceil(a/ratio+1)) and :961-968 (sinc(t)*sinc(t/a) inside
    |t|<a, normalised to sum 1).
    """
    n = math.ceil(a / ratio + 1)
    half = torch.empty([n])
    acc = 0
    for i in range(n):
        half[i] = acc
        acc += ratio
    pos = torch.cat([-half[1:].flip([0]), half])[1:-1]

    def sinc(t: Tensor) -> Tensor:
        return torch.where(t != 0, torch.sin(math.pi * t) / (math.pi * t), t.new_ones([]))

    inside = torch.logical_and(-a < pos, pos < a)
    w = torch.where(inside, sinc(pos) * sinc(pos / a), pos.new_zeros([]))
    return w / w.sum()


def _lanczos_resample(img: Tensor, size: tuple[int, int]) -> Tensor:
    """degradations.py:982-1001: per-axis reflect-padded Lanczos-3 prefilter at
    full input resolution (only on axes that shrink), then NON-antialiased
    bicubic (Keys a=-0.75) and clamp."""
    n, c, h, w = img.shape
    dh, dw = size
    x = img.reshape([n * c, 1, h, w])
    if dh < h:
        taps = _lanczos_taps(dh / h).to(x.device, x.dtype)
        p = (taps.shape[0] - 1) // 2
        x = F.conv2d(F.pad(x, (0, 0, p, p), "reflect"), taps[None, None, :, None], padding=0)
    if dw < w:
        taps = _lanczos_taps(dw / w).to(x.device, x.dtype)
        p = (taps.shape[0] - 1) // 2
        x = F.conv2d(F.pad(x, (p, p, 0, 0), "reflect"), taps[None, None, None, :], padding=0)
    x = x.view([n, c, h, w])
    return F.interpolate(x, size, mode="bicubic", align_corners=False).clamp(0, 1)


def resize_out_size(h: int, w: int, scale_factor: float = 0, size: tuple[int, int] = (0, 0)) -> tuple[int, int]:
    """degradations.py:1007-1011 — Python banker's ``round`` on H*s, W*s."""
    if scale_factor == 0 and tuple(size) == (0, 0):
        raise ValueError("scale_factor or size is required")
    if scale_factor != 0:
        return (round(h * scale_factor), round(w * scale_factor))
    return (int(size[0]), int(size[1]))


def resize_pt(img: Tensor, mode: str, scale_factor: float = 0, size: tuple[int, int] = (0, 0)) -> Tensor:
    """degradations.py:1004-1021.  ``size=`` is always what reaches ATen, so the
    sampling scale is in/out; bilinear and bicubic run with antialias=True."""
    size = resize_out_size(img.shape[2], img.shape[3], scale_factor, size)
    if mode == "lanczos":
        return _lanczos_resample(img, size)
    return F.interpolate(img, size=size, mode=mode, antialias=mode in _AA).clamp(0, 1)


# ----------------------------------------------------------------------------
# a4  Gaussian noise                 traiNNer/data/degradations.py:569-633,668-698
# ----------------------------------------------------------------------------


def _finish_noise(out: Tensor, clip: bool, rounds: bool) -> Tensor:
    # degradations.py:626-632 (same tail at :835-841, :691-697, :902-908)
    if clip and rounds:
        return torch.clamp((out * 255.0).round(), 0, 255) / 255.0
    if clip:
        return torch.clamp(out, 0, 1)
    if rounds:
        return (out * 255.0).round() / 255.0
    return out


def gaussian_noise_field(
    img: Tensor,
    sigma: Tensor,
    gray_flag: Tensor,
    n_color: Tensor,
    n_gray: Tensor | None,
) -> Tensor:
    """degradations.py:569-605 with the two ``torch.randn`` fields made explicit.

    ``n_color`` is the (B,C,h,w) standard-normal field of :598, ``n_gray`` the
    (h,w) field of :593-596 — ONE field shared by every sample of the batch
    (quirk Q3) — used only when some flag is set.  ``sigma`` (B,) is on the
    0..255 scale.
    """
    b, _, h, w = img.size()
    sg = sigma.view(b, 1, 1, 1)
    gf = gray_flag.view(b, 1, 1, 1)
    use_gray = bool(torch.sum(gf) > 0)
    noise = n_color * sg / 255.0
    if use_gray:
        assert n_gray is not None
        ng = (n_gray * sg / 255.0).view(b, 1, h, w)
        noise = noise * (1 - gf) + ng * gf
    return noise


def add_gaussian_noise(
    img: Tensor,
    sigma: Tensor,
    gray_flag: Tensor,
    n_color: Tensor,
    n_gray: Tensor | None,
    clip: bool = True,
    rounds: bool = False,
) -> Tensor:
    """degradations.py:608-633 / :683-698 with explicit per-sample draws."""
    return _finish_noise(img + gaussian_noise_field(img, sigma, gray_flag, n_color, n_gray), clip, rounds)


# ----------------------------------------------------------------------------
# a5  Poisson noise                  traiNNer/data/degradations.py:762-842,879-909
# ----------------------------------------------------------------------------

GRAY_WEIGHTS = (0.2989, 0.587, 0.114)  # torchvision rgb_to_grayscale, called at :787


def rgb_to_gray(img: Tensor) -> Tensor:
    """torchvision.transforms.functional.rgb_to_grayscale(num_output_channels=1):
    ``(0.2989*r + 0.587*g + 0.114*b)`` evaluated left to right in fp32."""
    r, g, b = img.unbind(dim=-3)
    return (GRAY_WEIGHTS[0] * r + GRAY_WEIGHTS[1] * g + GRAY_WEIGHTS[2] * b).unsqueeze(-3)


def quantise8(x: Tensor) -> Tensor:
    """``clamp(round(x*255),0,255)/255`` — degradations.py:789, :800 and
    realesrgan_model.py:616 (torch.round = half-to-even)."""
    return torch.clamp((x * 255.0).round(), 0, 255) / 255.0


def poisson_vals(q: Tensor) -> Tensor:
    """Per-sample ``2**ceil(log2(#distinct values))`` — degradations.py:791-793
    and :802-804.  ``q`` is already on the 8-bit lattice."""
    b = q.size(0)
    counts = [len(torch.unique(q[i])) for i in range(b)]
    return q.new_tensor([2 ** np.ceil(np.log2(v)) for v in counts]).view(b, 1, 1, 1)


def poisson_lambda(img: Tensor) -> tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Deterministic half of degradations.py:785-805: the quantised colour image,
    its ``vals`` and rate ``q*vals``; the same three for the gray image."""
    qc = quantise8(img)
    vc = poisson_vals(qc)
    qg = quantise8(rgb_to_gray(img))
    vg = poisson_vals(qg)
    return qc, vc, qc * vc, qg, vg, qg * vg


def poisson_noise_field(
    img: Tensor,
    scale: Tensor,
    gray_flag: Tensor,
    poisson_fn: Callable[[Tensor], Tensor] = torch.poisson,
    counts_color: Tensor | None = None,
    counts_gray: Tensor | None = None,
) -> Tensor:
    """degradations.py:762-811.  ``poisson_fn`` replaces ``torch.poisson`` so a
    test can hand both sides the same counts; alternatively pass the counts
    directly.  The noise is measured against the *quantised* image (quirk Q4).
    Order of draws when a gray flag is set: gray field first (:794), colour
    second (:805)."""
    b, _, h, w = img.size()
    gf = gray_flag.view(b, 1, 1, 1)
    use_gray = bool(torch.sum(gf) > 0)
    noise_gray = None
    if use_gray:
        qg = quantise8(rgb_to_gray(img))
        vg = poisson_vals(qg)
        cg = counts_gray if counts_gray is not None else poisson_fn(qg * vg)
        noise_gray = (cg / vg - qg).expand(b, 3, h, w)
    qc = quantise8(img)
    vc = poisson_vals(qc)
    cc = counts_color if counts_color is not None else poisson_fn(qc * vc)
    noise = cc / vc - qc
    if use_gray:
        noise = noise * (1 - gf) + noise_gray * gf
    return noise * scale.view(b, 1, 1, 1)


def add_poisson_noise(
    img: Tensor,
    scale: Tensor,
    gray_flag: Tensor,
    clip: bool = True,
    rounds: bool = False,
    **kw,
) -> Tensor:
    """degradations.py:814-842 / :894-909."""
    return _finish_noise(img + poisson_noise_field(img, scale, gray_flag, **kw), clip, rounds)


# ----------------------------------------------------------------------------
# a6  DiffJPEG                       traiNNer/utils/diffjpeg.py:18-527
# ----------------------------------------------------------------------------

# Annex-K luminance table, stored TRANSPOSED as at diffjpeg.py:18-31.
_Y_TABLE = torch.from_numpy(
    np.array(
        [
            [16, 11, 10, 16, 24, 40, 51, 61],
            [12, 12, 14, 19, 26, 58, 60, 55],
            [14, 13, 16, 24, 40, 57, 69, 56],
            [14, 17, 22, 29, 51, 87, 80, 62],
            [18, 22, 37, 56, 68, 109, 103, 77],
            [24, 35, 55, 64, 81, 104, 113, 92],
            [49, 64, 78, 87, 103, 121, 120, 101],
            [72, 92, 95, 98, 112, 100, 103, 99],
        ],
        dtype=np.float32,
    ).T.copy()
)


def _make_c_table() -> Tensor:
    # diffjpeg.py:32-37: 99 everywhere except the (symmetric) 4x4 corner.
    t = np.full((8, 8), 99, dtype=np.float32)
    t[:4, :4] = np.array([[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]]).T
    return torch.from_numpy(t)


_C_TABLE = _make_c_table()


def jpeg_tables() -> tuple[Tensor, Tensor]:
    return _Y_TABLE.clone(), _C_TABLE.clone()


def _cos_basis(inverse: bool) -> Tensor:
    # diffjpeg.py:155-160 (forward, indices [x,y,u,v]) and :338-343 (inverse,
    # indices [u... the same product with the roles of (x,y) and (u,v) swapped).
    t = np.zeros((8, 8, 8, 8), dtype=np.float32)
    for p in range(8):
        for q in range(8):
            for r in range(8):
                for s in range(8):
                    if inverse:
                        t[p, q, r, s] = np.cos((2 * r + 1) * p * np.pi / 16) * np.cos((2 * s + 1) * q * np.pi / 16)
                    else:
                        t[p, q, r, s] = np.cos((2 * p + 1) * r * np.pi / 16) * np.cos((2 * q + 1) * s * np.pi / 16)
    return torch.from_numpy(t).float()


_ALPHA = np.array([1.0 / np.sqrt(2)] + [1] * 7)
_DCT_T = _cos_basis(False)
_IDCT_T = _cos_basis(True)
_DCT_SCALE = torch.from_numpy(np.outer(_ALPHA, _ALPHA) * 0.25).float()  # :161-164
_IDCT_ALPHA = torch.from_numpy(np.outer(_ALPHA, _ALPHA)).float()  # :335-336
_RGB2YCC = torch.from_numpy(
    np.array(
        [[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5], [0.5, -0.418688, -0.081312]],
        dtype=np.float32,
    ).T.copy()
)  # :70-79
_YCC2RGB = torch.from_numpy(
    np.array([[1.0, 0.0, 1.402], [1, -0.344136, -0.714136], [1, 1.772, 0]], dtype=np.float32).T.copy()
)  # :415-420
_FWD_SHIFT = torch.tensor([0.0, 128.0, 128.0])
_INV_SHIFT = torch.tensor([0, -128.0, -128.0])


_G = globals()  # diffjpeg() re-binds the constants above per call on the input's device


def quality_to_factor(q: float) -> float:
    """diffjpeg.py:48-61."""
    q = 5000.0 / q if q < 50 else 200.0 - q * 2
    return q / 100.0


def jpeg_round(x: Tensor, differentiable: bool) -> Tensor:
    """torch.round, or diffjpeg.py:40-42 ``round(x)+(x-round(x))**3``."""
    if differentiable:
        return torch.round(x) + (x - torch.round(x)) ** 3
    return torch.round(x)


def _split8(plane: Tensor) -> Tensor:
    # diffjpeg.py:135-147
    b, h = plane.shape[0], plane.shape[1]
    return plane.view(b, h // 8, 8, -1, 8).permute(0, 1, 3, 2, 4).contiguous().view(b, -1, 8, 8)


def _merge8(blocks: Tensor, h: int, w: int) -> Tensor:
    # diffjpeg.py:363-377
    b = blocks.shape[0]
    return blocks.view(b, h // 8, w // 8, 8, 8).permute(0, 1, 3, 2, 4).contiguous().view(b, h, w)


def diffjpeg(x: Tensor, quality: float | Tensor, differentiable: bool = False) -> Tensor:
    """The whole of ``DiffJPEG.forward`` (diffjpeg.py:503-527) as one function.

    ``quality`` may be a number or a (B,) tensor; like the reference (:512-514,
    quirk Q1) a tensor argument is overwritten in place with the factors.
    Returns the (B,3,h,w) result as the same NHWC-strided view the reference
    yields (values are what matter; callers may ``.contiguous()``).
    """
    if isinstance(quality, (int, float)):
        factor: float | Tensor = quality_to_factor(quality)
    else:
        factor = quality
        for i in range(factor.size(0)):
            factor[i] = quality_to_factor(factor[i])
    h, w = x.shape[-2:]
    hp = (16 - h % 16) % 16
    wp = (16 - w % 16) % 16
    dv = x.device  # the constants follow the input (CPU: `.to` returns the tensor itself; CUDA: bench.py's eager-ATen arm)
    _RGB2YCC, _FWD_SHIFT, _DCT_SCALE, _DCT_T, _Y_TABLE, _C_TABLE, _IDCT_ALPHA, _IDCT_T, _INV_SHIFT, _YCC2RGB = (
        t.to(dv) for t in (_G["_RGB2YCC"], _G["_FWD_SHIFT"], _G["_DCT_SCALE"], _G["_DCT_T"], _G["_Y_TABLE"], _G["_C_TABLE"],
                           _G["_IDCT_ALPHA"], _G["_IDCT_T"], _G["_INV_SHIFT"], _G["_YCC2RGB"]))
    xp = F.pad(x, (0, wp, 0, hp), mode="constant", value=0)  # :515-522
    hh, ww = h + hp, w + wp
    nb = x.shape[0]

    def table_for(tab: Tensor):
        if isinstance(factor, (int, float)):
            return tab * factor
        return tab.expand(nb, 1, 8, 8) * factor.view(nb, 1, 1, 1)

    # ---- compress (:254-275)
    img = (xp * 255).permute(0, 2, 3, 1)
    ycc = (torch.tensordot(img, _RGB2YCC, dims=1) + _FWD_SHIFT).view(img.shape)  # :89-91
    planar = ycc.permute(0, 3, 1, 2).clone()  # :112
    cb = F.avg_pool2d(planar[:, 1].unsqueeze(1), kernel_size=2, stride=(2, 2), count_include_pad=False)
    cr = F.avg_pool2d(planar[:, 2].unsqueeze(1), kernel_size=2, stride=(2, 2), count_include_pad=False)
    comps = {
        "y": ycc[:, :, :, 0],
        "cb": cb.permute(0, 2, 3, 1).squeeze(3),
        "cr": cr.permute(0, 2, 3, 1).squeeze(3),
    }
    quant = {}
    for name, plane in comps.items():
        blk = _split8(plane) - 128  # :174
        coef = _DCT_SCALE * torch.tensordot(blk, _DCT_T, dims=2)  # :175
        tab = table_for(_Y_TABLE if name == "y" else _C_TABLE)
        quant[name] = jpeg_round(coef.float() / tab, differentiable)  # :207-214
    # ---- decompress (:450-479)
    planes = {}
    for name, qv in quant.items():
        tab = table_for(_Y_TABLE if name == "y" else _C_TABLE)
        deq = qv * tab  # :300-306
        pix = 0.25 * torch.tensordot(deq * _IDCT_ALPHA, _IDCT_T, dims=2) + 128  # :351-352
        if name == "y":
            planes[name] = _merge8(pix, hh, ww)
        else:
            planes[name] = _merge8(pix, int(hh / 2), int(ww / 2))

    def up2(p: Tensor) -> Tensor:  # :397-402 nearest x2
        ph, pw = p.shape[1:3]
        return p.unsqueeze(-1).repeat(1, 1, 2, 2).view(-1, ph * 2, pw * 2)

    stacked = torch.cat(
        [planes["y"].unsqueeze(3), up2(planes["cb"]).unsqueeze(3), up2(planes["cr"]).unsqueeze(3)], dim=3
    )
    rgb = torch.tensordot(stacked + _INV_SHIFT, _YCC2RGB, dims=1).view(stacked.shape).permute(0, 3, 1, 2)  # :430-431
    rgb = torch.min(255 * torch.ones_like(rgb), torch.max(torch.zeros_like(rgb), rgb))  # :476-478
    return (rgb / 255)[:, :, 0:h, 0:w]


# ----------------------------------------------------------------------------
# a7/a8  clamp-round and paired crop
# ----------------------------------------------------------------------------


def clamp_round(x: Tensor) -> Tensor:
    """traiNNer/models/realesrgan_model.py:616 (also :493)."""
    return quantise8(x)


def paired_crop(gt: Tensor, lq: Tensor, gt_patch_size: int, scale: int, top: int, left: int) -> tuple[Tensor, Tensor]:
    """Tensor branch of traiNNer/data/transforms.py:69-144 with the two
    ``random.randint`` draws (:119-120) made explicit."""
    h_lq, w_lq = lq.shape[-2:]
    h_gt, w_gt = gt.shape[-2:]
    p = gt_patch_size // scale
    if h_gt != h_lq * scale or w_gt != w_lq * scale:
        raise ValueError(f"Scale mismatches. GT ({h_gt}, {w_gt}) is not {scale}x ", f"multiplication of LQ ({h_lq}, {w_lq}). None")
    if h_lq < p or w_lq < p:
        raise ValueError(f"LQ ({h_lq}, {w_lq}) is smaller than patch size ({p}, {p}). Please remove None.")
    lq_c = lq[:, :, top : top + p, left : left + p]
    tg, lg = int(top * scale), int(left * scale)
    gt_c = gt[:, :, tg : tg + gt_patch_size, lg : lg + gt_patch_size]
    return gt_c, lq_c


# ----------------------------------------------------------------------------
# a9  pair pool                      traiNNer/models/realesrgan_model.py:403-453
# ----------------------------------------------------------------------------


class PairPool:
    """The training pair pool, with the CPU ``torch.randperm`` (:430) supplied
    by the caller so both sides can replay the same permutation."""

    def __init__(self, queue_size: int) -> None:
        self.queue_size = queue_size
        self.queue_lr: Tensor | None = None
        self.queue_gt: Tensor | None = None
        self.queue_ptr = 0

    def step(self, lq: Tensor, gt: Tensor, perm: Tensor | None = None) -> tuple[Tensor, Tensor]:
        b, c, h, w = lq.size()
        if self.queue_lr is None:
            assert self.queue_size % b == 0, f"queue size {self.queue_size} should be divisible by batch size {b}"
            self.queue_lr = torch.zeros(self.queue_size, c, h, w)
            self.queue_gt = torch.zeros(self.queue_size, *gt.shape[1:])
            self.queue_ptr = 0
        assert self.queue_gt is not None
        if self.queue_ptr == self.queue_size:
            idx = perm if perm is not None else torch.randperm(self.queue_size)
            self.queue_lr = self.queue_lr[idx]
            self.queue_gt = self.queue_gt[idx]
            lq_out = self.queue_lr[0:b].clone()
            gt_out = self.queue_gt[0:b].clone()
            self.queue_lr[0:b] = lq.clone()
            self.queue_gt[0:b] = gt.clone()
            return lq_out, gt_out
        self.queue_lr[self.queue_ptr : self.queue_ptr + b] = lq.clone()
        self.queue_gt[self.queue_ptr : self.queue_ptr + b] = gt.clone()
        self.queue_ptr += b
        return lq, gt


# ----------------------------------------------------------------------------
# chain compositions: order (B) classical, order (A) as shipped
# ----------------------------------------------------------------------------


def _apply_noise(out: Tensor, st: dict, noise: dict, key: str, poisson_fn, fields: dict | None = None) -> Tensor:
    """``add_*_noise_pt`` written as generate + add + tail (degradations.py:625-632 / :833-841 do exactly that), so the
    generated noise field can be exported (``fields[key + "_field"]``) for tests that inject the finished field."""
    kind = st["kind"]
    if kind == "gaussian":
        f = gaussian_noise_field(out, st["sigma"], st["gray"], noise[f"{key}_color"], noise.get(f"{key}_gray"))
    elif kind == "poisson":
        f = poisson_noise_field(out, st["scale"], st["gray"], poisson_fn=poisson_fn,
                                counts_color=noise.get(f"{key}_counts_color"), counts_gray=noise.get(f"{key}_counts_gray"))
    else:
        return out
    if fields is not None:
        fields[f"{key}_field"] = f.contiguous().clone()
    return _finish_noise(out + f, True, False)


def run_chain_b(
    gt: Tensor,
    kernel1: Tensor,
    kernel2: Tensor,
    sinc_kernel: Tensor,
    plan: dict,
    noise: dict,
    poisson_fn: Callable[[Tensor], Tensor] = torch.poisson,
    taps: dict | None = None,
    fields: dict | None = None,
) -> tuple[Tensor, Tensor]:
    """Classical Real-ESRGAN second-order chain, SURVEY.md §3.2(B): every stage
    is one of the reference primitives above, driven by an explicit ``plan``.

    plan keys: scale, gt_size, usm (None | dict(radius,weight,threshold)),
    blur1 (bool), resize1 (None | dict(scale,mode)), noise1 (None |
    dict(kind,sigma|scale,gray)), jpeg1 (None | Tensor[B] quality), blur2,
    resize2 (None | dict(scale,mode)), noise2, final_order ("resize_first" |
    "jpeg_first"), resize3_mode, jpeg2, crop (top,left).
    Returns (gt_crop, lq_crop).  ``taps`` collects every intermediate.
    """
    ori_h, ori_w = gt.shape[2:4]
    sc = plan["scale"]
    out = gt

    def tap(name: str, t: Tensor) -> Tensor:
        if taps is not None:
            taps[name] = t.contiguous().clone()
        return t

    if plan.get("usm"):
        u = plan["usm"]
        out = tap("usm", usm_sharp(out, usm_kernel(u["radius"]), u.get("weight", 0.5), u.get("threshold", 10)))
    if plan.get("blur1"):
        out = tap("blur1", filter2d(out, kernel1))
    if plan.get("resize1"):
        r = plan["resize1"]
        out = tap("resize1", resize_pt(out, scale_factor=r["scale"], mode=r["mode"]))
    if plan.get("noise1"):
        out = tap("noise1", _apply_noise(out, plan["noise1"], noise, "noise1", poisson_fn, fields))
    if plan.get("jpeg1") is not None:
        out = torch.clamp(out, 0, 1)
        out = tap("jpeg1", diffjpeg(out, plan["jpeg1"].clone(), differentiable=False).contiguous())
    if plan.get("blur2"):
        out = tap("blur2", filter2d(out, kernel2))
    if plan.get("resize2"):
        r = plan["resize2"]
        size = (int(ori_h / sc * r["scale"]), int(ori_w / sc * r["scale"]))  # quirk Q5: int() truncation
        out = tap("resize2", resize_pt(out, size=size, mode=r["mode"]))
    if plan.get("noise2"):
        out = tap("noise2", _apply_noise(out, plan["noise2"], noise, "noise2", poisson_fn, fields))

    def final_resize_sinc(t: Tensor) -> Tensor:
        t = tap("resize3", resize_pt(t, size=(ori_h // sc, ori_w // sc), mode=plan["resize3_mode"]))
        return tap("sinc", filter2d(t, sinc_kernel))

    def final_jpeg(t: Tensor) -> Tensor:
        if plan.get("jpeg2") is None:
            return t
        t = torch.clamp(t, 0, 1)
        return tap("jpeg2", diffjpeg(t, plan["jpeg2"].clone(), differentiable=False).contiguous())

    if plan.get("final_order", "resize_first") == "resize_first":
        out = final_jpeg(final_resize_sinc(out))
    else:
        out = final_resize_sinc(final_jpeg(out))
    lq = tap("lq_full", clamp_round(out))
    top, left = plan["crop"]
    gt_c, lq_c = paired_crop(gt, lq, plan["gt_size"], sc, top, left)
    return gt_c.contiguous(), lq_c.contiguous()


def run_chain_a(
    gt: Tensor,
    kernel1: Tensor,
    sinc_kernel: Tensor,
    plan: dict,
    taps: dict | None = None,
) -> tuple[Tensor, Tensor]:
    """The fork's as-shipped order with every probability-0 extra and the
    PIL codec stage removed — traiNNer/models/realesrgan_model.py:525-526,
    :564-574, :616-623: [filter2d(kernel1)] -> resize_pt(size=ori//scale, mode)
    -> filter2d(sinc) -> clamp/round -> paired crop.  ``plan["jpeg"]`` (optional
    Tensor[B]) routes the compression stage to DiffJPEG instead.
    """
    ori_h, ori_w = gt.shape[2:4]
    sc = plan["scale"]
    out = gt
    if plan.get("blur1"):
        out = filter2d(out, kernel1)
    out = resize_pt(out, size=(ori_h // sc, ori_w // sc), mode=plan["resize3_mode"])
    out = filter2d(out, sinc_kernel)
    if plan.get("jpeg") is not None:
        out = diffjpeg(torch.clamp(out, 0, 1), plan["jpeg"].clone(), differentiable=False).contiguous()
    for fmt, q in plan.get("compression", []):
        # the unified compression stage (paragon_otf_degradations.py:39-158): "jpeg" is the reference's PIL round — uint8
        # truncation (:119-120), libjpeg's baseline round trip at int(quality), / 255 (oracle/libjpeg_oracle.py, pinned
        # bit for bit against PIL); other formats are host codecs and pass through
        if fmt == "jpeg" and q is not None:
            from . import libjpeg_oracle

            out = libjpeg_oracle.jpeg_round(out, q)
    lq = clamp_round(out)
    if taps is not None:
        taps["lq_full"] = lq.clone()
    top, left = plan["crop"]
    gt_c, lq_c = paired_crop(gt, lq, plan["gt_size"], sc, top, left)
    return gt_c.contiguous(), lq_c.contiguous()


# ----------------------------------------------------------------------------
# synthetic inputs shared by tests and bench (SURVEY.md §8d "Synthetic inputs")
# ----------------------------------------------------------------------------


def synth_gt(b: int, h: int, w: int, kind: str = "natural", seed: int = 1234) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 3, h, w, generator=g)
    if kind == "uniform":
        return x
    if kind == "natural":
        x = F.avg_pool2d(F.pad(x, (2, 2, 2, 2), mode="reflect"), 5, stride=1)
        ramp = torch.linspace(0, 1, w).view(1, 1, 1, w) * 0.5 + torch.linspace(0, 1, h).view(1, 1, h, 1) * 0.3
        return (x * 0.6 + ramp * 0.5).clamp(0, 1)
    if kind == "flat":
        return torch.full((b, 3, h, w), 0.5)
    if kind == "twolevel":
        return (x > 0.5).float() * 0.75
    raise ValueError(kind)


def synth_blur_kernels(b: int, seed: int = 0, ksize_max: int = 21, kinds: Sequence[str] = ("iso", "aniso", "sinc")) -> Tensor:
    """Blur kernels with the structure of traiNNer/data/degradations.py:62-128
    (Gaussian exp(-0.5 g^T S^-1 g), normalised) and :472-507 (circular low-pass
    sinc via J1), odd true size in 7..ksize_max, zero-padded to 21x21 as at
    traiNNer/data/realesrgan_dataset.py:171-172.  Not bit-matched to the
    reference's generators (host-side producer) — they only feed
    both sides the same realistic kernels."""
    from scipy import special

    rng = np.random.default_rng(seed)
    out = np.zeros((b, 21, 21), dtype=np.float32)
    for i in range(b):
        k = int(rng.choice(np.arange(7, ksize_max + 1, 2)))
        kind = kinds[i % len(kinds)]
        ax = np.arange(k) - (k - 1) / 2
        xx, yy = np.meshgrid(ax, ax)
        if kind == "sinc":
            wc = rng.uniform(np.pi / 3, np.pi)
            rr = np.sqrt(xx**2 + yy**2)
            with np.errstate(divide="ignore", invalid="ignore"):
                ker = wc * special.j1(wc * rr) / (2 * np.pi * rr)
            ker[(k - 1) // 2, (k - 1) // 2] = wc**2 / (4 * np.pi)
        else:
            sx = rng.uniform(0.2, 3.0)
            sy = sx if kind == "iso" else rng.uniform(0.2, 3.0)
            th = 0.0 if kind == "iso" else rng.uniform(-np.pi, np.pi)
            d = np.array([[sx**2, 0], [0, sy**2]])
            u = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            inv = np.linalg.inv(u @ d @ u.T)
            g = np.stack([xx, yy], -1)
            ker = np.exp(-0.5 * np.einsum("...i,ij,...j->...", g, inv, g))
        ker = ker / ker.sum()
        p = (21 - k) // 2
        out[i, p : p + k, p : p + k] = ker
    return torch.from_numpy(out)


def synth_sinc_or_pulse(b: int, seed: int = 1, sinc_prob: float = 0.8) -> Tensor:
    """Final sinc kernel (realesrgan_dataset.py:200-206): 21x21 sinc w.p.
    final_sinc_prob, else the pulse of :110-113."""
    from scipy import special

    rng = np.random.default_rng(seed)
    out = np.zeros((b, 21, 21), dtype=np.float32)
    ax = np.arange(21) - 10
    xx, yy = np.meshgrid(ax, ax)
    rr = np.sqrt(xx**2 + yy**2)
    for i in range(b):
        if rng.uniform() < sinc_prob:
            wc = rng.uniform(np.pi / 3, np.pi)
            with np.errstate(divide="ignore", invalid="ignore"):
                ker = wc * special.j1(wc * rr) / (2 * np.pi * rr)
            ker[10, 10] = wc**2 / (4 * np.pi)
            out[i] = ker / ker.sum()
        else:
            out[i, 10, 10] = 1.0
    return torch.from_numpy(out)
