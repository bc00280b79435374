"""TEST INFRASTRUCTURE — CPU restatement of the JPEG round of the fork's unified compression stage.

The reference (traiNNer/models/paragon_otf_degradations.py:95-158, `_compress_with_format(..., "jpeg")`) truncates the
image to uint8, writes it with ``PIL.Image.save(buffer, format="JPEG", quality=int(q))`` and reads it back with
``Image.open(buffer).convert("RGB")``.  The arithmetic therefore lives in a third-party dependency that is not part of
/root/reference: Pillow (unpinned in the reference's pyproject.toml; 12.2.0 installed here) and the libjpeg-turbo it
bundles (libjpeg API 6.2).  Entropy coding is lossless, so the decoded pixels depend only on the lossy half of baseline
JPEG with libjpeg's defaults, restated here from libjpeg's published sources, function by function:

  jccolor.c   rgb_ycc_convert      fixed-point RGB -> YCbCr (SCALEBITS 16, the "ONE_HALF - 1" rounding of Cb / Cr)
  jcprepct.c  pre_process_data     edge expansion: full-resolution rows to a whole row group, columns to the component's
  jcsample.c  h2v2_downsample      block width (expand_right_edge), 2x2 box with the alternating 1, 2 bias, then every
                                   COMPONENT to whole blocks by repeating its own last row (expand_bottom_edge)
  jfdctint.c  jpeg_fdct_islow      the "slow" integer DCT (CONST_BITS 13, PASS1_BITS 2), output scaled by 8
  jcdctmgr.c  quantize             round-half-up division of |coef| by (quantval << 3)
  jcparam.c   jpeg_set_quality     Annex K tables scaled by jpeg_quality_scaling, forced to the baseline range 1..255
  jidctint.c  jpeg_idct_islow      dequantise, integer inverse DCT, range limit
  jdsample.c  h2v2_fancy_upsample  triangle-filter chroma up-sampling (box replication when the component is <= 2 wide)
  jdcolor.c   ycc_rgb_convert      fixed-point YCbCr -> RGB

Pinned bit for bit against PIL itself by tests/test_libjpeg_cpu.py (14 image sizes from 1x1 to 256x256, noise and
smooth content, qualities 5..100) — PIL is the reference's own call, so this oracle is pinned, not merely plausible.
Only tests/ import this module; the product path is trainner_redux_b200/csrc/libjpeg.cu.
"""

from __future__ import annotations

import numpy as np

STD_LUMINANCE = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99],
    np.int64).reshape(8, 8)
STD_CHROMINANCE = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99],
    np.int64).reshape(8, 8)

F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865, F_0_899976223, F_1_175875602 = 2446, 3196, 4433, 6270, 7373, 9633
F_1_501321110, F_1_847759065, F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 12299, 15137, 16069, 16819, 20995, 25172
CONST_BITS, PASS1_BITS = 13, 2


def quant_tables(quality: int) -> tuple[np.ndarray, np.ndarray]:
    """jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE) -> (luminance, chrominance), natural order."""
    q = max(1, min(100, int(quality)))
    scale = 5000 // q if q < 50 else 200 - 2 * q
    return tuple(np.clip((t * scale + 50) // 100, 1, 255) for t in (STD_LUMINANCE, STD_CHROMINANCE))  # type: ignore[return-value]


def _fix(x: float) -> int:
    return int(x * 65536 + 0.5)


def rgb_to_ycc(rgb: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    half = 1 << 15
    y = (_fix(0.29900) * r + _fix(0.58700) * g + _fix(0.11400) * b + half) >> 16
    cb = (-_fix(0.16874) * r - _fix(0.33126) * g + _fix(0.50000) * b + (128 << 16) + half - 1) >> 16
    cr = (_fix(0.50000) * r - _fix(0.41869) * g - _fix(0.08131) * b + (128 << 16) + half - 1) >> 16
    return y, cb, cr


def _pad_edge(a: np.ndarray, mh: int, mw: int) -> np.ndarray:
    h, w = a.shape
    return np.pad(a, ((0, -h % mh), (0, -w % mw)), mode="edge")


def h2v2_downsample(c: np.ndarray) -> np.ndarray:
    s = c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2]
    return (s + np.where(np.arange(s.shape[1]) % 2 == 0, 1, 2)[None, :]) >> 2


def _descale(x: np.ndarray, n: int) -> np.ndarray:
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d: np.ndarray, first: bool) -> np.ndarray:
    d0, d1, d2, d3, d4, d5, d6, d7 = (d[..., i] for i in range(8))
    t0, t7, t1, t6, t2, t5, t3, t4 = d0 + d7, d0 - d7, d1 + d6, d1 - d6, d2 + d5, d2 - d5, d3 + d4, d3 - d4
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    n = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS
    o = [None] * 8
    if first:
        o[0], o[4] = (t10 + t11) << PASS1_BITS, (t10 - t11) << PASS1_BITS
    else:
        o[0], o[4] = _descale(t10 + t11, PASS1_BITS), _descale(t10 - t11, PASS1_BITS)
    z1 = (t12 + t13) * F_0_541196100
    o[2], o[6] = _descale(z1 + t13 * F_0_765366865, n), _descale(z1 - t12 * F_1_847759065, n)
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * F_1_175875602
    t4, t5, t6, t7 = t4 * F_0_298631336, t5 * F_2_053119869, t6 * F_3_072711026, t7 * F_1_501321110
    z1, z2, z3, z4 = -z1 * F_0_899976223, -z2 * F_2_562915447, -z3 * F_1_961570560 + z5, -z4 * F_0_390180644 + z5
    o[7], o[5], o[3], o[1] = _descale(t4 + z1 + z3, n), _descale(t5 + z2 + z4, n), _descale(t6 + z2 + z3, n), _descale(t7 + z1 + z4, n)
    return np.stack(o, -1)


def fdct_islow(block: np.ndarray) -> np.ndarray:
    """(..., 8, 8) samples minus 128 -> coefficients scaled by 8: rows first, then columns."""
    x = _fdct_1d(block, True)
    return _fdct_1d(x.swapaxes(-1, -2), False).swapaxes(-1, -2)


def quantize(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    d = q << 3
    r = (np.abs(coef) + (d >> 1)) // d
    return np.where(coef < 0, -r, r)


def _idct_1d(v: np.ndarray, first: bool) -> np.ndarray:
    i0, i1, i2, i3, i4, i5, i6, i7 = (v[..., i] for i in range(8))
    z1 = (i2 + i6) * F_0_541196100
    t2, t3 = z1 - i6 * F_1_847759065, z1 + i2 * F_0_765366865
    t0, t1 = (i0 + i4) << CONST_BITS, (i0 - i4) << CONST_BITS
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = i7, i5, i3, i1
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F_1_175875602
    t0, t1, t2, t3 = t0 * F_0_298631336, t1 * F_2_053119869, t2 * F_3_072711026, t3 * F_1_501321110
    z1, z2, z3, z4 = -z1 * F_0_899976223, -z2 * F_2_562915447, -z3 * F_1_961570560 + z5, -z4 * F_0_390180644 + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    n = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS + 3
    return np.stack([_descale(t10 + t3, n), _descale(t11 + t2, n), _descale(t12 + t1, n), _descale(t13 + t0, n),
                     _descale(t13 - t0, n), _descale(t12 - t1, n), _descale(t11 - t2, n), _descale(t10 - t3, n)], -1)


def idct_islow(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Quantised coefficients -> samples 0..255: dequantise, columns first, then rows, + 128, range limit."""
    x = _idct_1d((coef * q).swapaxes(-1, -2), True).swapaxes(-1, -2)
    return np.clip(_idct_1d(x, False) + 128, 0, 255)


def _blocks(a: np.ndarray) -> np.ndarray:
    h, w = a.shape
    return a.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3)


def _unblocks(b: np.ndarray) -> np.ndarray:
    return b.transpose(0, 2, 1, 3).reshape(b.shape[0] * 8, b.shape[1] * 8)


def h2v2_fancy_upsample(c: np.ndarray) -> np.ndarray:
    """(h, w) -> (2h, 2w): 3/4 of the nearer sample + 1/4 of the further one in both directions, edges repeated."""
    h, w = c.shape
    out = np.empty((2 * h, 2 * w), np.int64)
    for v, nb in ((0, np.concatenate([c[:1], c[:-1]], 0)), (1, np.concatenate([c[1:], c[-1:]], 0))):
        cs = 3 * c + nb
        even = (cs * 3 + np.concatenate([cs[:, :1], cs[:, :-1]], 1) + 8) >> 4
        odd = (cs * 3 + np.concatenate([cs[:, 1:], cs[:, -1:]], 1) + 7) >> 4
        even[:, 0], odd[:, -1] = (cs[:, 0] * 4 + 8) >> 4, (cs[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2], out[v::2, 1::2] = even, odd
    return out


def ycc_to_rgb(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    half = 1 << 15
    xr, xb = cr - 128, cb - 128
    r = y + ((_fix(1.40200) * xr + half) >> 16)
    b = y + ((_fix(1.77200) * xb + half) >> 16)
    g = y + ((-_fix(0.34414) * xb + half - _fix(0.71414) * xr) >> 16)
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def jpeg_roundtrip_u8(rgb: np.ndarray, quality: int) -> np.ndarray:
    """(H, W, 3) uint8 -> what ``Image.open(save(rgb, "JPEG", quality)).convert("RGB")`` holds, bit for bit."""
    h, w, _ = rgb.shape
    ql, qc = quant_tables(quality)
    y, cb, cr = (_pad_edge(p, 2, 16) for p in rgb_to_ycc(rgb))
    y, cb, cr = (_pad_edge(p, 8, 8) for p in (y, h2v2_downsample(cb), h2v2_downsample(cr)))
    y, cb, cr = (_unblocks(idct_islow(quantize(fdct_islow(_blocks(p) - 128), q), q)) for p, q in ((y, ql), (cb, qc), (cr, qc)))
    ch, cw = -(-h // 2), -(-w // 2)  # the decoder knows the real extent of every component
    up = h2v2_fancy_upsample if cw > 2 else (lambda c: np.repeat(np.repeat(c, 2, 0), 2, 1))
    return ycc_to_rgb(y[:h, :w], up(cb[:ch, :cw])[:h, :w], up(cr[:ch, :cw])[:h, :w])


def jpeg_round(img: "np.ndarray | object", quality: float) -> "object":
    """The whole stage on a (B, 3, H, W) float tensor in the reference's order: clamp, truncate to uint8, codec round
    at ``int(quality)``, back to float32 / 255 (paragon_otf_degradations.py:119-149)."""
    import torch

    out = []
    for i in range(img.size(0)):
        a = (img[i].clamp(0, 1).numpy() * 255).astype("uint8").transpose(1, 2, 0)
        out.append(torch.from_numpy(jpeg_roundtrip_u8(a, int(quality))).float().div(255.0).permute(2, 0, 1))
    return torch.stack(out, dim=0)
