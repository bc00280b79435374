"""The libjpeg restatement (oracle/libjpeg_oracle.py) against PIL itself — the reference's own call for the "jpeg" round
of its compression stage (paragon_otf_degradations.py:128-146).  Bit for bit."""

import io

import numpy as np
import pytest
import torch

from oracle import libjpeg_oracle as LJ
from oracle import paragon_oracle as P

PIL = pytest.importorskip("PIL.Image")


def _pil(rgb: np.ndarray, q: int) -> np.ndarray:
    buf = io.BytesIO()
    PIL.fromarray(rgb).save(buf, format="JPEG", quality=int(q))
    buf.seek(0)
    return np.array(PIL.open(buf).convert("RGB"))


def _smooth(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 9.0) * np.cos(yy / 13.0), 90 + 0.8 * xx + 0.3 * yy, 200 - 0.5 * xx + 30 * np.sin(yy / 5.0)], -1)
    return np.clip(base + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("size", [(64, 64), (48, 80), (256, 256), (50, 70), (17, 33), (8, 8), (100, 3), (3, 100), (5, 5), (1, 1),
                                  (16, 6), (33, 17), (2, 40), (127, 129)])
def test_oracle_equals_pil(size):
    rng = np.random.default_rng(size[0] * 1000 + size[1])
    h, w = size
    for img in (rng.integers(0, 256, (h, w, 3), dtype=np.uint8), _smooth(rng, h, w), np.full((h, w, 3), 255, np.uint8)):
        for q in (1, 5, 30, 49, 50, 75, 90, 95, 100):
            assert np.array_equal(LJ.jpeg_roundtrip_u8(img, q), _pil(img, q)), (size, q)


def test_stage_equals_the_reference_composition():
    """Float in, float out: clamp, uint8 truncation, codec at int(quality), / 255 — against paragon_oracle.pil_jpeg."""
    img = torch.rand(2, 3, 40, 56, generator=torch.Generator().manual_seed(3)) * 1.2 - 0.1
    for q in (35.7, 80.2, 94.999):
        assert torch.equal(LJ.jpeg_round(img, q), P.pil_jpeg(img, q))
    ql, qc = LJ.quant_tables(50)
    assert np.array_equal(ql, LJ.STD_LUMINANCE) and np.array_equal(qc, LJ.STD_CHROMINANCE)
    assert LJ.quant_tables(100)[0].max() == 1 and LJ.quant_tables(1)[0].max() == 255
