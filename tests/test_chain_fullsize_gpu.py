"""The north_star end-to-end bar on the BASELINE configs themselves: the whole classical chain through
``RealESRGANFeed.feed_data`` at B=64 x 256^2 x4 (crop 224/56) and B=32 x 512^2 x2 (crop 480/240), both noise kinds,
both final orders, uniform and natural GT, with the oracle's random fields injected — the final 8-bit LQ must be within
1 LSB of the oracle's (reference primitives composed as traiNNer/models/realesrgan_model.py:564-627 composes them)
on >= 99.9 % of pixels, and the GT crop bit-identical."""

import pytest
import torch

from chain_cases import CONFIGS, lsb_fraction, make_case, run_oracle
from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed

pytestmark = pytest.mark.gpu
BAR = 0.999  # north_star: final 8-bit LR within 1 LSB on >= 99.9 % of pixels
# Poisson chains have one more rounding cliff than Gaussian ones: generate_poisson_noise_pt quantises the image to 8 bits
# (degradations.py:789, :800) BEFORE the noise is formed, so a 1e-7 difference in the blurred / resized input flips
# round(x * 255) on ~1.6e-5 of the pixels and each flip moves that pixel's noise by scale/255 — up to 3 LSB — in front
# of BOTH JPEG quantisers.  No fp32 implementation other than the reference's own binary can avoid those flips (ATen's
# CPU and CUDA convolutions already disagree by that much), so the protocol injects what north_star names — "the same
# injected noise fields": the FINISHED field the reference generated (``*_field``), as it does for Gaussian noise.  The
# harder variant (raw Poisson counts injected, the quantiser cliff inside the compared path) is measured too and held to
# 99.8 %; measured 99.83-99.97 % (profiles/r02_parity_fullsize.txt).
BAR_COUNTS = 0.998


@pytest.mark.parametrize("gt_kind", ["natural", "uniform"])
@pytest.mark.parametrize("final_order", ["resize_first", "jpeg_first"])
@pytest.mark.parametrize("noise_kind", ["gaussian", "poisson"])
@pytest.mark.parametrize("config", ["c2", "c3"])
def test_full_chain_final_lq_within_1_lsb(config, noise_kind, final_order, gt_kind, dev):
    case = make_case(config, noise_kind, final_order, gt_kind, seed=0)
    fields: dict = {}
    want_gt, want_lq, noise = run_oracle(case, fields=fields)
    b, size, scale, crop = CONFIGS[config]
    data = {k: case[k] for k in ("gt", "kernel1", "kernel2", "sinc_kernel")}
    variants = [("fields", {k: v.to(dev) for k, v in fields.items()}, BAR)] if noise_kind == "poisson" else []
    variants.append(("counts" if noise_kind == "poisson" else "N(0,1) fields", {k: v.to(dev) for k, v in noise.items()},
                     BAR_COUNTS if noise_kind == "poisson" else BAR))
    for what, inject, bar in variants:
        feed = RealESRGANFeed(OTFOptions(scale=scale, gt_size=crop), device=dev, use_pool=False)
        feed.feed_data(data, plan=case["plan"], inject=inject)
        torch.cuda.synchronize()
        assert tuple(feed.lq.shape) == (b, 3, crop // scale, crop // scale)
        assert torch.equal(feed.gt.cpu(), want_gt), "GT crop must be bit-identical"
        frac, worst = lsb_fraction(feed.lq, want_lq)
        print(f"[fullsize] {config} {noise_kind} {final_order} {gt_kind} (injected {what}): within 1 LSB on {frac * 100:.4f}% "
              f"(max {worst:.1f} LSB)")
        assert frac >= bar, f"{config} {noise_kind} {final_order} {gt_kind} ({what}): only {frac * 100:.4f}% within 1 LSB (max {worst:.1f} LSB)"
