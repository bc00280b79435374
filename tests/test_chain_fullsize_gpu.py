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
BAR = 0.999


@pytest.mark.parametrize("gt_kind", ["natural", "uniform"])
@pytest.mark.parametrize("final_order", ["resize_first", "jpeg_first"])
@pytest.mark.parametrize("noise_kind", ["gaussian", "poisson"])
@pytest.mark.parametrize("config", ["c2", "c3"])
def test_full_chain_final_lq_within_1_lsb(config, noise_kind, final_order, gt_kind, dev):
    case = make_case(config, noise_kind, final_order, gt_kind, seed=0)
    want_gt, want_lq, noise = run_oracle(case)
    b, size, scale, crop = CONFIGS[config]
    feed = RealESRGANFeed(OTFOptions(scale=scale, gt_size=crop), device=dev, use_pool=False)
    feed.feed_data({k: case[k] for k in ("gt", "kernel1", "kernel2", "sinc_kernel")}, plan=case["plan"],
                   inject={k: v.to(dev) for k, v in noise.items()})
    torch.cuda.synchronize()
    assert tuple(feed.lq.shape) == (b, 3, crop // scale, crop // scale)
    assert torch.equal(feed.gt.cpu(), want_gt), "GT crop must be bit-identical"
    frac, worst = lsb_fraction(feed.lq, want_lq)
    print(f"[fullsize] {config} {noise_kind} {final_order} {gt_kind}: within 1 LSB on {frac * 100:.4f}% (max {worst:.1f} LSB)")
    assert frac >= BAR, f"{config} {noise_kind} {final_order} {gt_kind}: only {frac * 100:.4f}% within 1 LSB (max {worst:.1f} LSB)"
