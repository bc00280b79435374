"""Soak test: hundreds of freshly drawn random plans through feed_data (both stage orders, wide option ranges, all resize
modes incl. extreme down-scales, both noise kinds, pool on) — every output finite, on the 8-bit lattice, right shape."""

import warnings

import pytest
import torch

from oracle import otf_oracle as O
from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed

pytestmark = pytest.mark.gpu
MODES = ("bilinear", "bicubic", "area", "nearest-exact", "lanczos")


@pytest.mark.parametrize("order", ["classic", "fork"])
def test_random_plans_soak(dev, order):
    warnings.simplefilter("ignore")
    opt = OTFOptions(order=order, scale=4, gt_size=96, queue_size=24, p_clean=0.0, lq_usm=(order == "classic"), lq_usm_radius_range=(1, 25),
                     blur_prob=0.8, resize_prob=(0.3, 0.5, 0.2), resize_range=(0.15, 1.5), resize_mode_list=MODES, resize_mode_prob=(0.2,) * 5,
                     gaussian_noise_prob=0.5, noise_range=(1, 30), poisson_scale_range=(0.05, 3), gray_noise_prob=0.4, jpeg_prob=0.8,
                     jpeg_range=(20, 98), blur_prob2=0.8, resize_prob2=(0.3, 0.4, 0.3), resize_range2=(0.3, 1.2), resize_mode_list2=MODES,
                     resize_mode_prob2=(0.2,) * 5, gaussian_noise_prob2=0.5, noise_range2=(1, 25), poisson_scale_range2=(0.05, 2.5),
                     gray_noise_prob2=0.4, jpeg_prob2=0.8, jpeg_range2=(20, 98), resize_mode_list3=MODES, resize_mode_prob3=(0.2,) * 5,
                     lens_distort_prob=0.3, chromatic_aberration_prob=0.3, motion_blur_prob=0.3, motion_blur_kernel_size=(3, 25),
                     demosaic_prob=0.3, sensor_noise_prob=0.3, rolling_shutter_prob=0.3, exposure_prob=0.3, color_temp_prob=0.3,
                     oversharpen_prob=0.3, aliasing_prob=0.3, recompression_prob=0.4, editing_prob=0.4, editing_exposure_prob=0.5,
                     codec_fallback="jpeg")
    feed = RealESRGANFeed(opt, device=dev, manual_seed=11)
    batches = [{"gt": O.synth_gt(4, 128, 128, kind, seed=30 + i).to(dev), "kernel1": O.synth_blur_kernels(4, seed=i).to(dev),
                "kernel2": O.synth_blur_kernels(4, seed=50 + i).to(dev), "sinc_kernel": O.synth_sinc_or_pulse(4, seed=i).to(dev)}
               for i, kind in enumerate(("natural", "uniform", "flat", "twolevel"))]
    seen = set()
    for step in range(240):
        feed.feed_data(batches[step % 4])
        lq, gt = feed.lq, feed.gt
        assert tuple(lq.shape) == (4, 3, 24, 24) and tuple(gt.shape) == (4, 3, 96, 96), (step, feed.last_plan)
        assert torch.isfinite(lq).all() and float(lq.min()) >= 0 and float(lq.max()) <= 1, (step, feed.last_plan)
        lq8 = lq.cpu() * 255
        assert (lq8 - lq8.round()).abs().max().item() < 1e-3, (step, feed.last_plan)
        p = feed.last_plan
        seen.update(k for k in p if k not in ("scale", "gt_size", "order", "crop"))
        if order == "classic":
            seen.update({p["resize1"]["mode"], p["resize2"]["mode"], p["noise1"]["kind"], p["noise2"]["kind"]})
    torch.cuda.synchronize()
    if order == "classic":
        assert set(MODES) <= seen and {"gaussian", "poisson", "usm"} <= seen
    else:
        assert {"lens", "chroma", "motion", "demosaic", "sensor", "shutter", "aliasing", "oversharpen", "compression"} <= seen
