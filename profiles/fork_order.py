"""Throughput of the fork's as-shipped order (A) (realesrgan_model.py:512-616) with the probabilities of the reference's
own ParagonSR option file (scripts/options/train_default_options_paragon_sr_otf.yml:125-176): every step draws a fresh
random plan, so launches are issued eagerly from Python (no graph); inputs are resident on the device.

    python profiles/fork_order.py [--batch 64] [--steps 300] [--cpu-steps 3] [--json out.json]

Also times the CPU oracle of the same order on the host cores (PIL JPEG as in the reference; WebP/AVIF/HEIF rounds pass
through on both sides), for the same kind of GPU-vs-host figure bench.py reports for the classical chain."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trainner_redux_b200 import _lib  # noqa: E402
from trainner_redux_b200 import synthetic as S  # noqa: E402
from trainner_redux_b200.kernels import synthesize_kernels  # noqa: E402
from trainner_redux_b200.realesrgan_feed import HostRNG, OTFOptions, RealESRGANFeed, draw_plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--cpu-steps", type=int, default=3)
ap.add_argument("--json", default=None)
ap.add_argument("--profile", action="store_true", help="cProfile the timed loop (host-side cost per stage)")
ap.add_argument("--order", default="fork", choices=["fork", "classic"],
                help="fork: the as-shipped order (A); classic: the second-order chain with the same option file's random schedule (:125-146)")
args = ap.parse_args()
dev = torch.device("cuda:0")
B, GT = args.batch, 256
opt = OTFOptions(order="fork", scale=4, gt_size=GT - 32, queue_size=B * 2, p_clean=0.0, blur_prob=0.7, oversharpen_prob=0.2,
                 chromatic_aberration_prob=0.1, demosaic_prob=0.1, aliasing_prob=0.2, motion_blur_prob=0.15, lens_distort_prob=0.1,
                 exposure_prob=0.2, color_temp_prob=0.15, sensor_noise_prob=0.25, rolling_shutter_prob=0.05)
if args.order == "classic":  # SURVEY.md §8d Config 2, second run: the option file's random schedule on the classical chain
    opt = OTFOptions(order="classic", scale=4, gt_size=GT - 32, queue_size=B * 2, blur_prob=0.7, resize_prob=(0.2, 0.7, 0.1),
                     resize_range=(0.4, 1.5), gaussian_noise_prob=0.7, noise_range=(0, 15), poisson_scale_range=(0.05, 3),
                     gray_noise_prob=0, jpeg_prob=0.8, jpeg_range=(75, 95), blur_prob2=0.7, resize_prob2=(0.3, 0.4, 0.3),
                     resize_range2=(0.6, 1.2), gaussian_noise_prob2=0.5, noise_range2=(0, 10), poisson_scale_range2=(0.05, 2.5),
                     gray_noise_prob2=0, jpeg_prob2=0.6, jpeg_range2=(75, 95))
feed = RealESRGANFeed(opt, device=dev, manual_seed=0, use_pool=False)
p1, p2, p3 = S.synth_kernel_params(B, 0)
data = {"gt": S.synth_gt(B, GT, GT, "uniform", seed=1).to(dev), "kernel1": synthesize_kernels(p1, dev),
        "kernel2": synthesize_kernels(p2, dev), "sinc_kernel": synthesize_kernels(p3, dev)}
import warnings  # noqa: E402

warnings.simplefilter("ignore")
for _ in range(30):
    feed.feed_data(data)
torch.cuda.synchronize()
l0 = _lib.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if args.profile:
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
t0 = time.perf_counter()
e0.record()
for _ in range(args.steps):
    feed.feed_data(data)
e1.record()
if args.profile:
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
res = {"workload": ("classical second-order chain, random schedule of the ParagonSR option file," if args.order == "classic" else "fork order (A), ParagonSR default probabilities,") + f" batch {B} x {GT}^2 GT x4, random plans, eager launches",
       "pairs_per_s": B / (ms / 1e3), "ms_per_step": ms, "host_issue_ms_per_step": 1e3 * t_issue / args.steps,
       "launches_per_step": (_lib.launch_count - l0) / args.steps}
if args.cpu_steps > 0 and args.order == "classic":
    from oracle import otf_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    gt, k1, k2, sk = (data[k].cpu() for k in ("gt", "kernel1", "kernel2", "sinc_kernel"))
    rng = HostRNG(1)
    t0 = time.perf_counter()
    for _ in range(args.cpu_steps):
        plan = draw_plan(opt, B, GT, GT, rng)
        h1 = round(GT * plan["resize1"]["scale"])
        h2 = int(GT / 4 * plan["resize2"]["scale"])
        noise = {"noise1_color": torch.randn(B, 3, h1, h1), "noise1_gray": torch.randn(h1, h1),
                 "noise2_color": torch.randn(B, 3, h2, h2), "noise2_gray": torch.randn(h2, h2)}
        O.run_chain_b(gt, k1, k2, sk, plan, noise)
    dt = (time.perf_counter() - t0) / args.cpu_steps
    res["cpu_oracle"] = {"pairs_per_s": B / dt, "ms_per_step": dt * 1e3, "cores": os.cpu_count(), "kind": "port",
                         "note": "oracle/otf_oracle.run_chain_b with the same random schedule (torch CPU, all host threads)"}
elif args.cpu_steps > 0:
    from oracle import paragon_oracle as P

    torch.set_num_threads(os.cpu_count() or 1)
    gt, k1, sk = data["gt"].cpu(), data["kernel1"].cpu(), data["sinc_kernel"].cpu()
    rng = HostRNG(1)
    t0 = time.perf_counter()
    for _ in range(args.cpu_steps):
        plan = draw_plan(opt, B, GT, GT, rng)
        inject = {"sensor_noise": torch.randn_like(gt)}
        P.apply_extras_a(gt, k1, sk, plan, inject)
    dt = (time.perf_counter() - t0) / args.cpu_steps
    res["cpu_oracle"] = {"pairs_per_s": B / dt, "ms_per_step": dt * 1e3, "cores": os.cpu_count(), "kind": "port",
                         "note": "oracle/paragon_oracle.apply_extras_a: torch CPU + PIL JPEG + numpy demosaic"}
print(json.dumps(res))
if args.json:
    json.dump(res, open(args.json, "w"), indent=1)
