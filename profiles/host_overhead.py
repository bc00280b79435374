"""Host-side cost of one eager feed_data() step with a freshly drawn random plan (what a training loop
pays per iteration): wall time per step with and without waiting for the GPU, and a cProfile of the host
work.      python profiles/host_overhead.py [--batch 8] [--steps 200] [--profile]"""
import argparse
import cProfile
import json
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trainner_redux_b200 import _lib  # noqa: E402
from trainner_redux_b200 import synthetic as S  # noqa: E402
from trainner_redux_b200.kernels import synthesize_kernels  # noqa: E402
from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--gt", type=int, default=256)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--native", type=int, default=-1, help="force the native chain executor on (1) / off (0) if the feed has one")
ap.add_argument("--stable", action="store_true", help="shape-stable schedule (fixed resize scales/modes): the captured-chain path")
ap.add_argument("--graphs", type=int, default=1, help="RealESRGANFeed.use_graphs")
args = ap.parse_args()
dev = torch.device("cuda:0")

opt = OTFOptions(scale=4, gt_size=args.gt - 32, blur_prob=1.0, blur_prob2=0.8, gaussian_noise_prob=0.5, gaussian_noise_prob2=0.5,
                 noise_range=(1, 30), noise_range2=(1, 25), poisson_scale_range=(0.05, 3), poisson_scale_range2=(0.05, 2.5),
                 gray_noise_prob=0.4, gray_noise_prob2=0.4, jpeg_prob=1.0, jpeg_range=(30, 95), jpeg_prob2=1.0, jpeg_range2=(30, 95),
                 resize_range=(0.15, 1.5), resize_range2=(0.3, 1.2), queue_size=args.batch * 4)
if args.stable:  # every draw still happens; only the shapes repeat (bench.py's headline workload)
    import dataclasses

    opt = dataclasses.replace(opt, blur_prob2=1.0, gaussian_noise_prob=1.0, gaussian_noise_prob2=1.0, resize_prob=(0, 0, 1), resize_mode_list=["bicubic"], resize_mode_prob=[1.0], resize_prob2=(0, 0, 1),
                              resize_mode_list2=["bilinear"], resize_mode_prob2=[1.0], resize_mode_list3=["area"], resize_mode_prob3=[1.0],
                              final_jpeg_first_prob=0.0)
feed = RealESRGANFeed(opt, device=dev, manual_seed=0)
feed.use_graphs = bool(args.graphs)
if args.native >= 0 and hasattr(feed, "native_chain"):
    feed.native_chain = bool(args.native)
p1, p2, p3 = S.synth_kernel_params(args.batch, 0)
data = {"gt": S.synth_gt(args.batch, args.gt, args.gt, "uniform", seed=1).to(dev), "kernel1": synthesize_kernels(p1, dev),
        "kernel2": synthesize_kernels(p2, dev), "sinc_kernel": synthesize_kernels(p3, dev)}


def loop(n):
    for _ in range(n):
        feed.feed_data(data)


loop(20)
torch.cuda.synchronize()
l0 = _lib.launch_count
t0 = time.perf_counter()
loop(args.steps)
t_issue = time.perf_counter() - t0  # host time to ISSUE the steps
torch.cuda.synchronize()
t_done = time.perf_counter() - t0
res = {"batch": args.batch, "gt": args.gt, "steps": args.steps, "host_issue_ms_per_step": round(1e3 * t_issue / args.steps, 4),
       "wall_ms_per_step": round(1e3 * t_done / args.steps, 4), "launches_per_step": (_lib.launch_count - l0) / args.steps,
       "pairs_per_s": round(args.batch * args.steps / t_done, 1), "native_chain": getattr(feed, "native_chain", False),
       "stable": args.stable, "graphs": {"on": feed.use_graphs, "hits": feed.graphs.hits, "misses": feed.graphs.misses,
                                         "captures": feed.graphs.captures}}
print(json.dumps(res))
if args.profile:
    pr = cProfile.Profile()
    pr.enable()
    loop(args.steps)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
