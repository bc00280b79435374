"""BASELINE.json configs[4] (SURVEY.md §8d "Config 5"): share of a training step spent in the OTF degradation
when the GPU feed drives a x4 SR training step under DDP.

    python profiles/train_share.py [--steps 50] [--batch 16] [--json out.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/train_share.py

Shapes follow the reference's SPAN OTF template (options/_templates/train/SPAN/SPAN_OTF_fidelity.yml: batch 16 per GPU,
lq_size 64, scale 4, AdamW 5e-4, Charbonnier loss, bf16 autocast, channels_last); the dataset hands feed_data
(gt_size + 32)^2 GT crops.  The network is SPAN's structure restated here (SpanShapedSR below: training-time Conv3XC
branches, six SPAB blocks, conv_cat fusion, pixel-shuffle tail, 52 channels — traiNNer/archs/span_arch.py:97-345), random
initialised, NOT the reference's arch file.  Networks, losses and optimisers are outside this repo's scope; only their
cost relative to the feed is measured.  Per rank: pinned-host batch -> H2D on a copy stream -> RealESRGANFeed.feed_data (classical order, random
plans, pair pool on) -> forward / loss / backward (DDP all-reduce) / AdamW step.  CUDA events bracket the feed and the
optimisation separately and are read after the run (the host queues work ahead of the GPU, as a training loop does;
--sync-every-step shows the host-bound figure instead); the report is the median over the timed steps, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trainner_redux_b200 import synthetic as S  # noqa: E402
from trainner_redux_b200.kernels import synthesize_kernels  # noqa: E402
from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed  # noqa: E402


class RepConv3(nn.Module):
    """Training-time form of SPAN's re-parameterisable 3x3 convolution (traiNNer/archs/span_arch.py:97-216 `Conv3XC`):
    a 1x1 widening (gain), an unpadded 3x3 on the zero-padded input and a 1x1 narrowing, plus a 1x1 skip branch.  The
    eval-time merged 3x3 of the reference is not needed here: only the training step is timed."""

    def __init__(self, c_in: int, c_out: int, gain: int = 2) -> None:
        super().__init__()
        self.skip = nn.Conv2d(c_in, c_out, 1)
        self.widen = nn.Conv2d(c_in, c_in * gain, 1)
        self.mix = nn.Conv2d(c_in * gain, c_out * gain, 3)
        self.narrow = nn.Conv2d(c_out * gain, c_out, 1)

    def forward(self, x):
        return self.narrow(self.mix(self.widen(nn.functional.pad(x, (1, 1, 1, 1))))) + self.skip(x)


class AttnBlock(nn.Module):
    """SPAN's parameter-free attention block (span_arch.py:219-248 `SPAB`): three RepConv3 with SiLU in between, the
    residual sum gated by sigmoid(out) - 0.5; also returns the first convolution's output (block 6's feeds conv_cat)."""

    def __init__(self, c: int) -> None:
        super().__init__()
        self.c1, self.c2, self.c3 = RepConv3(c, c), RepConv3(c, c), RepConv3(c, c)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        o1 = self.c1(x)
        o3 = self.c3(self.act(self.c2(self.act(o1.clone()))))
        return (o3 + x) * (torch.sigmoid(o3) - 0.5), o1


class SpanShapedSR(nn.Module):
    """The SPAN network restated from its structure (span_arch.py:251-345; option template
    options/_templates/train/SPAN/SPAN_OTF_fidelity.yml: `type: span`, 52 feature channels, x4): head RepConv3, six
    attention blocks, RepConv3, a 1x1 fusion of [head, tail, block 1, first conv of block 6], 3x3 + pixel shuffle.
    Random-initialised; written here, the reference's arch file is neither imported nor copied."""

    def __init__(self, c: int = 52, scale: int = 4) -> None:
        super().__init__()
        self.head = RepConv3(3, c)
        self.blocks = nn.ModuleList(AttnBlock(c) for _ in range(6))
        self.tail = RepConv3(c, c)
        self.fuse = nn.Conv2d(4 * c, c, 1)
        self.up = nn.Sequential(nn.Conv2d(c, 3 * scale * scale, 3, padding=1), nn.PixelShuffle(scale))

    def forward(self, x):
        f = self.head(x)
        y, b1, o1 = f, None, None
        for i, blk in enumerate(self.blocks):
            y, o1 = blk(y)
            if i == 0:
                b1 = y
        return self.up(self.fuse(torch.cat([f, self.tail(y), b1, o1], 1)))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=15)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--json", default=None)
    ap.add_argument("--sync-every-step", action="store_true", help="synchronise after every step (exposes the host-side issue cost of the feed)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, GTS, SC = args.batch, 256, 4
    opt = OTFOptions(scale=SC, gt_size=GTS, queue_size=B * 5, blur_prob=1.0, blur_prob2=0.8, gaussian_noise_prob=0.5, noise_range=(1, 30),
                     poisson_scale_range=(0.05, 3), gray_noise_prob=0.4, jpeg_range=(30, 95), gaussian_noise_prob2=0.5, noise_range2=(1, 25),
                     poisson_scale_range2=(0.05, 2.5), gray_noise_prob2=0.4, jpeg_range2=(30, 95),
                     resize_mode_list=("bilinear", "bicubic", "area"), resize_mode_prob=(1 / 3,) * 3,
                     resize_mode_list2=("bilinear", "bicubic", "area"), resize_mode_prob2=(1 / 3,) * 3,
                     resize_mode_list3=("bilinear", "bicubic", "area"), resize_mode_prob3=(1 / 3,) * 3)
    feed = RealESRGANFeed(opt, device=dev, manual_seed=0, rank=rank)
    host = []
    for i in range(4):  # rotating pinned host batches, as a dataloader would hand over
        p1, p2, p3 = S.synth_kernel_params(B, rank * 4 + i)
        host.append({"gt": S.synth_gt(B, GTS + 32, GTS + 32, "uniform", seed=rank * 4 + i).pin_memory(),
                     "kernel1": synthesize_kernels(p1, dev).cpu().pin_memory(), "kernel2": synthesize_kernels(p2, dev).cpu().pin_memory(),
                     "sinc_kernel": synthesize_kernels(p3, dev).cpu().pin_memory()})
    net = SpanShapedSR(scale=SC).to(dev).to(memory_format=torch.channels_last)
    if world > 1:
        net = nn.parallel.DistributedDataParallel(net, device_ids=[local])
    optim = torch.optim.AdamW(net.parameters(), lr=5e-4, betas=(0.9, 0.99), fused=True)
    copy_stream = torch.cuda.Stream()

    def upload(i):
        with torch.cuda.stream(copy_stream):
            d = {k: v.to(dev, non_blocking=True) for k, v in host[i % 4].items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d, ev

    t_feed, t_opt, evs = [], [], []
    nxt = upload(0)
    t_wall0 = 0.0
    for step in range(args.warmup + args.steps):
        if step == args.warmup:
            torch.cuda.synchronize()
            t_wall0 = time.perf_counter()
        d, ev = nxt
        nxt = upload(step + 1)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.current_stream().wait_event(ev)
        for v in d.values():
            v.record_stream(torch.cuda.current_stream())
        e[0].record()
        feed.feed_data(d)
        e[1].record()
        lq = feed.lq.contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            sr = net(lq)
        loss = torch.sqrt((sr.float() - feed.gt) ** 2 + 1e-6).mean()  # Charbonnier
        optim.zero_grad(set_to_none=True)
        loss.backward()
        optim.step()
        e[2].record()
        if step >= args.warmup:
            evs.append(e)
        if args.sync_every_step:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t_wall0) / max(1, len(evs))
    for e in evs:
        t_feed.append(e[0].elapsed_time(e[1]))
        t_opt.append(e[1].elapsed_time(e[2]))
    res = torch.tensor([statistics.median(t_feed), statistics.median(t_opt), wall * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        f, o, wall_ms = res.tolist()
        line = {"config": "x4 SR training step with GPU OTF feed (BASELINE.json configs[4])", "n_gpus": world, "batch_per_gpu": B,
                "gt": GTS, "scale": SC, "net": "SPAN structure restated (52 ch, 6 SPAB, training-time Conv3XC branches, x4 pixel shuffle), random init, bf16 autocast, channels_last, AdamW fused",
                "params": sum(p.numel() for p in net.parameters()), "feed_ms": f, "optimize_ms": o, "degradation_share": f / (f + o),
                "wall_ms_per_step": wall_ms, "pairs_per_s_job": world * B / (wall_ms / 1e3),
                "timing": "per-step sync" if args.sync_every_step else "events read after the run: the host queues ahead as in a real training loop", "steps": args.steps, "loss": float(loss)}
        print(json.dumps(line), flush=True)
        if args.json:
            with open(args.json, "w") as fh:
                json.dump(line, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
