#!/usr/bin/env python
"""bench.py — degraded LR/HR pairs/sec of the OTF second-order degradation path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md §8d "Config 2"): batch 64 of synthetic 256x256 RGB
GT per GPU, scale 4, the classical second-order chain with every stage on:
  blur1 (per-sample 21x21 kernels from the reference's random_mixed_kernels mix, true sizes 7..21) ->
  bicubic resize x0.75 (192^2) ->
  Gaussian noise sigma~U[1,30], gray 40 % -> DiffJPEG q~U[30,95] -> blur2 -> bilinear resize to 64^2 ->
  Gaussian noise -> area resize to 64^2 -> sinc filter -> DiffJPEG q~U[30,95] -> clamp/round -> paired crop 224/56.
A "step" is one pass of that chain over one batch.  The batch shards per sample, so every rank
runs its own batch with no collective on the data path ("scaling": "weak").

One JSON line on stdout (rank 0).  `value` = pairs/s with inputs resident in HBM, `--streams` (default 4)
batches in flight as CUDA-graph replays on as many streams (`value_one_batch_in_flight` = one stream);
`e2e` = the same metric through RealESRGANFeed.feed_data() from pinned HOST fp32 buffers (H2D of the batch
and its kernels prefetched on a copy stream) with the finished LQ read back every step; `e2e_u8` = the
uint8-GT upload extension; `roofline` describes the dominant kernel (blur1 filter2d) from CUDA-event
timings of graph replays; `cpu_baseline` is the oracle port of the reference pipeline on this box's host
cores (`--impl reference` runs only that).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "degraded LR/HR pairs/sec at 256^2 GT x4"
UNIT = "pairs/s"
GT, SCALE, BATCH, GT_CROP = 256, 4, 64, 224
S1, S2 = 0.75, 1.0
N_ROTATE = int(os.environ.get("OTF_BENCH_ROTATE", "4"))  # distinct input batches rotated through (4 x 50 MB > 126 MB L2)


def peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)", "sm_max_mhz": d.get("sm_max_mhz", 1965.0)}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)", "sm_max_mhz": 1965.0}


def make_inputs(seed: int, batch: int, device=None):
    """Synthetic GT + kernels as SURVEY.md §8d specifies: U[0,1) GT; blur kernels from the reference's
    `random_mixed_kernels` distribution (default kernel_list / kernel_prob, odd sizes 7..21 zero-padded to 21,
    sinc_prob 0.1), final sinc w.p. 0.8 else the pulse.  The parameter tables are drawn once (dataset order);
    the GPU arm evaluates them with the product's own synthesis kernel (device given), the CPU reference arm with
    the oracle of those generators (device None) — the two agree to 1e-7 (tests/test_parity_gpu.py)."""
    from trainner_redux_b200.synthetic import synth_gt, synth_kernel_params

    p1, p2, p3 = synth_kernel_params(batch, seed)
    if device is None:
        from oracle import kernel_synth_oracle as KS  # CPU reference arm only

        k1, k2, k3 = (torch.from_numpy(KS.synthesize(p)) for p in (p1, p2, p3))
    else:
        from trainner_redux_b200.kernels import synthesize_kernels

        k1, k2, k3 = (synthesize_kernels(p, device).cpu() for p in (p1, p2, p3))
    return {"gt": synth_gt(batch, GT, GT, "uniform", seed=1234 + seed), "kernel1": k1, "kernel2": k2, "sinc_kernel": k3}


NOISE = "gaussian"  # --noise poisson: the Poisson variant of SURVEY.md §8d Config 2 (scale ~ U[0.05, 3] / U[0.05, 2.5])


def make_plan(batch: int, seed: int) -> dict:
    g = torch.Generator().manual_seed(seed)
    plan = _make_plan(batch, g)
    if NOISE == "poisson":
        for key, hi in (("noise1", 3.0), ("noise2", 2.5)):
            plan[key] = {"kind": "poisson", "scale": torch.rand(batch, generator=g) * (hi - 0.05) + 0.05, "gray": plan[key]["gray"]}
    return plan


def _make_plan(batch: int, g) -> dict:
    return {
        "scale": SCALE, "gt_size": GT_CROP, "order": "classic", "blur1": True,
        "resize1": {"scale": S1, "mode": "bicubic"},
        "noise1": {"kind": "gaussian", "sigma": torch.rand(batch, generator=g) * 29 + 1, "gray": (torch.rand(batch, generator=g) < 0.4).float()},
        "jpeg1": torch.rand(batch, generator=g) * 65 + 30, "blur2": True,
        "resize2": {"scale": S2, "mode": "bilinear"},
        "noise2": {"kind": "gaussian", "sigma": torch.rand(batch, generator=g) * 24 + 1, "gray": (torch.rand(batch, generator=g) < 0.4).float()},
        "final_order": "resize_first", "resize3_mode": "area", "jpeg2": torch.rand(batch, generator=g) * 65 + 30,
        "crop": (4, 4),
    }


def algorithmic_bytes_per_pair() -> int:
    """Stage-sum model of SURVEY.md §8d (noise counted with its own read+write here because it is a
    separate kernel in this round)."""
    a = 3 * GT * GT * 4
    b1 = int(round(GT * S1)) ** 2 * 3 * 4
    c = (GT // SCALE) ** 2 * 3 * 4
    b2 = int(GT / SCALE * S2) ** 2 * 3 * 4
    return 2 * a + (a + b1) + 2 * b1 + 2 * b1 + (b1 + b2) + (b2 + c) + 2 * c + 2 * c


# ------------------------------------------------------------------ CPU reference arm ----
def cpu_chain_pairs_per_s(batch: int, steps: int, warmup: int, threads: int) -> tuple[float, float]:
    """Times the oracle port of the reference pipeline (same plan, same inputs) on host cores."""
    from oracle import otf_oracle as O

    torch.set_num_threads(threads)
    data = make_inputs(0, batch)
    plan = make_plan(batch, 0)
    g = torch.Generator().manual_seed(0)
    h1 = int(round(GT * S1))
    h2 = int(GT / SCALE * S2)

    def one():
        noise = {"noise1_color": torch.randn(batch, 3, h1, h1, generator=g), "noise1_gray": torch.randn(h1, h1, generator=g),
                 "noise2_color": torch.randn(batch, 3, h2, h2, generator=g), "noise2_gray": torch.randn(h2, h2, generator=g)}
        return O.run_chain_b(data["gt"], data["kernel1"], data["kernel2"], data["sinc_kernel"], plan, noise)

    with torch.no_grad():
        for _ in range(warmup):
            one()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # each step = one bounded sample of the workload, sized so that K + W steps end within ~2 minutes on this host
    # (the port runs at roughly 22 pairs/s per host thread)
    budget_pairs = 120.0 * 22.0 * threads
    batch = BATCH
    while batch > 1 and batch * (args.steps + args.warmup) > budget_pairs:
        batch //= 2
    val, ms = cpu_chain_pairs_per_s(batch, args.steps, args.warmup, threads)
    sample = f"each step = one batch of {batch} x {GT}^2 GT through the oracle port of the reference chain (torch CPU, {threads} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "batch_per_step": batch, "gt": GT, "scale": SCALE},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


SCALING = "weak"


def set_workload(name: str, world: int) -> None:
    """c2 (default, the config the metric is quoted on): B=64 x 256^2 GT x4 per GPU, weak scaling.
    c3 (BASELINE.json configs[2]): ONE batch of 32 x 512^2 GT x2 sharded per sample over the ranks, strong scaling."""
    global GT, SCALE, BATCH, GT_CROP, SCALING, METRIC
    if name == "c3":
        GT, SCALE, GT_CROP, SCALING = 512, 2, 480, "strong"
        BATCH = max(1, 32 // world)
        METRIC = "degraded LR/HR pairs/sec at 512^2 GT x2 (batch 32 sharded over the ranks)"


def workload_name() -> str:
    lq = GT // SCALE
    return (f"Real-ESRGAN OTF second-order chain incl. DiffJPEG + sinc, batch {BATCH} synthetic {GT}^2 GT x{SCALE} per GPU "
            f"(blur1, bicubic x0.75, {NOISE}, jpeg, blur2, bilinear, {NOISE}, area->{lq}^2, sinc, jpeg, clamp/round, "
            f"crop {GT_CROP}/{GT_CROP // SCALE})")


# ------------------------------------------------------------------------- clocks ----
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index: int):
    """Best effort: run this rank (and therefore first-touch its pinned host buffers) on the CPUs of the NUMA
    node the GPU hangs off, so that N ranks do not all pull their H2D traffic through one socket."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:  # noqa: BLE001  (topology files missing in a container: stay unbound)
        return None
    return None


# ---------------------------------------------------------------------------- main ----
def run_b200(args) -> None:
    import torch.distributed as dist

    from trainner_redux_b200 import _lib
    from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed
    from trainner_redux_b200.transforms import crop_pair

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)  # pinned staging buffers should live next to this rank's GPU
    if world > 1:
        # stdout carries the one JSON line only: NCCL's banner ("NCCL version ...", NCCL_DEBUG output) goes to stderr —
        # file descriptor 1 points at stderr while the communicator is created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.load()

    feed = RealESRGANFeed(OTFOptions(scale=SCALE, gt_size=GT_CROP, queue_size=BATCH * 2), device=dev, manual_seed=0, rank=rank,
                          use_pool=False)
    feed.stage_times = {}
    host = [make_inputs(rank * N_ROTATE + i, BATCH, dev) for i in range(N_ROTATE)]
    for d in host:
        for k in d:
            d[k] = d[k].pin_memory()
    devd = [{k: v.to(dev) for k, v in d.items()} for d in host]
    plans = []
    for i in range(N_ROTATE):
        p = make_plan(BATCH, rank * N_ROTATE + i)
        for key in ("noise1", "noise2"):
            for kk in ("sigma", "scale", "gray"):
                if kk in p[key]:
                    p[key][kk] = p[key][kk].to(dev)
        for key in ("jpeg1", "jpeg2"):
            p[key] = p[key].to(dev)
        plans.append(p)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i: int):
        d, p = devd[i % N_ROTATE], plans[i % N_ROTATE]
        lq_full = feed.degrade(d["gt"], d["kernel1"], d["kernel2"], d["sinc_kernel"], p)
        return crop_pair(d["gt"], lq_full, GT_CROP, SCALE, *p["crop"])

    # ---- value: device-resident inputs, CUDA events, max over ranks ----
    # The chain has fixed shapes here, so each rotated input set is captured once into a CUDA graph
    # (one launch of ~16 kernel nodes per step) and the timed region replays the graphs.  --no-graph
    # times the eager Python-driven launches instead.
    for i in range(max(args.warmup, N_ROTATE)):
        step_resident(i)
    barrier()
    graphs, graph_out, kernels_per_step = [], [], None
    if not args.no_graph:
        from trainner_redux_b200.degradations import pin_resize_tables

        pin_resize_tables()  # the warm-up steps filled the weight-table cache; captured graphs may read it
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(N_ROTATE):
                g = torch.cuda.CUDAGraph()
                l_before = _lib.launch_count
                with torch.cuda.graph(g, stream=side):
                    graph_out.append(step_resident(i))
                kernels_per_step = _lib.launch_count - l_before
                graphs.append(g)
        torch.cuda.current_stream().wait_stream(side)
        for i in range(args.warmup):
            graphs[i % N_ROTATE].replay()
    barrier()
    def timed_region(n_streams: int) -> float:
        """K steps, one batch each.  With 2 streams consecutive batches overlap (step i on stream i % 2): every
        batch runs the identical chain, but the tail of one step's small kernels fills with the next step's work —
        what a prefetching loader gets when it degrades batch i+1 while batch i is being consumed."""
        main = torch.cuda.current_stream()
        lanes = [torch.cuda.Stream() for _ in range(n_streams)] if n_streams > 1 else [main]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for ln in lanes:
            if ln is not main:
                ln.wait_stream(main)
        for i in range(args.steps):
            with torch.cuda.stream(lanes[i % len(lanes)]):
                if graphs:
                    graphs[i % N_ROTATE].replay()  # lanes divides N_ROTATE: a graph always replays on the same lane
                else:
                    step_resident(i)
        for ln in lanes:
            if ln is not main:
                main.wait_stream(ln)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count
    n_streams = 1 if args.no_graph else max(1, min(args.streams, N_ROTATE))
    while N_ROTATE % n_streams:
        n_streams -= 1
    ms_total = timed_region(n_streams)
    launches = kernels_per_step * args.steps if graphs else _lib.launch_count - l0
    clocks = sampler.stop() if sampler else None
    value = world * BATCH * args.steps / (ms_total / 1e3)
    ms_single = timed_region(1) if n_streams > 1 else ms_total
    value_single = world * BATCH * args.steps / (ms_single / 1e3)
    # per-stage GPU durations for the roofline.  CUDA events cannot be read back from inside a replayed
    # graph, and around eager launches they would include the Python launch gap, so every stage of the
    # chain (on the tensors of one real pass) is re-captured into its own graph of REP launches and that
    # graph's replay is timed with an event pair: kernel time only, measured live in this run.
    feed.record_stage_fns = not args.no_stage_timing
    step_resident(0)
    feed.record_stage_fns = False
    torch.cuda.synchronize()
    REP = 20
    stage_ms = {}
    side = torch.cuda.Stream()
    for name, fn in feed.stage_fns.items():
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            with torch.cuda.graph(g, stream=side):
                for _ in range(REP):
                    fn()
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        best = float("inf")
        for _ in range(3):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            g.replay()
            s1.record()
            torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1) / REP)
        stage_ms[name] = best
        del g
    feed.stage_fns = {}

    # ---- e2e: feed_data() from pinned HOST buffers + D2H of the LQ batch, wall clock, max over ranks ----
    # As in the reference's training loop the upload of batch i+1 overlaps the degradation of batch i
    # (CUDAPrefetcher: side-stream H2D joined by wait_stream, prefetch_dataloader.py:476-493); the GT
    # crop stays on the device for the network, the LQ result is read back every step.
    lq_host = torch.empty((BATCH, 3, GT_CROP // SCALE, GT_CROP // SCALE), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    plans_host = [make_plan(BATCH, rank * N_ROTATE + i) for i in range(N_ROTATE)]

    def upload(i: int):
        with torch.cuda.stream(copy_stream):
            d = {k: v.to(dev, non_blocking=True) for k, v in host[i % N_ROTATE].items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d, ev

    def run_e2e(n: int):
        nxt = upload(0)
        for i in range(n):
            d, ev = nxt
            if i + 1 < n:
                nxt = upload(i + 1)
            torch.cuda.current_stream().wait_event(ev)
            for v in d.values():
                v.record_stream(torch.cuda.current_stream())
            feed.feed_data(d, plan=plans_host[i % N_ROTATE])
            lq_host.copy_(feed.lq, non_blocking=True)
        torch.cuda.synchronize()

    def timed_e2e() -> float:
        run_e2e(args.warmup)
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * BATCH * args.steps / t.item()

    e2e_value = timed_e2e()
    h2d = sum(v.numel() * v.element_size() for v in host[0].values()) + 6 * BATCH * 4
    d2h = lq_host.numel() * 4
    # extension, reported beside the contract's fp32 number: the same loop with the GT batch uploaded as uint8
    # (what the dataset decodes) and normalised on the device — SURVEY.md §8 f4
    for d in host:
        d["gt"] = (d["gt"] * 255.0).round().clamp(0, 255).to(torch.uint8).pin_memory()
    e2e_u8_value = timed_e2e()
    h2d_u8 = sum(v.numel() * v.element_size() for v in host[0].values()) + 6 * BATCH * 4

    if rank == 0:
        pk = peaks()
        k_ms = stage_ms.get("blur1", float("nan"))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):  # dram__bytes_read+write of this kernel from the committed ncu --set full capture
            with open(tpath) as f:
                traffic = json.load(f).get("filter2d_blur1_dram_bytes_per_launch")
        blur_bytes = BATCH * (2 * 3 * GT * GT * 4 + 21 * 21 * 4)
        achieved = blur_bytes / (k_ms * 1e-3) / 1e9
        true_k2 = float((devd[0]["kernel1"] != 0).flatten(1).sum(1).float().mean().item())
        flops = 2.0 * 21 * 21 * 3 * GT * GT * BATCH
        fma_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        chain_bytes = algorithmic_bytes_per_pair()
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cval, cms = cpu_chain_pairs_per_s(BATCH, 8, 1, threads)
            cpu = {"value": cval, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"8 batches of {BATCH} x {GT}^2 GT (+1 warm-up) through the oracle port of the reference chain, torch CPU {threads} threads, {cms:.0f} ms/batch"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "batch_per_gpu": BATCH, "gt": GT, "scale": SCALE,
                       "l2": f"inputs rotate over {N_ROTATE} distinct batches ({N_ROTATE * BATCH * 3 * GT * GT * 4 / 1e6:.0f} MB > 126 MB L2)",
                       "launch": "eager" if args.no_graph else f"CUDA graph replay ({kernels_per_step} kernels/step), {n_streams} batches in flight on {n_streams} streams; stage_ms: each stage re-captured alone (x20) and replayed",
                       "parallelism": f"per-sample shards x{world}, no collective", "numa_node_rank0": numa},
            "value_one_batch_in_flight": value_single, "ms_per_step_one_batch_in_flight": ms_single / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_u8": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                       "note": "extension: uint8 GT upload + on-device /255 (not the reference's fp32 host format)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": f"filter2d_kernel (blur1, {BATCH}x3x{GT}x{GT}, per-sample kernels zero-padded to 21x21: default kernel_list mix, sizes 7..21)", "bound": "hbm",
                         "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": traffic, "peak_source": pk["source"], "ms_per_launch": k_ms,
                         "fma": {"achieved_tflops_true_taps": flops * (true_k2 / 441.0) / (k_ms * 1e-3) / 1e12,
                                 "frac_of_fma_peak": flops * (true_k2 / 441.0) / (k_ms * 1e-3) / 1e12 / fma_peak,
                                 "peak_tflops": fma_peak, "mean_nonzero_taps": true_k2,
                                 "note": "filter2d is FP32-FMA bound above K~9 (SURVEY.md H1); both roofs reported"}},
            "chain": {"algorithmic_bytes_per_pair": chain_bytes, "achieved_gbs": chain_bytes * value / world / 1e9,
                      "frac_of_hbm_peak": chain_bytes * value / world / 1e9 / pk["hbm_gbs"], "stage_ms": stage_ms},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"], help="c2: 64 x 256^2 x4 per GPU (default); c3: 32 x 512^2 x2 sharded")
    ap.add_argument("--noise", default="gaussian", choices=["gaussian", "poisson"], help="noise kind of both noise stages")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--streams", type=int, default=4, help="batches in flight during the timed region (graph mode)")
    ap.add_argument("--no-stage-timing", action="store_true", help="skip the per-stage re-capture pass (for ncu launch lists)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    set_workload(args.workload, int(os.environ.get("WORLD_SIZE", "1")))
    global NOISE
    NOISE = args.noise
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
