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

EVERY number of the GPU arm goes through the product's public call, ``RealESRGANFeed.feed_data`` (the drop-in for
traiNNer/models/realesrgan_model.py:455-650), with a plan drawn afresh for every step by the product's own
``draw_plan`` — per-sample sigma / gray flags / JPEG qualities, crop offsets and the Philox position change every step;
only the resize scale of stage 1 is pinned to the workload's 0.75.  The chain's shapes repeat, so feed_data replays its
captured chain (trainner_redux_b200/chain_graph.py) after the second step.

One JSON line on stdout (rank 0):
  value                        pairs/s, inputs resident in HBM (rotating over 4 distinct batches), `--streams` (default 4)
                               feed_data calls in flight on as many streams;
  value_feed_data              the same loop on ONE stream (what a plain training loop gets);
  value_graph_replay[_one_stream]  bare replays of the same captured chains, no host work per step (upper bounds of the two above);
  pcie_measured_gbs            raw pinned-host <-> device copy bandwidth of this box (the ceiling of e2e);
  e2e                          feed_data fed by the package's CUDAPrefetcher from pinned HOST batches in the dataset's
                               decoded format — uint8 GT + the (3,B,8) kernel-parameter table (/255 and
                               kernel synthesis happen on the device) — H2D inside the timed region, the whole LQ batch read
                               back to pinned host memory every step as bytes (CUDAReadback.read(as_u8=True): the LQ lies on the
                               8-bit lattice, lossless), `--streams` batches in flight, wall clock;
  e2e_readback_f32             the same loop with the LQ read back as fp32 (4x the D2H bytes);
  e2e_one_stream               the same loop on ONE stream (fp32 read-back);
  timed_regions_ms, e2e.timed_regions   K <= 100: five timed regions of exactly K steps each, the median is reported;
  e2e_f32                      the same loop fed the reference's host format (fp32 GT + three (B,21,21) kernels,
                               realesrgan_dataset.py:213-219): 4x the bytes, PCIe-bound;
  roofline                     the dominant kernel (blur1 filter2d) from CUDA-event timings of graph replays;
  chain                        stage-sum bytes per pair, per-launch times (`stage_ms`) and their fractions of the HBM roof;
  cpu_baseline / --impl reference   the oracle port of the reference pipeline on this box's host cores;
  reference_torch_cuda         the same oracle (the reference's own ATen call sequence) run on the CUDA device: PyTorch-eager
                               on the same B200, the GPU path a user of the reference has today;
  parity                       final-LQ agreement of feed_data with the oracle on this workload (injected noise fields);
  poisson, c3                  short runs of Config 2's Poisson variant and of BASELINE configs[2] (32 x 512^2 x2).
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

UNIT = "pairs/s"
S1, S2 = 0.75, 1.0
N_ROTATE = int(os.environ.get("OTF_BENCH_ROTATE", "4"))  # distinct input batches rotated through (4 x 50 MB > 126 MB L2)
E2E_SHORT_STEPS, E2E_REPS_SHORT = 100, 5  # e2e timed regions of <= 100 steps are measured 5 times, median reported (Arm.e2e)
E2E_SLOTS = int(os.environ.get("OTF_BENCH_E2E_SLOTS", "2"))  # static upload slots of the e2e prefetcher (one captured chain each)


_RECORD_OUT = None


def emit(line: dict) -> None:
    """The one JSON record, on the process's real stdout (see main)."""
    out = _RECORD_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class Workload:
    """c2 (default, the config the metric is quoted on): B=64 x 256^2 GT x4 per GPU, weak scaling.
    c3 (BASELINE.json configs[2]): ONE batch of 32 x 512^2 GT x2 sharded per sample over the ranks, strong scaling."""

    def __init__(self, name: str, world: int, noise: str = "gaussian") -> None:
        self.name, self.noise, self.world = name, noise, world
        if name == "c3":
            self.gt, self.scale, self.crop, self.scaling = 512, 2, 480, "strong"
            self.batch = max(1, 32 // world)
            self.metric = "degraded LR/HR pairs/sec at 512^2 GT x2 (batch 32 sharded over the ranks)"
        else:
            self.gt, self.scale, self.crop, self.scaling, self.batch = 256, 4, 224, "weak", 64
            self.metric = "degraded LR/HR pairs/sec at 256^2 GT x4"

    def describe(self) -> str:
        lq = self.gt // self.scale
        return (f"Real-ESRGAN OTF second-order chain incl. DiffJPEG + sinc, batch {self.batch} synthetic {self.gt}^2 GT x{self.scale} "
                f"per GPU (blur1, bicubic x0.75, {self.noise}, jpeg, blur2, bilinear, {self.noise}, area->{lq}^2, sinc, jpeg, "
                f"clamp/round, crop {self.crop}/{self.crop // self.scale})")

    def config(self, batch: int | None = None) -> dict:
        """Identical keys in both arms (the driver compares them)."""
        return {"workload": self.describe(), "batch_per_step": batch or self.batch, "gt": self.gt, "scale": self.scale,
                "noise": self.noise,
                "l2": f"inputs rotate over {N_ROTATE} distinct batches ({N_ROTATE * self.batch * 3 * self.gt * self.gt * 4 / 1e6:.0f} MB > 126 MB L2)",
                "parallelism": f"per-sample shards x{self.world}, no collective"}

    def options(self):
        from trainner_redux_b200.realesrgan_feed import OTFOptions

        g = 1.0 if self.noise == "gaussian" else 0.0
        return OTFOptions(
            scale=self.scale, gt_size=self.crop, queue_size=self.batch * 2, blur_prob=1.0, blur_prob2=1.0, gaussian_noise_prob=g,
            noise_range=(1, 30), poisson_scale_range=(0.05, 3.0), gray_noise_prob=0.4, jpeg_prob=1.0, jpeg_range=(30, 95),
            gaussian_noise_prob2=g, noise_range2=(1, 25), poisson_scale_range2=(0.05, 2.5), gray_noise_prob2=0.4, jpeg_prob2=1.0,
            jpeg_range2=(30, 95), resize_prob=(0, 0, 1), resize_mode_list=["bicubic"], resize_mode_prob=[1.0], resize_prob2=(0, 0, 1),
            resize_mode_list2=["bilinear"], resize_mode_prob2=[1.0], resize_mode_list3=["area"], resize_mode_prob3=[1.0],
            final_jpeg_first_prob=0.0)

    def stage_bytes_per_pair(self) -> dict:
        """Algorithmic bytes of each launch of the step (unique input + unique output at fp32; a fused launch counts
        its external input and final output), keyed like `stage_ms`."""
        a = 3 * self.gt * self.gt * 4
        b1 = int(round(self.gt * S1)) ** 2 * 3 * 4
        c = (self.gt // self.scale) ** 2 * 3 * 4
        b2 = int(self.gt / self.scale * S2) ** 2 * 3 * 4
        crop = 3 * (self.crop // self.scale) ** 2 * 4
        return {"blur1": 2 * a, "fused resize1+noise1": a + b1, "jpeg1": 2 * b1, "blur2": 2 * b1, "fused resize2+noise2": b1 + b2,
                "sinc": 2 * c, "fused jpeg2+round+lq_crop": c + crop}

    def algorithmic_bytes_per_pair(self) -> int:
        """Stage-sum model of SURVEY.md §8d."""
        a = 3 * self.gt * self.gt * 4
        b1 = int(round(self.gt * S1)) ** 2 * 3 * 4
        c = (self.gt // self.scale) ** 2 * 3 * 4
        b2 = int(self.gt / self.scale * S2) ** 2 * 3 * 4
        return 2 * a + (a + b1) + 2 * b1 + 2 * b1 + (b1 + b2) + (b2 + c) + 2 * c + 2 * c


def peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)", "sm_max_mhz": d.get("sm_max_mhz", 1965.0)}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)", "sm_max_mhz": 1965.0}


def make_inputs(wl: Workload, seed: int, device=None) -> dict:
    """Synthetic GT + kernels as SURVEY.md §8d specifies: U[0,1) GT; blur kernels from the reference's
    `random_mixed_kernels` distribution (default kernel_list / kernel_prob, odd sizes 7..21 zero-padded to 21,
    sinc_prob 0.1), final sinc w.p. 0.8 else the pulse.  The parameter tables are drawn once (dataset order);
    the GPU arm evaluates them with the product's own synthesis kernel (device given), the CPU reference arm with
    the oracle of those generators (device None) — the two agree to 1e-7 (tests/test_parity_gpu.py)."""
    from trainner_redux_b200.synthetic import synth_gt, synth_kernel_params

    p1, p2, p3 = synth_kernel_params(wl.batch, seed)
    if device is None:
        from oracle import kernel_synth_oracle as KS  # CPU reference arm only

        k1, k2, k3 = (torch.from_numpy(KS.synthesize(p)) for p in (p1, p2, p3))
    else:
        from trainner_redux_b200.kernels import synthesize_kernels

        k1, k2, k3 = (synthesize_kernels(p, device).cpu() for p in (p1, p2, p3))
    return {"gt": synth_gt(wl.batch, wl.gt, wl.gt, "uniform", seed=1234 + seed), "kernel1": k1, "kernel2": k2, "sinc_kernel": k3,
            "kernel_params": (p1, p2, p3)}


def fixed_plan(wl: Workload, seed: int) -> dict:
    """The workload's plan with seeded per-sample draws (the CPU arms and the parity check use it)."""
    from trainner_redux_b200.realesrgan_feed import HostRNG, draw_plan

    plan = draw_plan(wl.options(), wl.batch, wl.gt, wl.gt, HostRNG(1000 + seed))
    plan["resize1"] = {"scale": S1, "mode": "bicubic"}
    plan["crop"] = (4, 4)
    return plan


# ------------------------------------------------------------------ CPU reference arm ----
def oracle_chain(wl: Workload, batch: int, steps: int, warmup: int, threads: int, device: str = "cpu", keep: dict | None = None):
    """Times the oracle port of the reference pipeline (same plan shape, same inputs): (pairs/s, ms/step)."""
    from oracle import otf_oracle as O

    torch.set_num_threads(threads)
    data = {k: (v[:batch] if torch.is_tensor(v) else v) for k, v in make_inputs(wl, 0).items()}
    plan = fixed_plan(wl, 0)
    for key in ("noise1", "noise2"):
        plan[key] = {k: (v[:batch] if torch.is_tensor(v) else v) for k, v in plan[key].items()}
    plan["jpeg1"], plan["jpeg2"] = plan["jpeg1"][:batch], plan["jpeg2"][:batch]
    g = torch.Generator().manual_seed(0)
    h1, h2 = int(round(wl.gt * S1)), int(wl.gt / wl.scale * S2)
    dev = torch.device(device)

    def mv(v):
        if torch.is_tensor(v):
            return v.to(dev)
        if isinstance(v, dict):
            return {k: mv(x) for k, x in v.items()}
        return v

    d_dev, plan_dev = mv({k: data[k] for k in ("gt", "kernel1", "kernel2", "sinc_kernel")}), mv(plan)

    def one():
        noise = {}
        if wl.noise == "gaussian":
            noise = {"noise1_color": torch.randn(batch, 3, h1, h1, generator=g), "noise1_gray": torch.randn(h1, h1, generator=g),
                     "noise2_color": torch.randn(batch, 3, h2, h2, generator=g), "noise2_gray": torch.randn(h2, h2, generator=g)}
        fields: dict = {} if keep is not None else None  # type: ignore[assignment]
        out = O.run_chain_b(d_dev["gt"], d_dev["kernel1"], d_dev["kernel2"], d_dev["sinc_kernel"], plan_dev, mv(noise), fields=fields)
        if keep is not None:
            keep.update(plan=plan, data=data, fields=fields, out=out)
        return out

    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    with torch.no_grad():
        for _ in range(warmup):
            one()
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        sync()
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


def run_reference(args, wl: Workload) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # each step = one bounded sample of the workload, sized so that K + W steps end within ~2 minutes on this host
    # (the port runs at roughly 22 pairs/s per host thread at 256^2)
    budget_pairs = 120.0 * 22.0 * threads * (256.0 / wl.gt) ** 2
    batch = wl.batch
    while batch > 1 and batch * (args.steps + args.warmup) > budget_pairs:
        batch //= 2
    val, ms = oracle_chain(wl, batch, args.steps, args.warmup, threads)
    sample = f"each step = one batch of {batch} x {wl.gt}^2 GT through the oracle port of the reference chain (torch CPU, {threads} threads)"
    line = {
        "impl": "reference", "metric": wl.metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": wl.config(),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------- clocks ----
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
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
        # the sampler spans warm-up, the timed regions and the e2e loops: report the clock under load (upper half)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def gpu_numa_node(index: int):
    """(node, why): the NUMA node the GPU hangs off, from sysfs."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        path = f"/sys/bus/pci/devices/{bdf}/numa_node"
        if not os.path.exists(path):
            return None, f"{path} does not exist (PCI topology not exposed in this container)"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None, f"{path} = {node}: the platform reports no NUMA affinity for this device (single-node guest)"
        return node, "sysfs"
    except Exception as e:  # noqa: BLE001
        return None, f"{type(e).__name__}: {e}"


def bind_to_gpu_numa_node(index: int):
    """Best effort: run this rank (and therefore first-touch its pinned host buffers) on the CPUs of the NUMA
    node the GPU hangs off, so that N ranks do not all pull their H2D traffic through one socket."""
    node, why = gpu_numa_node(index)
    if node is None:
        return None, why
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node, f"bound to {len(allowed)} CPUs of node {node}"
        return None, f"none of node {node}'s CPUs is in this process's affinity mask"
    except Exception as e:  # noqa: BLE001
        return None, f"{type(e).__name__}: {e}"


# ---------------------------------------------------------------------------- main ----
class Arm:
    """One workload on this rank: device-resident inputs, a feed, and the loops that are timed."""

    def __init__(self, wl: Workload, dev, rank: int, world: int, dist) -> None:
        from trainner_redux_b200.realesrgan_feed import RealESRGANFeed, draw_plan

        self.wl, self.dev, self.rank, self.world, self.dist = wl, dev, rank, world, dist
        self.opt = wl.options()
        self.feed = RealESRGANFeed(self.opt, device=dev, manual_seed=0, rank=rank, use_pool=False)
        # this process builds several prefetchers in a row (three e2e loops, each with its own upload slots, after the
        # device-resident loops): enough capture credits for all of their chains even in a 20-step run
        self.feed.graphs.credits = self.feed.graphs.max_credits = 32.0
        self.feed.graphs.capacity = 32
        self.draw_plan = draw_plan
        self.host = [make_inputs(wl, rank * N_ROTATE + i, dev) for i in range(N_ROTATE)]
        keys = ("gt", "kernel1", "kernel2", "sinc_kernel")
        self.devd = [{k: d[k].to(dev) for k in keys} for d in self.host]

    def plan(self) -> dict:
        """A fresh plan for this step from the product's own draw_plan: every host draw of the chain happens (per-sample
        sigma, gray flags, qualities, crop offsets); the workload pins stage 1 to its x0.75 bicubic."""
        p = self.draw_plan(self.opt, self.wl.batch, self.wl.gt, self.wl.gt, self.feed.rng)
        p["resize1"] = {"scale": S1, "mode": "bicubic"}
        return p

    def step(self, i: int) -> None:
        self.feed.feed_data(self.devd[i % N_ROTATE], plan=self.plan())

    def barrier(self) -> None:
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, steps: int, n_streams: int, replay_only: bool = False) -> float:
        """K steps, one batch each; max over ranks of the device time (CUDA events on the launching stream).  With
        several streams consecutive feed_data calls run on different streams (step i on stream i % n): every call is the
        whole chain of its own batch, but the tail of one step's small kernels overlaps the next step's work — what a
        prefetching loader gets when it degrades batch i+1 while batch i is being consumed."""
        main = torch.cuda.current_stream()
        lanes = [torch.cuda.Stream() for _ in range(n_streams)] if n_streams > 1 else [main]
        entries = list(self.feed.graphs.entries.values())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for ln in lanes:
            if ln is not main:
                ln.wait_stream(main)
        for i in range(steps):
            with torch.cuda.stream(lanes[i % len(lanes)]):
                if replay_only:
                    entries[i % len(entries)].graph.replay()
                else:
                    self.step(i)  # N_ROTATE is a multiple of the lane count: a captured chain always replays on the same lane
        for ln in lanes:
            if ln is not main:
                main.wait_stream(ln)
        e1.record()
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def pcie_gbs(self) -> dict:
        """Raw pinned-host <-> device copy bandwidth of THIS box's link (256 MB, best of 3, CUDA events): the ceiling the
        end-to-end number is measured against."""
        n = 64 * 1024 * 1024
        h = torch.empty(n, dtype=torch.float32).pin_memory()
        d = torch.empty(n, dtype=torch.float32, device=self.dev)
        out = {}
        for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
            best = float("inf")
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dst.copy_(src, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name] = n * 4 / best / 1e6
        return out

    def pairs_per_s(self, steps: int, ms_total: float) -> float:
        return self.world * self.wl.batch * steps / (ms_total / 1e3)

    # ---- e2e: CUDAPrefetcher (pinned host batches -> static device slots on a copy stream) + feed_data + D2H of the LQ ----
    def e2e(self, steps: int, warmup: int, u8: bool, lanes: int = 1, readback: str = "f32") -> tuple[float, int, int]:
        from trainner_redux_b200.prefetch import CUDAPrefetcher, CUDAReadback

        wl = self.wl
        batches = []
        for d in self.host:
            if u8:  # extension: 8-bit GT (what the dataset decodes) + the three (B,8) kernel-parameter tables stacked
                b = {"gt": (d["gt"] * 255.0).round().clamp(0, 255).to(torch.uint8).pin_memory(),
                     "kernel_params": torch.stack([torch.as_tensor(p, dtype=torch.float64) for p in d["kernel_params"]]).pin_memory()}
            else:  # the reference's host format: fp32 GT and three (B,21,21) kernels (realesrgan_dataset.py:213-219)
                b = {k: d[k].pin_memory() for k in ("gt", "kernel1", "kernel2", "sinc_kernel")}
            batches.append(b)
        rb = CUDAReadback(self.dev)  # the LQ goes back to pinned host memory on a side stream, every step
        # readback "u8": the LQ lies on the 8-bit lattice, so it crosses PCIe as bytes (CUDAReadback.read(as_u8=True), lossless)
        lq_bytes = {"f32": 4, "u8": 1, "none": 0}[readback] * wl.batch * 3 * (wl.crop // wl.scale) ** 2

        def loader(n):
            for i in range(n):
                yield batches[i % N_ROTATE]

        n_slots = lanes * -(-E2E_SLOTS // lanes)  # a multiple of the lane count: a slot (and its captured chain) stays on one lane
        main = torch.cuda.current_stream()
        lane = [torch.cuda.Stream(self.dev) for _ in range(lanes)] if lanes > 1 else [main]
        n_warm = n_slots * -(-max(warmup, 2 * n_slots) // n_slots)  # (static slots: the chain of each is captured on its second sighting)
        # ONE loader for warm-up and timed steps, as in a training loop: the prefetcher stays one batch ahead, so when the
        # clock starts the first timed batch is already on its way and every timed step uploads the batch of the step
        # after it — K steps, K uploads, K read-backs inside the timed region (the loader yields one batch more than is
        # consumed so that the last step's preload is a real upload too)
        reps = E2E_REPS_SHORT if steps <= E2E_SHORT_STEPS else 1
        pf = CUDAPrefetcher(loader(n_warm + reps * steps + 1), device=self.dev, slots=n_slots)
        state = {"t": 0, "h2d": 0}

        def run(n: int) -> None:
            t = state["t"]
            torch.cuda.set_stream(lane[t % lanes])  # batch t is handed out, degraded and read back on lane t % lanes
            for _ in range(n):
                batch = pf.next()
                state["h2d"] = pf.h2d_bytes
                self.feed.feed_data(batch, plan=self.plan())
                if readback != "none":  # ("none": profiles/e2e_readback_ab.py only)
                    rb.read(self.feed.lq, as_u8=readback == "u8")
                t += 1
                if lanes > 1:
                    torch.cuda.set_stream(lane[t % lanes])
            state["t"] = t
            torch.cuda.set_stream(main)
            rb.wait()
            torch.cuda.synchronize()

        run(n_warm)
        # a short timed region (the driver's K = 20 is under 5 ms of wall clock, host work included) is at the mercy of one
        # scheduling hiccup on the host: it is then measured `reps` times — each EXACTLY `steps` steps between a barrier +
        # synchronise on both sides, max over ranks — and the median is reported, every repetition listed beside it
        vals = []
        for _ in range(reps):
            self.barrier()
            t0 = time.perf_counter()
            run(steps)
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device=self.dev)
            if self.world > 1:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            vals.append(self.world * wl.batch * steps / t.item())
        h2d = state["h2d"]
        params_h2d = 16 + 4 * wl.batch * 6  # the chain's per-step parameter block (chain_graph.ParamBlock)
        self.e2e_reps = [round(v, 1) for v in vals]
        return sorted(vals)[len(vals) // 2], h2d + params_h2d, lq_bytes

    # ---- per-stage GPU durations: every stage re-captured alone (x REP) and replayed, CUDA events ----
    def stage_ms(self) -> dict:
        feed = self.feed
        feed.record_stage_fns = True
        d, p = self.devd[0], self.plan()
        for key in ("noise1", "noise2"):
            p[key] = {k: (v.to(self.dev) if torch.is_tensor(v) else v) for k, v in p[key].items()}
        p["jpeg1"], p["jpeg2"] = p["jpeg1"].to(self.dev), p["jpeg2"].to(self.dev)
        from trainner_redux_b200.transforms import crop_pair

        lq_full = feed.degrade(d["gt"], d["kernel1"], d["kernel2"], d["sinc_kernel"], p)
        feed.stage_fns["crop"] = lambda: crop_pair(d["gt"], lq_full, self.wl.crop, self.wl.scale, 4, 4)
        feed.record_stage_fns = False
        # the launches feed_data's native executor actually makes where it fuses adjacent stages (csrc/chain.cu, row g1):
        # resize + Gaussian noise, and the last DiffJPEG + 8-bit lattice + both crops.  Inputs: the unfused path's intermediates.
        from trainner_redux_b200.stages import StageList

        feed.collect_taps = {}
        feed.degrade(d["gt"], d["kernel1"], d["kernel2"], d["sinc_kernel"], p)
        taps, feed.collect_taps = feed.collect_taps, None
        sc, h0 = self.wl.scale, self.wl.gt

        def fused_resize_noise(src, rs, ns, size=None):
            from trainner_redux_b200 import degradations as D

            hh, ww = src.shape[2:]
            oh, ow = size if size is not None else (round(hh * rs["scale"]), round(ww * rs["scale"]))
            D.pinned_resize_table(self.dev, hh, ww, int(oh), int(ow), D._MODE_ID[rs["mode"]])  # tables outside the timed launches

            def fn():
                sl = StageList(src)
                if size is None:
                    sl.resize(rs["mode"], scale_factor=rs["scale"])
                else:
                    sl.resize(rs["mode"], size=size)
                sl.gaussian_noise(ns["sigma"], ns["gray"], feed.rng.philox)
                return sl.run()
            return fn

        def fused_tail(gt_copy: bool):
            from trainner_redux_b200 import _lib

            def fn():
                x, gt, pch = taps["sinc"], d["gt"], self.wl.crop // sc
                b = x.size(0)
                gt_out = torch.empty((b, 3, self.wl.crop, self.wl.crop), dtype=torch.float32, device=self.dev) if gt_copy else None
                lq_out = torch.empty((b, 3, pch, pch), dtype=torch.float32, device=self.dev)
                _lib.call("otf_diffjpeg_crop_pair_f32", _lib.ptr(x), b, x.size(2), x.size(3), _lib.ptr(p["jpeg2"]), 0.0, 1, 0, 1, _lib.ptr(gt),
                          gt.size(2), gt.size(3), 4, 4, None, pch, sc, _lib.ptr(gt_out), _lib.ptr(lq_out), _lib.stream())
                return lq_out
            return fn

        if p["noise1"]["kind"] == "gaussian" and "blur1" in taps and "blur2" in taps:
            feed.stage_fns["fused resize1+noise1"] = fused_resize_noise(taps["blur1"], p["resize1"], p["noise1"])
            s2 = p["resize2"]["scale"]
            feed.stage_fns["fused resize2+noise2"] = fused_resize_noise(taps["blur2"], p["resize2"], p["noise2"],
                                                                        size=(int(h0 / sc * s2), int(h0 / sc * s2)))
        if "sinc" in taps and p.get("final_order", "resize_first") == "resize_first":
            feed.stage_fns["fused jpeg2+round+lq_crop"] = fused_tail(False)  # what feed_data launches: the GT window stays a view
            feed.stage_fns["fused jpeg2+round+crop"] = fused_tail(True)  # with the dense GT copy (pool / MoA / gt_view=False)
        torch.cuda.synchronize()
        from trainner_redux_b200.degradations import pin_resize_tables

        pin_resize_tables()
        rep, out = 20, {}
        side = torch.cuda.Stream()
        for name, fn in feed.stage_fns.items():
            g = torch.cuda.CUDAGraph()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(rep):
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
                best = min(best, s0.elapsed_time(s1) / rep)
            out[name] = best
            del g
        feed.stage_fns = {}
        return out

    def executed_ops(self) -> dict:
        """FP32-pipe lane operations the blur1 launch EXECUTES per output pixel, from the device-side kernel analysis
        (true radius R, rank-1 and mirror-symmetry flags): dense K^2 FMAs; rank-1 K + K(64+K-1)/64 (the horizontal pass runs
        once per row of the CTA's 64-row tile, through shared memory); folded K(R+1) FMAs + R(4+K-1)/4 adds."""
        from trainner_redux_b200.img_process_util import KernelAnalysis

        k1 = self.devd[0]["kernel1"]
        ka = KernelAnalysis([k1])
        torch.cuda.synchronize()
        kb = k1.size(0)
        s = ka.scratch[: 3 * kb].cpu().tolist()
        ops, kinds = 0.0, {"rank1": 0, "folded": 0, "dense": 0, "identity": 0}
        for r, fl in zip(s[:kb], s[2 * kb:3 * kb]):
            kt = 2 * r + 1
            if r == 0:
                ops, kinds["identity"] = ops + 1, kinds["identity"] + 1
            elif r >= 2 and fl & 1:
                ops, kinds["rank1"] = ops + kt + kt * (64 + kt - 1) / 64, kinds["rank1"] + 1
            elif r >= 2 and fl & 2:
                ops, kinds["folded"] = ops + kt * (r + 1) + r * (4 + kt - 1) / 4, kinds["folded"] + 1
            else:
                ops, kinds["dense"] = ops + kt * kt, kinds["dense"] + 1
        return {"lane_ops_per_output_mean": ops / kb, "kinds": kinds,
                "nonzero_taps_mean": float((k1 != 0).flatten(1).sum(1).float().mean().item())}


def run_b200(args, wl: Workload) -> None:
    import torch.distributed as dist

    from trainner_redux_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("OTF_BENCH_NO_SAMPLER") else None  # spans warm-up, every timed region and the e2e loops
    numa, numa_why = bind_to_gpu_numa_node(local)  # pinned staging buffers should live next to this rank's GPU
    if world > 1:
        # (NCCL's banner and NCCL_DEBUG lines land on stderr: see main)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
    _lib.load()

    arm = Arm(wl, dev, rank, world, dist)
    n_streams = max(1, min(args.streams, N_ROTATE))
    while N_ROTATE % n_streams:
        n_streams -= 1
    for i in range(max(args.warmup, 3 * N_ROTATE)):  # every rotated input set is seen (and its chain captured) before timing
        arm.step(i)
    arm.barrier()
    arm.timed(min(args.steps, 2 * N_ROTATE), n_streams)  # lanes warmed up
    l0 = _lib.launch_count
    ms_total = arm.timed(args.steps, n_streams)
    launches = _lib.launch_count - l0
    # a short region (the driver's K = 20 is 3.5 ms) sees single host hiccups as 10 % swings: measured 5 times then — each
    # region EXACTLY K steps between barrier + synchronise, CUDA events, max over ranks — and the median reported
    regions_ms = [ms_total] + [arm.timed(args.steps, n_streams) for _ in range(E2E_REPS_SHORT - 1 if args.steps <= E2E_SHORT_STEPS else 0)]
    ms_total = sorted(regions_ms)[len(regions_ms) // 2]
    value = arm.pairs_per_s(args.steps, ms_total)
    ms_single = arm.timed(args.steps, 1)
    ms_replay1 = arm.timed(args.steps, 1, replay_only=True) if arm.feed.graphs.entries else float("nan")
    pcie = arm.pcie_gbs()
    ms_replay = arm.timed(args.steps, n_streams, replay_only=True) if arm.feed.graphs.entries else float("nan")
    graphs = arm.feed.graphs
    stage_ms = {} if args.no_stage_timing else arm.stage_ms()

    e2e_f32_value, h2d_f32, d2h = arm.e2e(args.steps, args.warmup, u8=False)
    e2e_one_value, h2d_u8, _ = arm.e2e(args.steps, args.warmup, u8=True)
    e2e_rb32_value, _, _ = arm.e2e(args.steps, args.warmup, u8=True, lanes=n_streams)
    e2e_value, _, d2h_u8 = arm.e2e(args.steps, args.warmup, u8=True, lanes=n_streams, readback="u8")
    e2e_reps = list(arm.e2e_reps)

    extras = {}
    if not args.no_extras and wl.name == "c2" and wl.noise == "gaussian":
        # short runs of the two other workloads SURVEY.md §8d names, so that they are measured by the same command
        for key, w2 in (("poisson", Workload("c2", world, "poisson")), ("c3", Workload("c3", world))):
            a2 = Arm(w2, dev, rank, world, dist)
            for i in range(3 * N_ROTATE):
                a2.step(i)
            n2 = max(20, min(args.steps, 200))
            a2.timed(2 * N_ROTATE, n_streams)
            ms2 = a2.timed(n2, n_streams)
            extras[key] = {"metric": w2.metric, "value": a2.pairs_per_s(n2, ms2), "unit": UNIT, "ms_per_step": ms2 / n2, "steps": n2,
                           "scaling": w2.scaling, "config": w2.config()}
            del a2
            torch.cuda.empty_cache()

    if rank == 0:
        pk = peaks()
        k_ms = stage_ms.get("blur1", float("nan"))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):  # dram__bytes_read+write of this kernel from the committed ncu --set full capture, per workload
            with open(tpath) as f:
                traffic = (json.load(f).get(wl.name) or {}).get("filter2d_blur1_dram_bytes_per_launch")
        blur_bytes = wl.batch * (2 * 3 * wl.gt * wl.gt * 4 + 21 * 21 * 4)
        achieved = blur_bytes / (k_ms * 1e-3) / 1e9
        ex = arm.executed_ops()
        outputs = 3.0 * wl.gt * wl.gt * wl.batch
        fma_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        exec_tflops = 2.0 * ex["lane_ops_per_output_mean"] * outputs / (k_ms * 1e-3) / 1e12
        chain_bytes = wl.algorithmic_bytes_per_pair()
        cpu = parity = torch_cuda = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            keep: dict = {}
            n_cpu = 8 if wl.gt <= 256 else 2
            cval, cms = oracle_chain(wl, wl.batch, n_cpu, 1, threads, keep=keep)
            cpu = {"value": cval, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{n_cpu} batches of {wl.batch} x {wl.gt}^2 GT (+1 warm-up) through the oracle port of the reference chain, torch CPU {threads} threads, {cms:.0f} ms/batch"}
            # parity on this very workload: feed_data with the oracle's finished noise fields injected vs the oracle's LQ
            from trainner_redux_b200.realesrgan_feed import RealESRGANFeed

            pf = RealESRGANFeed(arm.opt, device=dev, use_pool=False)
            pf.feed_data({k: keep["data"][k] for k in ("gt", "kernel1", "kernel2", "sinc_kernel")}, plan=keep["plan"],
                         inject={k: v.to(dev) for k, v in keep["fields"].items()})
            diff = (pf.lq.cpu() - keep["out"][1]).abs()
            parity = {"lq_within_1lsb_frac": (diff <= 1 / 255 + 1e-6).float().mean().item(), "max_lsb": diff.max().item() * 255,
                      "gt_crop_bit_identical": bool(torch.equal(pf.gt.cpu(), keep["out"][0])), "bar": 0.999,
                      "how": "feed_data vs oracle.run_chain_b, same plan / inputs / injected noise fields, full batch"}
            tval, tms = oracle_chain(wl, wl.batch, 5, 2, threads, device=f"cuda:{local}")
            torch_cuda = {"value": tval, "unit": UNIT, "ms_per_step": tms,
                          "what": "the oracle's ATen call sequence (= the reference's primitives) on the CUDA device: PyTorch-eager on this B200"}
        line = {
            "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": wl.config(),  # (the same dict in both arms: the driver compares them)
            "run": {"launch": f"RealESRGANFeed.feed_data, a fresh draw_plan per step, captured chains replayed ({graphs.captures} captures, "
                              f"{graphs.hits} replays so far), {n_streams} calls in flight on {n_streams} streams",
                    "gt_crop": "the GT half of the pair is the reference's view of the GT batch (transforms.py:124-129: a slice, no copy); the LQ crop is dense",
                    "numa_node_rank0": numa, "numa_note": numa_why},
            "timed_regions_ms": [round(m, 4) for m in regions_ms],
            "value_feed_data": arm.pairs_per_s(args.steps, ms_single), "ms_per_step_feed_data": ms_single / args.steps,
            "value_graph_replay": arm.pairs_per_s(args.steps, ms_replay), "ms_per_step_graph_replay": ms_replay / args.steps,
            "value_graph_replay_one_stream": arm.pairs_per_s(args.steps, ms_replay1),
            "pcie_measured_gbs": pcie,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h_u8,
                    "h2d_gbs_per_rank": h2d_u8 * e2e_value / world / wl.batch / 1e9,
                    "timed_regions": e2e_reps, "estimator": f"median of {len(e2e_reps)} timed regions of exactly {args.steps} steps each" if len(e2e_reps) > 1 else f"one timed region of {args.steps} steps",
                    "feed": "pinned host uint8 GT + (3,B,8) kernel-parameter table (the dataset's decoded format; /255 and kernel "
                            f"synthesis on the device) -> CUDAPrefetcher -> feed_data -> CUDAReadback of the whole LQ batch as bytes "
                            f"(it lies on the 8-bit lattice: lossless, u8 / 255 restores it bit for bit), {n_streams} batches in flight"},
            "e2e_readback_f32": {"value": e2e_rb32_value, "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                                 "note": "same loop with the LQ read back as fp32 (4x the D2H bytes): equal on one GPU, slower when eight "
                                         "ranks share the host path (profiles/r02_e2e_readback_n8.json)"},
            "e2e_one_stream": {"value": e2e_one_value, "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h},
            "e2e_f32": {"value": e2e_f32_value, "unit": UNIT, "h2d_bytes_per_step": h2d_f32, "d2h_bytes_per_step": d2h,
                        "h2d_gbs_per_rank": h2d_f32 * e2e_f32_value / world / wl.batch / 1e9,
                        "feed": "the reference's host format: pinned fp32 GT + three (B,21,21) kernels (PCIe-bound: 4x the bytes)"},
            "gpu_launches": launches,
            "clocks": sampler.stop() if sampler else None,
            "roofline": {"kernel": f"filter2d_kernel (blur1, {wl.batch}x3x{wl.gt}x{wl.gt}, per-sample kernels zero-padded to 21x21: default kernel_list mix, sizes 7..21)",
                         "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                         "traffic": traffic, "peak_source": pk["source"], "ms_per_launch": k_ms,
                         "fma": {"executed_tflops": exec_tflops, "frac_of_fma_peak": exec_tflops / fma_peak, "peak_tflops": fma_peak,
                                 **ex, "note": "filter2d is FP32-FMA bound above K~9 (SURVEY.md H1); executed FP32-pipe operations, not nominal K^2 taps"}},
            "chain": {"algorithmic_bytes_per_pair": chain_bytes, "achieved_gbs": chain_bytes * value / world / 1e9,
                      "frac_of_hbm_peak": chain_bytes * value / world / 1e9 / pk["hbm_gbs"], "stage_ms": stage_ms,
                      "stage_frac_of_hbm_peak": {k: round(wl.batch * nb / (stage_ms[k] * 1e-3) / 1e9 / pk["hbm_gbs"], 4)
                                                 for k, nb in wl.stage_bytes_per_pair().items() if stage_ms.get(k)},
                      "kernels_per_step": launches / args.steps},
            "cpu_baseline": cpu, "reference_torch_cuda": torch_cuda, "parity": parity, **extras,
        }
        emit(line)
    elif sampler:
        sampler.stop()
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
    ap.add_argument("--no-extras", action="store_true", help="skip the short Poisson / c3 runs of the default line")
    ap.add_argument("--streams", type=int, default=4, help="feed_data calls in flight during the timed region")
    ap.add_argument("--no-stage-timing", action="store_true", help="skip the per-stage re-capture pass (for ncu launch lists)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record.  Libraries that write to file descriptor 1 on their own (NCCL prints
    # "NCCL version ..." and, under NCCL_DEBUG, its communicator lines there) are not silenced: for the whole run fd 1 is an
    # alias of stderr, where those lines stay visible to whoever collects them, and the record goes to the saved stdout.
    global _RECORD_OUT
    sys.stdout.flush()
    _RECORD_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = Workload(args.workload, int(os.environ.get("WORLD_SIZE", "1")), args.noise)
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
