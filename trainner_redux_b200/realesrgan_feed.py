"""Drop-in for the OTF half of traiNNer/models/realesrgan_model.py: ``feed_data`` (:455-650) and
the training-pair pool ``_dequeue_and_enqueue`` (:403-453).

``RealESRGANFeed`` owns what ``RealESRGANModel`` owns for this path — ``self.gt`` / ``self.lq``, the
``jpeger``, the ``usm_sharpener``, the pool — and nothing else (networks, losses and optimisers are
out of scope).  A maintainer swaps it in by making ``RealESRGANModel.feed_data`` delegate to it
(INTEGRATION.md).

Two stage orders are implemented:
  * ``order="classic"`` — the second-order Real-ESRGAN chain the option schema still describes
    (traiNNer/utils/redux_options.py:720-851): [USM] blur1, resize1, noise1, jpeg1, blur2, resize2,
    noise2, {resize3+sinc, jpeg2} in random order, clamp/round, crop, pool.
  * ``order="fork"`` — the as-shipped order of this fork (realesrgan_model.py:525-526, :564-574,
    :616-627) minus the probability-0 extras and the PIL codec round trip: [blur1], resize3, sinc,
    clamp/round, crop, pool.

All host-side decisions are drawn in Python in the reference's order by
``draw_plan``; the bulk random fields come from the kernels' Philox streams.  Nothing between the
H2D copies and the finished pair synchronises with the host.
"""

from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass
from typing import Any, Callable, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib
from . import degradations as D
from .chain_graph import ChainGraph, ChainGraphCache, ParamBlock, ParamCollector, fill_recipe, plan_signature
from .diffjpeg import DiffJPEG
from .img_process_util import USMSharp
from .stages import StageList
from .transforms import crop_pair


# --------------------------------------------------------------------------- options ----
@dataclass
class OTFOptions:
    """The fields of ``ReduxOptions`` that govern the path, with the reference's names and
    defaults (traiNNer/utils/redux_options.py:720-901).  Any object exposing these attributes
    (e.g. a real ``ReduxOptions``) can be passed instead."""

    scale: int = 4
    gt_size: int = 256  # opt.datasets["train"].gt_size (realesrgan_model.py:619)
    lq_usm: bool = False
    lq_usm_radius_range: tuple[int, int] = (1, 25)
    blur_prob: float = 0
    resize_prob: Sequence[float] = (0.2, 0.7, 0.1)
    resize_mode_list: Sequence[str] = ("bilinear", "bicubic", "nearest-exact", "lanczos")
    resize_mode_prob: Sequence[float] = (0.25, 0.25, 0.25, 0.25)
    resize_range: tuple[float, float] = (0.4, 1.5)
    gaussian_noise_prob: float = 0
    noise_range: tuple[float, float] = (0, 0)
    poisson_scale_range: tuple[float, float] = (0, 0)
    gray_noise_prob: float = 0
    jpeg_prob: float = 1
    jpeg_range: tuple[float, float] = (75, 95)
    blur_prob2: float = 0
    resize_prob2: Sequence[float] = (0.3, 0.4, 0.3)
    resize_mode_list2: Sequence[str] = ("bilinear", "bicubic", "nearest-exact", "lanczos")
    resize_mode_prob2: Sequence[float] = (0.25, 0.25, 0.25, 0.25)
    resize_range2: tuple[float, float] = (0.6, 1.2)
    gaussian_noise_prob2: float = 0
    noise_range2: tuple[float, float] = (0, 0)
    poisson_scale_range2: tuple[float, float] = (0, 0)
    gray_noise_prob2: float = 0
    jpeg_prob2: float = 1
    jpeg_range2: Sequence[float] = (75, 95)
    resize_mode_list3: Sequence[str] = ("bilinear", "bicubic", "nearest-exact", "lanczos")
    resize_mode_prob3: Sequence[float] = (0.25, 0.25, 0.25, 0.25)
    queue_size: int = 120
    p_clean: float = 0
    # MoA batch augment on the finished pair (redux_options.py:295-322 `train.use_moa` ...; realesrgan_model.py:649-650)
    use_moa: bool = False
    # the schema's own defaults (4 names, 5 weights) fail batch_aug's length check in the reference too: set both
    moa_augs: Sequence[str] = ("none", "mixup", "cutmix", "resizemix")
    moa_probs: Sequence[float] = (0.4, 0.084, 0.084, 0.084, 0.348)
    # the fork's extra stages (redux_options.py:525-700, :854-894; all off by default, as in the schema)
    lens_distort_prob: float = 0
    lens_distort_strength_range: tuple[float, float] = (-0.3, 0.3)
    chromatic_aberration_prob: float = 0
    motion_blur_prob: float = 0
    motion_blur_kernel_size: tuple[int, int] = (5, 15)
    motion_blur_angle_range: tuple[float, float] = (0, 360)
    demosaic_prob: float = 0
    sensor_noise_prob: float = 0
    sensor_noise_std_range: tuple[float, float] = (0.01, 0.1)
    rolling_shutter_prob: float = 0
    rolling_shutter_strength_range: tuple[float, float] = (-0.1, 0.1)
    exposure_prob: float = 0
    exposure_factor_range: tuple[float, float] = (0.5, 2.0)
    color_temp_prob: float = 0
    color_temp_shift_range: tuple[float, float] = (-0.2, 0.2)
    oversharpen_prob: float = 0
    oversharpen_strength: tuple[float, float] = (1.0, 2.0)
    aliasing_prob: float = 0
    aliasing_scale_range: tuple[float, float] = (0.6, 0.9)
    compression_formats: Sequence[str] = ("jpeg", "webp", "avif", "heif")
    compression_weights: Sequence[float] = (0.60, 0.25, 0.10, 0.05)
    compression_jpeg_range: tuple[float, float] = (45, 95)
    compression_webp_range: tuple[float, float] = (45, 95)
    compression_avif_range: tuple[float, float] = (35, 90)
    compression_heif_range: tuple[float, float] = (40, 90)
    recompression_prob: float = 0
    recompression_formats: Sequence[str] = ("jpeg", "webp", "avif", "heif")
    recompression_weights: Sequence[float] = (0.50, 0.35, 0.10, 0.05)
    editing_prob: float = 0
    editing_exposure_prob: float = 0
    editing_exposure_range: tuple[float, float] = (0.9, 1.1)
    editing_oversharpen_prob: float = 0
    editing_oversharpen_strength: tuple[float, float] = (1.0, 1.3)
    fork_compression: bool = True  # order="fork": run the unified compression stage (realesrgan_model.py:581)
    codec_fallback: str = "passthrough"  # webp/avif/heif rounds: "passthrough" (reference without the plugin) or "jpeg" (DiffJPEG round)
    # not in the reference schema: which composition to run, and the upstream 50/50 final-order coin
    order: str = "classic"
    final_jpeg_first_prob: float = 0.5
    noise_enabled: bool = True  # classic order draws a noise stage (Gaussian else Poisson) as upstream does


class HostRNG:
    """The reference's three host generators: a numpy ``Generator`` seeded WITHOUT the rank offset
    (traiNNer/utils/rng.py:19-25), and Python ``random`` + the torch CPU generator seeded with
    ``manual_seed + rank`` (train.py:283 -> utils/misc.py:19-22)."""

    def __init__(self, manual_seed: int = 0, rank: int = 0) -> None:
        self.np = np.random.default_rng(manual_seed)
        self.py = random.Random(manual_seed + rank)
        self.torch = torch.Generator().manual_seed(manual_seed + rank)
        self.philox = D.PhiloxState(seed=(manual_seed + rank) * 0x9E3779B97F4A7C15 + 0xB200)


def _opt(o: Any, name: str, default: Any = None) -> Any:
    return getattr(o, name, default)


def _draw_resize(rng: HostRNG, probs, rrange, modes, mode_probs) -> dict:
    kind = rng.py.choices(["up", "down", "keep"], probs)[0]
    if kind == "up":
        s = rng.np.uniform(1, rrange[1])
    elif kind == "down":
        s = rng.np.uniform(rrange[0], 1)
    else:
        s = 1.0
    mode = rng.py.choices(list(modes), weights=list(mode_probs))[0]
    return {"scale": float(s), "mode": mode}


def _draw_noise(rng: HostRNG, b: int, gaussian_prob, noise_range, poisson_range, gray_prob) -> dict:
    # gate first, then the primitive's own draws: rand(B) value, rand(B) gray (degradations.py:673-679)
    gaussian = rng.np.uniform() < gaussian_prob
    rr = noise_range if gaussian else poisson_range
    # (in-place forms of `rand * (hi - lo) + lo` and `(rand < p).float()`: same values, a third of the dispatches)
    val = torch.rand(b, generator=rng.torch).mul_(rr[1] - rr[0]).add_(rr[0])
    gray = torch.rand(b, generator=rng.torch).lt_(gray_prob)
    if gaussian:
        return {"kind": "gaussian", "sigma": val, "gray": gray}
    return {"kind": "poisson", "scale": val, "gray": gray}


def _draw_quality(rng: HostRNG, b: int, qrange) -> Tensor:
    # out.new_zeros(B).uniform_(*jpeg_range)
    return torch.empty(b).uniform_(float(qrange[0]), float(qrange[1]), generator=rng.torch)


EXTRA_KEYS = ("lens", "chroma", "motion", "demosaic", "sensor", "shutter", "exposure", "color_temp", "oversharpen", "aliasing",
              "compression", "editing_exposure")


def _draw_fork(opt: Any, rng: HostRNG) -> dict:
    """Host draws of the fork's order (A), realesrgan_model.py:512-611 with the `hasattr` guards of
    paragon_otf_degradations.py: a stage whose option fields exist draws its gate even at probability 0, and its
    parameters only when the gate passes.  Pinned against the reference by oracle/make_paragon_goldens.py."""
    u = rng.np.uniform
    has = lambda *names: all(hasattr(opt, n) for n in names)  # noqa: E731
    p: dict[str, Any] = {}
    if has("lens_distort_prob", "lens_distort_strength_range") and u() < opt.lens_distort_prob:
        p["lens"] = float(u(*opt.lens_distort_strength_range))
    if has("chromatic_aberration_prob") and u() < opt.chromatic_aberration_prob:
        p["chroma"] = True
    if has("motion_blur_prob") and u() < opt.motion_blur_prob:
        ks = rng.py.randint(opt.motion_blur_kernel_size[0], opt.motion_blur_kernel_size[1])
        p["motion"] = (ks, float(u(*opt.motion_blur_angle_range)))
    p["blur1"] = bool(u() < _opt(opt, "blur_prob", 0))  # :525
    if has("demosaic_prob") and u() < opt.demosaic_prob:
        p["demosaic"] = True
    if has("sensor_noise_prob", "sensor_noise_std_range") and u() < opt.sensor_noise_prob:
        p["sensor"] = float(u(*opt.sensor_noise_std_range))
    if has("rolling_shutter_prob", "rolling_shutter_strength_range") and u() < opt.rolling_shutter_prob:
        p["shutter"] = float(u(*opt.rolling_shutter_strength_range))
    if has("exposure_prob", "exposure_factor_range") and u() < opt.exposure_prob:
        p["exposure"] = float(u(*opt.exposure_factor_range))
    if has("color_temp_prob", "color_temp_shift_range") and u() < opt.color_temp_prob:
        p["color_temp"] = float(u(*opt.color_temp_shift_range))
    if has("oversharpen_prob", "oversharpen_strength") and u() < opt.oversharpen_prob:
        p["oversharpen"] = float(u(*opt.oversharpen_strength))
    if has("aliasing_prob", "aliasing_scale_range") and u() < opt.aliasing_prob:
        p["aliasing"] = float(u(*opt.aliasing_scale_range))
    p["resize3_mode"] = rng.py.choices(list(opt.resize_mode_list3), weights=list(opt.resize_mode_prob3))[0]  # :564
    if _opt(opt, "fork_compression", True) and has("compression_formats"):  # :581 -> paragon :39-87
        def one(formats, weights):
            fmt = str(rng.np.choice(list(formats), p=list(weights)))
            rr = getattr(opt, f"compression_{fmt}_range", None)
            return fmt, (float(u(rr[0], rr[1])) if rr is not None else None)

        comp = [one(opt.compression_formats, opt.compression_weights)]
        if u() < opt.recompression_prob:
            comp.append(one(opt.recompression_formats, opt.recompression_weights))
        p["compression"] = comp
    if u() < _opt(opt, "editing_prob", 0):  # :589-611
        if has("editing_exposure_prob") and u() < opt.editing_exposure_prob:
            p["editing_exposure"] = float(u(*_opt(opt, "editing_exposure_range", (0.9, 1.1))))
        if has("editing_oversharpen_prob"):
            u()  # the gate is drawn; the branch applies nothing (:606-611)
    return p


def resolve_gt_size(opt: Any, default: int | None = None) -> int | None:
    """``gt_size`` as the reference reads it: ``opt.datasets["train"].gt_size`` (realesrgan_model.py:619); a flat
    ``opt.gt_size`` (OTFOptions) wins when present."""
    v = getattr(opt, "gt_size", None)
    if v is None:
        ds = getattr(opt, "datasets", None)
        train = ds.get("train") if isinstance(ds, dict) else getattr(ds, "train", None)
        v = getattr(train, "gt_size", None) if train is not None else None
        if v is None and isinstance(train, dict):
            v = train.get("gt_size")
    return default if v is None else int(v)


def resolve_order(opt: Any) -> str:
    """Which composition runs: ``opt.order`` when the options object has one (OTFOptions: "classic" by default); an
    object without the field — a real ``ReduxOptions`` — gets the order this fork's ``feed_data`` executes ("fork")."""
    return getattr(opt, "order", None) or "fork"


def draw_plan(opt: Any, b: int, ori_h: int, ori_w: int, rng: HostRNG, gt_size: int | None = None, order: str | None = None) -> dict:
    """Host-side decisions for one ``feed_data`` call, in the order of SURVEY.md appendix A.
    The result is a plain dict (the format ``oracle.otf_oracle.run_chain_b`` consumes).  ``gt_size`` / ``order``
    override what the options object says (a frozen ``ReduxOptions`` struct cannot be given new attributes)."""
    scale = _opt(opt, "scale", 4)
    plan: dict[str, Any] = {"scale": scale, "gt_size": gt_size or resolve_gt_size(opt, ori_h), "order": order or resolve_order(opt)}
    # the gate is drawn whenever the option exists — ReduxOptions always defines p_clean (default 0), so the reference
    # consumes one numpy uniform here on every call (realesrgan_model.py:487-489, SURVEY.md appendix A draw #1)
    if hasattr(opt, "p_clean") and rng.np.uniform() < opt.p_clean:  # realesrgan_model.py:487-503
        plan["clean"] = True
    elif plan["order"] == "fork":
        plan.update(_draw_fork(opt, rng))
    else:
        if _opt(opt, "lq_usm", False):
            lo, hi = opt.lq_usm_radius_range
            plan["usm"] = {"radius": rng.py.randint(lo, hi), "weight": 0.5, "threshold": 10}
        plan["blur1"] = bool(rng.np.uniform() < _opt(opt, "blur_prob", 0))
        plan["resize1"] = _draw_resize(rng, opt.resize_prob, opt.resize_range, opt.resize_mode_list, opt.resize_mode_prob)
        if _opt(opt, "noise_enabled", True):
            plan["noise1"] = _draw_noise(rng, b, opt.gaussian_noise_prob, opt.noise_range, opt.poisson_scale_range, opt.gray_noise_prob)
        plan["jpeg1"] = _draw_quality(rng, b, opt.jpeg_range) if rng.np.uniform() < opt.jpeg_prob else None
        plan["blur2"] = bool(rng.np.uniform() < _opt(opt, "blur_prob2", 0))
        plan["resize2"] = _draw_resize(rng, opt.resize_prob2, opt.resize_range2, opt.resize_mode_list2, opt.resize_mode_prob2)
        if _opt(opt, "noise_enabled", True):
            plan["noise2"] = _draw_noise(rng, b, opt.gaussian_noise_prob2, opt.noise_range2, opt.poisson_scale_range2, opt.gray_noise_prob2)
        plan["final_order"] = "resize_first" if rng.np.uniform() >= _opt(opt, "final_jpeg_first_prob", 0.5) else "jpeg_first"
        plan["resize3_mode"] = rng.py.choices(list(opt.resize_mode_list3), weights=list(opt.resize_mode_prob3))[0]
        plan["jpeg2"] = _draw_quality(rng, b, opt.jpeg_range2) if rng.np.uniform() < opt.jpeg_prob2 else None
    # paired_random_crop: two random.randint draws (transforms.py:119-120)
    p = plan["gt_size"] // scale
    h_lq, w_lq = ori_h // scale, ori_w // scale
    if h_lq < p or w_lq < p:
        raise ValueError(f"LQ ({h_lq}, {w_lq}) is smaller than patch size ({p}, {p}). Please remove None.")
    plan["crop"] = (rng.py.randint(0, h_lq - p), rng.py.randint(0, w_lq - p))
    return plan


# --------------------------------------------------------------------- small kernels ----
def clamp_round(x: Tensor) -> Tensor:
    """``clamp(round(x*255),0,255)/255`` — realesrgan_model.py:616."""
    x = _lib.dense_f32(x)
    out = torch.empty_like(x)
    _lib.call("otf_clamp_round_f32", _lib.ptr(x), x.numel(), _lib.ptr(out), _lib.stream())
    return out


# ------------------------------------------------------------------------- pair pool ----
class SlotMover:
    """Moves pool slots with the library's gather/scatter kernels."""

    def gather(self, src: Tensor, idx: Sequence[int]) -> Tensor:
        n = len(idx)
        out = torch.empty((n, *src.shape[1:]), dtype=src.dtype, device=src.device)
        arr = (C.c_int32 * n)(*idx)
        _lib.call("otf_gather_slots_f32", _lib.ptr(src), arr, n, src[0].numel(), _lib.ptr(out), _lib.stream())
        return out

    def scatter(self, dst: Tensor, idx: Sequence[int], src: Tensor) -> None:
        n = len(idx)
        arr = (C.c_int32 * n)(*idx)
        _lib.call("otf_scatter_slots_f32", _lib.ptr(_lib.dense_f32(src)), arr, n, dst[0].numel(), _lib.ptr(dst), _lib.stream())

    def exchange(self, queue_lr: Tensor, queue_gt: Tensor, idx: Sequence[int], lq: Tensor, gt: Tensor,
                 dequeue: bool) -> tuple[Tensor, Tensor]:
        """One pool step in one launch: enqueue ``lq`` / ``gt`` into the listed slots and, when ``dequeue``, hand out
        what they held (otherwise the new pair is passed through, as the reference does while the pool fills)."""
        n = len(idx)
        arr = (C.c_int32 * n)(*idx)
        lq, gt = _lib.dense_f32(lq), _lib.dense_f32(gt)
        lq_out = torch.empty_like(lq) if dequeue else None
        gt_out = torch.empty_like(gt) if dequeue else None
        _lib.call("otf_pool_exchange_f32", _lib.ptr(queue_lr), _lib.ptr(queue_gt), arr, n, lq[0].numel(), gt[0].numel(), _lib.ptr(lq),
                  _lib.ptr(gt), _lib.ptr(lq_out), _lib.ptr(gt_out), _lib.stream())
        return (lq_out, gt_out) if dequeue else (lq, gt)


class PairPool:
    """Training pair pool (realesrgan_model.py:403-453) without the full-queue gather.

    The reference shuffles by materialising ``queue[idx]`` for both queues every iteration
    (~100 MB of traffic at queue_size 120).  Here the queues never move: a host-side table maps
    logical positions to physical slots, ``randperm`` permutes the table, and only the ``b``
    dequeued / enqueued slots are touched.  Given the same ``randperm`` results the returned
    batches are bit-identical to the reference's (tests/test_pool_cpu.py)."""

    def __init__(self, queue_size: int, mover: Any | None = None, randperm: Callable[[int], Tensor] | None = None) -> None:
        self.queue_size = queue_size
        self.mover = mover or SlotMover()
        self.randperm = randperm or (lambda n: torch.randperm(n))
        self.queue_lr: Tensor | None = None
        self.queue_gt: Tensor | None = None
        self.queue_ptr = 0
        self.slot_of = list(range(queue_size))  # logical position -> physical slot

    def step(self, lq: Tensor, gt: Tensor) -> tuple[Tensor, Tensor]:
        b = lq.size(0)
        if self.queue_lr is None:
            assert self.queue_size % b == 0, f"queue size {self.queue_size} should be divisible by batch size {b}"
            self.queue_lr = torch.zeros((self.queue_size, *lq.shape[1:]), dtype=lq.dtype, device=lq.device)
            self.queue_gt = torch.zeros((self.queue_size, *gt.shape[1:]), dtype=gt.dtype, device=gt.device)
            self.queue_ptr = 0
        assert self.queue_gt is not None
        if self.queue_ptr == self.queue_size:  # the pool is full: shuffle, dequeue b, enqueue b
            idx = self.randperm(self.queue_size).tolist()
            self.slot_of = [self.slot_of[i] for i in idx]
            slots = self.slot_of[:b]
            if hasattr(self.mover, "exchange"):
                return self.mover.exchange(self.queue_lr, self.queue_gt, slots, lq, gt, True)
            lq_out = self.mover.gather(self.queue_lr, slots)
            gt_out = self.mover.gather(self.queue_gt, slots)
            self.mover.scatter(self.queue_lr, slots, lq)
            self.mover.scatter(self.queue_gt, slots, gt)
            return lq_out, gt_out
        slots = self.slot_of[self.queue_ptr : self.queue_ptr + b]
        self.queue_ptr += b
        if hasattr(self.mover, "exchange"):
            return self.mover.exchange(self.queue_lr, self.queue_gt, slots, lq, gt, False)
        self.mover.scatter(self.queue_lr, slots, lq)
        self.mover.scatter(self.queue_gt, slots, gt)
        return lq, gt


# ------------------------------------------------------------------------------ feed ----
class RealESRGANFeed:
    """``feed_data`` for OTF training: turns ``{gt, kernel1, kernel2, sinc_kernel}`` into
    ``self.gt`` / ``self.lq`` on the device."""

    def __init__(self, opt: Any, device: torch.device | str = "cuda", manual_seed: int = 0, rank: int = 0,
                 use_pool: bool = True, gt_size: int | None = None, order: str | None = None) -> None:
        """``opt``: an ``OTFOptions`` or the reference's ``ReduxOptions`` (read-only: nothing is ever set on it).
        ``gt_size`` defaults to ``opt.gt_size`` or ``opt.datasets["train"].gt_size`` (realesrgan_model.py:619);
        ``order`` to ``opt.order`` or, for an options object without that field, "fork" — the order the reference's
        own ``feed_data`` runs.  MoA fields are read from ``opt.train`` when ``opt`` has one (base_model.py:875-876)."""
        self.opt = opt
        self.gt_size = gt_size or resolve_gt_size(opt)
        self.order = order or resolve_order(opt)
        if self.order not in ("classic", "fork"):
            raise ValueError(f"order must be 'classic' or 'fork', got {self.order!r}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RealESRGANFeed runs on a CUDA device only (no CPU fallback)")
        _lib.load()
        self.is_train = True
        self.rng = HostRNG(manual_seed, rank)
        self.jpeger = DiffJPEG(differentiable=False)  # realesrgan_model.py:81-83
        with torch.cuda.device(self.device):
            D.poisson_tables(self.device, wait=True)  # the Poisson sampler's universal CDF tables: built once per device, outside any capture
        self._usm: dict[int, USMSharp] = {}
        self.queue_size = _opt(opt, "queue_size", 120)
        self.pool = PairPool(self.queue_size, randperm=lambda n: torch.randperm(n, generator=self.rng.torch)) if use_pool else None
        self.batch_augment = None  # base_model.py:875-876: BatchAugment(self.opt.scale, self.opt.train)
        train_opt = getattr(opt, "train", None)
        moa_opt = train_opt if train_opt is not None and hasattr(train_opt, "use_moa") else opt
        if _opt(moa_opt, "use_moa", False):
            from .batchaug import BatchAugment

            self.batch_augment = BatchAugment(_opt(opt, "scale", 4), moa_opt, rng=self.rng)
        self.gt: Tensor | None = None
        self.lq: Tensor | None = None
        self.last_plan: dict | None = None
        # optional per-stage CUDA-event timing (bench.py): name -> [(start, stop), ...]
        self.time_stages = False
        self.stage_times: dict[str, list] = {}
        self.record_stage_fns = False
        self.stage_fns: dict[str, Callable[[], Tensor]] = {}
        self.collect_taps: dict[str, Tensor] | None = None
        # launch the whole chain from ONE library call (stages.py / otf_run_stages_f32) instead of one Python call
        # per stage; same kernels, same arguments, bit-identical results.  The per-stage path stays for the stage hooks.
        self.native_chain = True
        # Shape-stable schedules replay a captured chain (chain_graph.py): a plan signature seen twice is captured into
        # a CUDA graph whose per-step numbers (per-sample sigma / scale / gray / quality, crop offsets, Philox position)
        # live in a small device block, so a step costs the host one plan draw, one tiny H2D and one graph launch.
        # NOTE: a replayed chain returns ``self.gt`` / ``self.lq`` in buffers the NEXT replay of the same chain
        # overwrites (the reference returns fresh tensors); a caller that keeps batches across steps sets
        # ``use_graphs = False`` or clones.  The pool (``use_pool``) copies what it keeps, so it is unaffected.
        self.use_graphs = True
        self.graphs = ChainGraphCache(capacity=16, credits=8.0)
        # The GT half of the pair as a VIEW of the GT batch, as the reference's paired_random_crop returns it
        # (transforms.py:124-129: a slice; only the LQ is made contiguous, realesrgan_model.py:627) — no GT bytes move.
        # The view aliases the batch feed_data was given (a prefetcher's static slot is rewritten `slots` calls
        # later).  Whenever something downstream needs a dense tensor anyway (pair pool, MoA) the window is copied by
        # the chain's last launch instead; ``gt_view = False`` forces that copy.
        self.gt_view = True
        self._synth_out: dict[tuple, Tensor] = {}  # kernel-synthesis outputs per upload slot (feed_data, `kernel_params`)
        self._gt_f32: dict[tuple, Tensor] = {}  # normalised fp32 GT per uint8 upload slot

    def _synth_stacked(self, kp: Tensor) -> tuple[Tensor, Tensor, Tensor]:
        """(3, B, 8) float64 parameter tables -> three (B, 21, 21) fp32 kernels from one launch.  One output buffer per
        table address (a prefetcher's static slot), so the kernels' addresses repeat and captured chains are replayed."""
        if not (kp.is_cuda and kp.dtype == torch.float64 and kp.is_contiguous()):
            kp = kp.to(self.device, dtype=torch.float64, non_blocking=True).contiguous()
            hit = None
        else:
            hit = self._synth_out.get((kp.data_ptr(), 3))
        if kp.size(2) != 8:
            raise ValueError("kernel_params must have shape (3, B, 8)")
        if hit is None or hit[0].size(0) != kp.size(1):
            buf = torch.empty((3, kp.size(1), 21, 21), dtype=torch.float32, device=self.device)
            hit = (buf[0], buf[1], buf[2], buf)
            if len(self._synth_out) > 64:
                self._synth_out.clear()
            self._synth_out[(kp.data_ptr(), 3)] = hit
        _lib.call("otf_synth_kernels_f32", _lib.ptr(kp), 3 * kp.size(1), _lib.ptr(hit[3]), _lib.stream())
        return hit[0], hit[1], hit[2]

    def _timed(self, name: str, fn: Callable[[], Tensor]) -> Tensor:
        if self.record_stage_fns:
            self.stage_fns[name] = fn  # closure over this call's inputs: bench.py re-launches it in a graph
        if self.collect_taps is not None:  # parity localisation (profiles/parity_localise.py): keep every intermediate
            self.collect_taps[name] = out = fn()
            return out
        if not self.time_stages:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.stage_times.setdefault(name, []).append((e0, e1))
        return out

    # -- the chain, described once ----------------------------------------------------------
    def _record(self, sl: Any, kernel1: Tensor, kernel2: Tensor, sinc_kernel: Tensor, plan: dict,
                inject: dict | None = None) -> Any:
        """THE description of the chain: which stages run, in which order, with which arguments and Philox offsets.
        ``sl`` is whatever consumes it — a ``StageList`` (launch records for the native executor), a
        ``ParamCollector`` (the per-step numbers of a captured chain) or a ``StageByStage`` (runs each stage at once
        through the public per-primitive API, behind the stage hooks)."""
        inject = inject or {}
        ori_h, ori_w, sc = sl.h, sl.w, plan["scale"]
        if plan.get("clean"):
            sl.at("round").clamp_round()
            return sl
        joint = kernel1.shape == kernel2.shape == sinc_kernel.shape and kernel1.size(-1) <= 21 and kernel1.size(0) == sl.b
        fork = plan.get("order") == "fork"
        if joint:  # ONE launch analyses all three kernel tensors (img_process_util.KernelAnalysis on the per-stage path)
            sl.analyse([kernel1, kernel2, sinc_kernel])
        an = (lambda i: i) if joint else (lambda i: None)

        def noise(st: dict, key: str) -> None:
            if inject.get(f"{key}_field") is not None:  # the finished noise field of the reference (parity tests)
                sl.at(key).noise_field(inject[f"{key}_field"])
            elif st["kind"] == "gaussian":
                sl.at(key).gaussian_noise(st["sigma"], st["gray"], self.rng.philox, noise=inject.get(f"{key}_color"),
                                  noise_gray=inject.get(f"{key}_gray"))
            else:
                sl.at(key).poisson_noise(st["scale"], st["gray"], self.rng.philox, counts=inject.get(f"{key}_counts_color"),
                                 counts_gray=inject.get(f"{key}_counts_gray"))

        def tail(jpeg: Tensor | None) -> None:  # resize3, sinc, then the 8-bit lattice (fused into the last JPEG)
            sl.at("resize3").resize(plan["resize3_mode"], size=(ori_h // sc, ori_w // sc))
            sl.at("sinc").filter2d(sinc_kernel, an(2))
            if jpeg is not None:
                sl.at("jpeg2+round").jpeg(jpeg, round8=True)
            else:
                sl.at("round").clamp_round()

        if fork:  # realesrgan_model.py:512-616, stage by stage as degrade() runs it
            from . import paragon_otf as PO

            if "lens" in plan:
                sl.at("lens").warp(_lib.WARP_LENS, plan["lens"])
            if plan.get("chroma") and sl.c == 3:
                sl.at("chroma").warp(_lib.WARP_CHROMA, 0.0)
            if "motion" in plan:
                sl.at("motion").taps_zero(PO.motion_blur_kernel(*plan["motion"]))
            if plan.get("blur1"):
                sl.at("blur1").filter2d(kernel1, an(0))
            if plan.get("demosaic"):
                sl.at("demosaic").demosaic()
            if "sensor" in plan:
                sl.at("sensor").sensor_noise(plan["sensor"], self.rng.philox, noise=inject.get("sensor_noise"))
            if "shutter" in plan:
                sl.at("shutter").warp(_lib.WARP_SHUTTER, plan["shutter"] * sl.h / sl.w)
            if "exposure" in plan:
                sl.at("exposure").gain((plan["exposure"],) * 3)
            if "color_temp" in plan and sl.c == 3:
                sl.at("color_temp").gain(PO.color_temperature_gains(plan["color_temp"]))
            if "oversharpen" in plan:
                sl.at("oversharpen").taps_zero(PO._BOX5, _lib.TAPS_OVERSHARPEN, plan["oversharpen"])
            if "aliasing" in plan:
                h, w = sl.h, sl.w
                sl.at("aliasing_down").resize_raw(_lib.RESIZE_NEAREST, int(h * plan["aliasing"]), int(w * plan["aliasing"]), False)
                sl.at("aliasing").resize_raw(_lib.RESIZE_NEAREST, h, w, False)
            sl.at("resize3").resize(plan["resize3_mode"], size=(ori_h // sc, ori_w // sc))
            sl.at("sinc").filter2d(sinc_kernel, an(2))
            if plan.get("jpeg") is not None:  # per-sample qualities through DiffJPEG (this repo's earlier routing)
                sl.at("jpeg+round").jpeg(plan["jpeg"], round8=True)
                return sl
            fallback = _opt(self.opt, "codec_fallback", "passthrough")
            for fmt, q in plan.get("compression", []):
                if PO.codec_runs_jpeg(fmt, q, fallback):  # uint8 truncation + libjpeg's round trip at int(quality), bit for bit PIL's
                    sl.at(f"compress_{fmt}").libjpeg(int(q))
            if "editing_exposure" in plan:
                sl.at("editing_exposure").gain((plan["editing_exposure"],) * 3)
            sl.at("round").clamp_round()
            return sl
        if plan.get("usm"):
            r = plan["usm"]["radius"]
            if r not in self._usm:
                self._usm[r] = USMSharp(radius=r)
            sl.at("usm").usm(self._usm[r], plan["usm"].get("weight", 0.5), plan["usm"].get("threshold", 10))
        if plan.get("blur1"):
            sl.at("blur1").filter2d(kernel1, an(0))
        if plan.get("resize1"):
            sl.at("resize1").resize(plan["resize1"]["mode"], scale_factor=plan["resize1"]["scale"])
        if plan.get("noise1"):
            noise(plan["noise1"], "noise1")
        if plan.get("jpeg1") is not None:
            sl.at("jpeg1").jpeg(plan["jpeg1"])
        if plan.get("blur2"):
            sl.at("blur2").filter2d(kernel2, an(1))
        if plan.get("resize2"):
            s2 = plan["resize2"]["scale"]
            sl.at("resize2").resize(plan["resize2"]["mode"], size=(int(ori_h / sc * s2), int(ori_w / sc * s2)))
        if plan.get("noise2"):
            noise(plan["noise2"], "noise2")
        jpeg2 = plan.get("jpeg2")
        if plan.get("final_order", "resize_first") == "resize_first":
            tail(jpeg2)
            return sl
        if jpeg2 is not None:
            sl.at("jpeg2").jpeg(jpeg2)
        tail(None)
        return sl

    def _gt_as_view(self) -> bool:
        return self.gt_view and self.pool is None and not self.batch_augment

    @staticmethod
    def _gt_window(gt: Tensor, plan: dict) -> Tensor:
        top, left = plan["crop"]
        sc, size = plan["scale"], plan["gt_size"] // plan["scale"] * plan["scale"]
        return gt[:, :, top * sc : top * sc + size, left * sc : left * sc + size]

    def _native(self, plan: dict | None = None) -> bool:
        # (the fork's extra stages are in the native executor's op table too: one library call per chain in both orders)
        return self.native_chain and not (self.time_stages or self.record_stage_fns or self.collect_taps is not None)

    # -- captured chains -------------------------------------------------------------------
    def _graph_key(self, gt: Tensor, kernels: Sequence[Tensor], plan: dict, inject: dict | None) -> tuple | None:
        if not self.use_graphs or inject or not self._native(plan) or torch.cuda.is_current_stream_capturing():
            return None
        if any(k in plan for k in EXTRA_KEYS):
            return None  # the fork's extras take their drawn scalars by value and come and go from step to step: run eagerly
        if any(k.dtype != torch.float32 or not k.is_contiguous() or not k.is_cuda for k in kernels):
            return None  # the stage list would work on a converted copy whose address is not the caller's
        sig = plan_signature(plan, gt.size(2), gt.size(3))
        if sig is None:
            return None
        return (gt.data_ptr(), tuple(gt.shape), *((k.data_ptr(), tuple(k.shape)) for k in kernels), sig, self._gt_as_view())

    def _fill_params(self, entry: ChainGraph, b: int, h: int, w: int, kernels: Sequence[Tensor], plan: dict) -> None:
        """This step's numbers into the chain's parameter block.  A chain whose rows all came from the plan's own
        per-sample vectors has a fill recipe (chain_graph.fill_recipe): row r <- plan[key][sub]; anything else walks the
        same branches as the capture did with a ParamCollector."""
        params = entry.params
        filled = False
        if entry.recipe is not None:
            try:
                for row, (key, sub) in zip(params.row_views(), entry.recipe):
                    v = plan[key]
                    row.copy_(v if sub is None else v[sub])
                filled = True
            except (KeyError, TypeError, RuntimeError):  # a hand-made plan with other shapes / scalars: the general walk
                filled = False
        if not filled:
            pc = ParamCollector(params, b, h, w)
            self._record(pc, *kernels, plan)  # type: ignore[arg-type]
            if (params.rows, params.noise_stages) != (entry.rows, entry.noise_stages):
                raise RuntimeError("captured chain and plan disagree about the parameter block layout")  # signature bug
        top, left = plan["crop"]
        params.set_header(self.rng.philox.offset, top, left)
        self.rng.philox.offset += entry.noise_stages
        cur = _lib.stream().value or 0
        if entry.last_stream != cur:
            # the block (and the chain's buffers) must not change under a replay still in flight on another stream: an
            # event recorded on that stream NOW covers the replay issued there earlier (recording one after every
            # replay, for the rare caller that hops streams, cost every step a stream lookup and an event record)
            if entry.last_stream is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.ExternalStream(entry.last_stream, device=self.device) if entry.last_stream
                          else torch.cuda.default_stream(self.device))
                torch.cuda.current_stream().wait_event(ev)
            entry.last_stream = cur
        params.upload()

    def _replay(self, entry: ChainGraph, gt: Tensor, plan: dict) -> tuple[Tensor, Tensor]:
        entry.graph.replay()  # type: ignore[union-attr]
        _lib.launch_count += entry.launches
        gt_out = entry.gt_out if entry.gt_out is not None else self._gt_window(gt, plan)
        return gt_out, entry.lq_out  # type: ignore[return-value]

    def _capture(self, key: tuple, gt: Tensor, kernels: Sequence[Tensor], plan: dict) -> ChainGraph | None:
        """Record the chain against a parameter block and capture its launches into a CUDA graph."""
        b, _, h, w = gt.shape
        entry = ChainGraph(ParamBlock(b, self.device))
        try:
            sl = self._record(StageList(gt, params=entry.params), *kernels, plan)  # resize tables are built here, eagerly
            entry.rows, entry.noise_stages = entry.params.rows, entry.params.noise_stages
            entry.recipe = fill_recipe(plan, entry.params.sources)
            top, left = plan["crop"]
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                entry.gt_out, entry.lq_out = sl.run(crop=(gt, plan["gt_size"], plan["scale"], top, left), gt_view=self._gt_as_view())
            entry.launches = _lib.launch_count - l0
            _lib.launch_count = l0  # nothing ran yet: replays are what launches
            entry.graph, entry.keep = g, sl
        except Exception as e:  # noqa: BLE001  a chain that cannot be captured keeps running eagerly
            import warnings

            warnings.warn(f"RealESRGANFeed: chain capture failed ({type(e).__name__}: {e}); this plan shape stays eager", stacklevel=2)
            self.graphs.seen[key] = -(1 << 30)
            return None
        self.graphs.put(key, entry)
        return entry

    def degrade(self, gt: Tensor, kernel1: Tensor, kernel2: Tensor, sinc_kernel: Tensor, plan: dict,
                inject: dict | None = None) -> Tensor:
        """Run the chain described by ``plan`` on device tensors; returns the full-size LQ on the
        8-bit lattice (before the crop).  ONE description of the chain (``_record``) drives both executions: the native
        stage executor (one library call), or — behind the stage hooks — ``StageByStage``, one public per-primitive
        call per stage."""
        if self._native(plan):
            return self._record(StageList(gt), kernel1, kernel2, sinc_kernel, plan, inject).run()
        from .stage_by_stage import StageByStage

        return self._record(StageByStage(self, gt), kernel1, kernel2, sinc_kernel, plan, inject).img

    # -- the reference entry point --------------------------------------------------------
    @torch.no_grad()
    def feed_data(self, data: dict, plan: dict | None = None, inject: dict | None = None) -> None:
        """Accept data from the dataloader and synthesise the LQ batch (realesrgan_model.py:455-650)."""
        if self.is_train:
            if "kernel_params" in data and "kernel1" not in data:
                # extension: the dataset ships 3 x (B,8) parameter tables instead of three
                # (B,21,21) kernels; the kernels are synthesised on the device (kernels.py)
                from .kernels import synthesize_kernels

                def synth(p: Any, which: int) -> Tensor:
                    # one output buffer per (parameter-table address, role): with a prefetcher's static slots the
                    # kernels then live at repeating addresses, which is what lets the captured chain be replayed
                    key = (p.data_ptr() if isinstance(p, Tensor) and p.is_cuda else None, which)
                    buf = self._synth_out.get(key) if key[0] is not None else None
                    k = synthesize_kernels(p, self.device, out=buf)
                    if key[0] is not None:
                        if len(self._synth_out) > 64:
                            self._synth_out.clear()
                        self._synth_out[key] = k
                    return k

                kp = data["kernel_params"]
                if isinstance(kp, Tensor) and kp.dim() == 3 and kp.size(0) == 3:
                    # the three tables stacked as one (3, B, 8) tensor: one upload, ONE synthesis launch for 3B kernels
                    k = self._synth_stacked(kp)
                    data = dict(data, kernel1=k[0], kernel2=k[1], sinc_kernel=k[2])
                else:
                    p1, p2, p3 = kp
                    data = dict(data, kernel1=synth(p1, 0), kernel2=synth(p2, 1), sinc_kernel=synth(p3, 2))
            assert "gt" in data and "kernel1" in data and "kernel2" in data and "sinc_kernel" in data
            gt = data["gt"].to(self.device, non_blocking=True)
            kernel1 = data["kernel1"].to(self.device, non_blocking=True)
            kernel2 = data["kernel2"].to(self.device, non_blocking=True)
            sinc_kernel = data["sinc_kernel"].to(self.device, non_blocking=True)
            if gt.dtype == torch.uint8:
                # extension: an 8-bit GT batch is normalised on the device (x / 255, the division
                # img2tensor does on the host) — a quarter of the PCIe bytes of the reference's fp32 upload
                gt8 = gt.contiguous()
                key = (gt8.data_ptr(), tuple(gt8.shape))
                gt = self._gt_f32.pop(key, None)  # same reasoning as for the kernels: one fp32 buffer per upload slot
                if gt is None:
                    while len(self._gt_f32) >= 8:  # least recently used first (a few slot rings at most are alive at a time)
                        self._gt_f32.pop(next(iter(self._gt_f32)))
                    gt = torch.empty(gt8.shape, dtype=torch.float32, device=self.device)
                self._gt_f32[key] = gt  # (re-inserted: most recently used last)
                _lib.call("otf_u8_to_f32", _lib.ptr(gt8), gt8.numel(), _lib.ptr(gt), _lib.stream())
            gt = _lib.dense_f32(gt)
            ori_h, ori_w = gt.shape[2:4]
            if plan is None:
                plan = draw_plan(self.opt, gt.size(0), ori_h, ori_w, self.rng, gt_size=self.gt_size, order=self.order)
            self.last_plan = plan
            top, left = plan["crop"]
            if plan.get("clean") and plan["scale"] != 1:
                # the clean pass-through keeps LQ at GT size, so the reference's paired_random_crop raises
                # (realesrgan_model.py:491-499 -> transforms.py:106-110)
                raise ValueError(f"Scale mismatches. GT ({ori_h}, {ori_w}) is not {plan['scale']}x ",
                                 f"multiplication of LQ ({ori_h}, {ori_w}). None")
            if self._native(plan):  # chain + crop from one library call
                if plan["gt_size"] % plan["scale"]:
                    raise _lib.OtfError(f"gt_patch_size {plan['gt_size']} must be a multiple of scale {plan['scale']}")
                kernels = (kernel1, kernel2, sinc_kernel)
                key = self._graph_key(gt, kernels, plan, inject)
                entry = self.graphs.get(key) if key is not None else None
                if entry is None and key is not None and self.graphs.should_capture(key):
                    entry = self._capture(key, gt, kernels, plan)
                if entry is not None:  # a captured chain: refresh its parameter block, replay
                    self._fill_params(entry, gt.size(0), ori_h, ori_w, kernels, plan)
                    self.gt, self.lq = self._replay(entry, gt, plan)
                else:
                    sl = self._record(StageList(gt), kernel1, kernel2, sinc_kernel, plan, inject)
                    gt_out, self.lq = sl.run(crop=(gt, plan["gt_size"], plan["scale"], top, left), gt_view=self._gt_as_view())
                    self.gt = gt_out if gt_out is not None else self._gt_window(gt, plan)
            else:
                lq_full = self.degrade(gt, kernel1, kernel2, sinc_kernel, plan, inject)
                self.gt, self.lq = crop_pair(gt, lq_full, plan["gt_size"], plan["scale"], top, left)
            if self.pool is not None:
                self.lq, self.gt = self.pool.step(self.lq, self.gt)
            if self.batch_augment:  # realesrgan_model.py:649-650 (is_train holds on this branch)
                self.gt, self.lq = self.batch_augment(self.gt, self.lq)
        else:
            assert "lq" in data
            self.lq = data["lq"].to(self.device, non_blocking=True)
            if "gt" in data:
                self.gt = data["gt"].to(self.device, non_blocking=True)


class RealESRGANPairedFeed(RealESRGANFeed):
    """The path's second caller, ``RealESRGANPairedModel.feed_data`` (traiNNer/models/realesrgan_paired_model.py:34-67,
    chosen when ``dataroot_lq_prob > 0``): one numpy-generator coin per call; with probability ``dataroot_lq_prob`` the
    pre-made ``paired_lq`` / ``paired_gt`` batch is uploaded as is (MoA applied when it is on), otherwise the ``otf_``
    keys lose their prefix and go through the OTF ``feed_data``."""

    def __init__(self, opt: Any, *args: Any, **kw: Any) -> None:
        super().__init__(opt, *args, **kw)
        self.dataroot_lq_prob = _opt(opt, "dataroot_lq_prob", 0)

    @torch.no_grad()
    def feed_data(self, data: dict, plan: dict | None = None, inject: dict | None = None) -> None:
        if self.rng.np.uniform() < self.dataroot_lq_prob:
            new_data = {k.replace("paired_", ""): v for k, v in data.items() if k.startswith("paired_")}
            assert "lq" in new_data
            self.lq = new_data["lq"].to(self.device, non_blocking=True)
            if "gt" in new_data:
                self.gt = new_data["gt"].to(self.device, non_blocking=True)
            if self.is_train and self.batch_augment and self.gt is not None:
                self.gt, self.lq = self.batch_augment(self.gt, self.lq)
            return
        super().feed_data({k.replace("otf_", ""): v for k, v in data.items() if k.startswith("otf_")}, plan=plan, inject=inject)


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous per-rank slice of a global batch of ``n`` samples: the path has
    no exchange step, so every rank degrades its own samples and keeps its own pool."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
