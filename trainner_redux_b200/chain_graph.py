"""CUDA-graph cache for ``RealESRGANFeed.feed_data``: one captured chain per plan *shape signature*.

``feed_data`` (traiNNer/models/realesrgan_model.py:455-650) draws a fresh plan per call, but everything that varies
per call in a shape-stable schedule is a handful of numbers: per-sample sigma / scale / gray flag / JPEG quality
(degradations.py:673-679, realesrgan_model.py jpeg draws), the crop offsets (transforms.py:119-120) and the position of
the Philox streams.  Those live in ONE small device block (``ParamBlock``) that the captured kernels read through
pointers; a step then costs: draw the plan on the host, one H2D copy of the block (a few hundred bytes), one graph
launch.  What is baked into a graph — and therefore forms the cache key — is the chain's structure (which stages run,
modes, every image extent), the batch shape and the addresses of the input tensors.

Block layout (bytes): [0, 8) uint64 Philox offset base (``otf_gaussian_noise_f32`` / ``otf_poisson_noise_f32``
``offset_dev``); [8, 16) int32 crop (top, left) (``otf_crop_pair_f32`` ``top_left_dev``); then rows of B fp32.
"""

from __future__ import annotations

from collections import OrderedDict
from typing import Any

import torch
from torch import Tensor

HEADER_BYTES = 16


class ParamBlock:
    """The per-step numbers of one chain: a host staging buffer and its device twin."""

    def __init__(self, b: int, device: torch.device, max_rows: int = 16) -> None:
        self.b, self.max_rows = b, max_rows
        nbytes = HEADER_BYTES + 4 * b * max_rows
        # pageable on purpose: the runtime stages a small pageable source before cudaMemcpyAsync returns, so the buffer
        # may be refilled for the next step while this step's copy is still queued (a pinned source could not)
        self.host = torch.zeros(nbytes, dtype=torch.uint8)
        self.dev = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self._rows = self.host[HEADER_BYTES:].view(torch.float32).view(max_rows, b)
        self._views: tuple | None = None
        self._hdr = self.host[:HEADER_BYTES].numpy().view("int64"), self.host[:HEADER_BYTES].numpy().view("int32")
        self.rows = 0  # rows in use
        self.noise_stages = 0  # Philox offsets one pass of the chain consumes
        self.sources: list[Any] = []  # what went into each row, in order (the capture turns it into a fill recipe)

    # -- addresses the stage records point at --------------------------------------------------
    @property
    def offset_ptr(self) -> int:
        return self.dev.data_ptr()

    @property
    def crop_ptr(self) -> int:
        return self.dev.data_ptr() + 8

    def row_ptr(self, r: int) -> int:
        return self.dev.data_ptr() + HEADER_BYTES + 4 * self.b * r

    # -- filling (host side, once per step) -----------------------------------------------------
    def begin(self) -> None:
        self.rows = 0
        self.noise_stages = 0
        self.sources = []

    def add_row(self, v: float | Tensor) -> int:
        """Append one per-sample vector (or a scalar broadcast over the batch); returns its device address."""
        if self.rows >= self.max_rows:
            raise RuntimeError(f"ParamBlock: more than {self.max_rows} per-sample vectors in one chain")
        r = self.rows
        if isinstance(v, (int, float)):
            self._rows[r].fill_(float(v))
        else:
            self._rows[r].copy_(v.reshape(self.b))
        self.rows += 1
        self.sources.append(v)
        return self.row_ptr(r)

    def next_noise_index(self) -> int:
        k = self.noise_stages
        self.noise_stages += 1
        return k

    def set_header(self, philox_offset: int, top: int, left: int) -> None:
        q, d = self._hdr  # numpy views of the same bytes (a scalar store into a torch tensor costs ~10x as much)
        q[0] = philox_offset
        d[2] = top
        d[3] = left

    def upload(self) -> None:
        n = HEADER_BYTES + 4 * self.b * self.rows
        if self._views is None or self._views[0] != n:
            self._views = (n, self.dev[:n], self.host[:n], [self._rows[r] for r in range(self.rows)])
        self._views[1].copy_(self._views[2], non_blocking=True)

    def row_views(self) -> list[Tensor]:
        """Host views of the rows in use (built once per layout: indexing a tensor costs more than filling 64 floats)."""
        n = HEADER_BYTES + 4 * self.b * self.rows
        if self._views is None or self._views[0] != n:
            self._views = (n, self.dev[:n], self.host[:n], [self._rows[r] for r in range(self.rows)])
        return self._views[3]


class ParamCollector:
    """Stands in for ``StageList`` when a cached graph is replayed: ``RealESRGANFeed._record`` walks the very same
    branches, and this object only collects what goes into the parameter block, in the order the capture laid it out."""

    def __init__(self, params: ParamBlock, b: int, h: int, w: int) -> None:
        self.params, self.b, self.h, self.w = params, b, h, w
        params.begin()

    def at(self, name: str) -> "ParamCollector":
        return self

    def analyse(self, kernels: Any) -> None:
        pass

    def filter2d(self, kernel: Tensor, analysed_set: int | None = None) -> None:
        pass

    def usm(self, taps: Any, weight: float, threshold: float) -> None:
        pass

    def resize(self, mode: str, scale_factor: float = 0, size: tuple[int, int] = (0, 0)) -> None:
        if scale_factor != 0:
            size = (round(self.h * scale_factor), round(self.w * scale_factor))
        self.h, self.w = int(size[0]), int(size[1])

    def _noise(self, value: float | Tensor, gray: float | Tensor | None) -> None:
        self.params.next_noise_index()
        self.params.add_row(value)
        if not (gray is None or (isinstance(gray, (int, float)) and gray <= 0)):
            self.params.add_row(gray)

    def gaussian_noise(self, sigma: Any, gray: Any, gen: Any, **kw: Any) -> None:
        self._noise(sigma, gray)

    def poisson_noise(self, scale: Any, gray: Any, gen: Any, **kw: Any) -> None:
        self._noise(scale, gray)

    def jpeg(self, quality: float | Tensor, differentiable: bool = False, clamp_in: bool = True, round8: bool = False) -> None:
        self.params.add_row(quality)

    def clamp_round(self) -> None:
        pass


def _resize_sig(h: int, w: int, st: dict | None, size: tuple[int, int] | None = None) -> tuple:
    if st is None:
        return (None, h, w)
    if size is None:
        size = (round(h * st["scale"]), round(w * st["scale"]))
    return (st["mode"], int(size[0]), int(size[1]))


def _noise_sig(st: dict | None) -> tuple | None:
    if not st:
        return None
    g = st.get("gray")
    return (st["kind"], not (g is None or (isinstance(g, (int, float)) and g <= 0)))


def plan_signature(plan: dict, h: int, w: int) -> tuple | None:
    """Everything about ``plan`` that a captured chain bakes in; None when the plan cannot be captured (per-sample
    vectors that already live on the device are pointers the graph would freeze)."""
    for key in ("noise1", "noise2"):
        st = plan.get(key)
        if st and any(isinstance(v, Tensor) and v.is_cuda for v in st.values()):
            return None
    for key in ("jpeg1", "jpeg2", "jpeg"):
        q = plan.get(key)
        if isinstance(q, Tensor) and q.is_cuda:
            return None
    sc = plan["scale"]
    head = (plan.get("order", "classic"), sc, plan["gt_size"], bool(plan.get("clean")))
    if plan.get("clean"):
        return head
    if plan.get("order") == "fork":
        return head + (bool(plan.get("blur1")), plan["resize3_mode"], plan.get("jpeg") is not None)
    usm = plan.get("usm")
    r1 = _resize_sig(h, w, plan.get("resize1"))
    r2 = plan.get("resize2")
    r2s = _resize_sig(r1[1], r1[2], r2, (int(h / sc * r2["scale"]), int(w / sc * r2["scale"]))) if r2 else (None, r1[1], r1[2])
    return head + (
        (usm["radius"], usm.get("weight", 0.5), usm.get("threshold", 10)) if usm else None, bool(plan.get("blur1")), r1,
        _noise_sig(plan.get("noise1")), plan.get("jpeg1") is not None, bool(plan.get("blur2")), r2s, _noise_sig(plan.get("noise2")),
        plan.get("final_order", "resize_first"), plan["resize3_mode"], plan.get("jpeg2") is not None)


class ChainGraph:
    """One captured chain: graph, parameter block, outputs."""

    def __init__(self, params: ParamBlock) -> None:
        self.params = params
        self.graph: torch.cuda.CUDAGraph | None = None
        self.gt_out: Tensor | None = None
        self.lq_out: Tensor | None = None
        self.launches = 0
        self.noise_stages = 0
        self.rows = 0
        self.last_stream: int | None = None  # raw handle of the stream of the last replay (None: never replayed)
        self.keep: Any = None  # the stage list of the capture: owns the host tap arrays / tensors the launches point into
        # fill recipe: for each row of the parameter block, where in a plan its vector comes from — (key, sub) for
        # plan[key][sub], (key, None) for plan[key].  None = unknown, walk the chain with a ParamCollector instead
        self.recipe: list[tuple[str, str | None]] | None = None


def fill_recipe(plan: dict, sources: list) -> list[tuple[str, str | None]] | None:
    """Match the tensors a capture put into its parameter block against the plan they came from (by identity)."""
    where: dict[int, tuple[str, str | None]] = {}
    for key in ("noise1", "noise2"):
        st = plan.get(key)
        if isinstance(st, dict):
            for sub, v in st.items():
                if isinstance(v, Tensor):
                    where[id(v)] = (key, sub)
    for key in ("jpeg1", "jpeg2", "jpeg"):
        v = plan.get(key)
        if isinstance(v, Tensor):
            where[id(v)] = (key, None)
    out = []
    for v in sources:
        if not isinstance(v, Tensor) or id(v) not in where:
            return None
        out.append(where[id(v)])
    return out


class ChainGraphCache:
    """LRU of captured chains keyed by (input addresses, batch shape, plan signature).  A key is captured the SECOND
    time it is seen — a freshly drawn resize scale rarely repeats, a shape-stable schedule repeats at once.

    A capture costs milliseconds (torch synchronises the device and collects garbage around it), so captures are
    rationed: the cache starts with a few capture credits, every replay earns a fraction of one back.  A shape-stable
    schedule pays for its handful of graphs at once and then only replays; a random schedule whose shapes recur now and
    then (a resize scale drawn from a continuum) spends its credits, finds that replays do not pay them back, and stays
    on the eager path instead of capturing-and-evicting forever."""

    def __init__(self, capacity: int = 8, capture_after: int = 2, credits: float = 4.0, credit_per_hit: float = 0.02) -> None:
        self.capacity, self.capture_after = capacity, capture_after
        self.credits, self.max_credits, self.credit_per_hit = credits, credits, credit_per_hit
        self.entries: OrderedDict[tuple, ChainGraph] = OrderedDict()
        self.seen: OrderedDict[tuple, int] = OrderedDict()
        self.hits = self.misses = self.captures = 0

    def get(self, key: tuple) -> ChainGraph | None:
        e = self.entries.get(key)
        if e is not None:
            self.entries.move_to_end(key)
            self.hits += 1
            self.credits = min(self.max_credits, self.credits + self.credit_per_hit)
        else:
            self.misses += 1
        return e

    def should_capture(self, key: tuple) -> bool:
        n = self.seen.get(key, 0) + 1
        self.seen[key] = n
        self.seen.move_to_end(key)
        while len(self.seen) > 512:
            self.seen.popitem(last=False)
        if n < self.capture_after or self.credits < 1.0:
            return False
        self.credits -= 1.0
        return True

    def put(self, key: tuple, entry: ChainGraph) -> None:
        self.entries[key] = entry
        self.captures += 1
        while len(self.entries) > self.capacity:
            self.entries.popitem(last=False)  # the graph, its workspace and its outputs go with the entry
