"""CPU tests: the C-ABI library loads and exports every symbol the header declares; host-side
logic (plan draws, pair pool indirection, shard ranges, world_size-2 gloo run); numpy restatement
of ATen's resampling rules.  No kernel is launched here."""

import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import np_resample as NR
from oracle import otf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from trainner_redux_b200 import _lib

    lib = _lib.load()
    header = open(_lib.HEADER_PATH).read()
    declared = set(re.findall(r"\b(otf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.otf_abi_version() == _lib.ABI_VERSION == 6
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (otf_[a-z0-9_]+)", out))
    assert declared <= exported


def test_library_is_sm100a_only():
    from trainner_redux_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    import trainner_redux_b200 as T

    x = torch.rand(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        T.filter2d(x, torch.ones(1, 3, 3) / 9)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        T.resize_pt(x, "bilinear", scale_factor=0.5)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        T.DiffJPEG(differentiable=False)(x, quality=50)
    with pytest.raises(ValueError, match="Wrong kernel size"):
        T.filter2d(x, torch.ones(1, 4, 4))
    with pytest.raises(ValueError, match="scale_factor or size is required"):
        T.resize_pt(x, "bilinear")
    import glob

    mods = sorted(os.path.basename(f)[:-3] for f in glob.glob(os.path.join(ROOT, "trainner_redux_b200", "*.py")))
    assert {"img_process_util", "diffjpeg", "degradations", "transforms", "realesrgan_feed", "paragon_otf", "batchaug", "stages", "_lib"} <= set(mods)
    for mod in mods:  # every module of the product package
        src = open(os.path.join(ROOT, "trainner_redux_b200", mod + ".py")).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{mod} must not import the oracle"


def test_draw_plan_order_and_rank_semantics():
    from trainner_redux_b200.realesrgan_feed import HostRNG, OTFOptions, draw_plan

    opt = OTFOptions(blur_prob=0.5, blur_prob2=0.5, gaussian_noise_prob=0.5, noise_range=(1, 30), poisson_scale_range=(0.05, 3),
                     gray_noise_prob=0.4, gaussian_noise_prob2=0.5, noise_range2=(1, 25), poisson_scale_range2=(0.05, 2.5),
                     gray_noise_prob2=0.4, jpeg_prob=0.8, jpeg_prob2=0.8, jpeg_range=(30, 95), jpeg_range2=(30, 95), gt_size=224)
    a = draw_plan(opt, 8, 256, 256, HostRNG(7, rank=0))
    b = draw_plan(opt, 8, 256, 256, HostRNG(7, rank=0))
    c = draw_plan(opt, 8, 256, 256, HostRNG(7, rank=1))

    def flat(v):
        if torch.is_tensor(v):
            return v.tolist()
        if isinstance(v, dict):
            return {k: flat(x) for k, x in v.items()}
        return v

    assert flat(a) == flat(b), "same (seed, rank) -> same plan"
    # numpy coins are NOT rank-offset (rng.py:19-25): stage on/off identical across ranks...
    for k in ("blur1", "blur2", "final_order"):
        assert a[k] == c[k]
    assert (a["jpeg1"] is None) == (c["jpeg1"] is None) and a["noise1"]["kind"] == c["noise1"]["kind"]
    # ...while torch/random draws differ per rank (train.py:283)
    va = a["noise2"].get("sigma", a["noise2"].get("scale"))
    vc = c["noise2"].get("sigma", c["noise2"].get("scale"))
    assert a["noise1"]["gray"].shape == (8,) and not torch.equal(va, vc)
    assert 0 <= a["crop"][0] <= 64 - 56 and 0 <= a["crop"][1] <= 64 - 56
    assert a["resize1"]["mode"] in opt.resize_mode_list and 0.4 <= a["resize1"]["scale"] <= 1.5
    with pytest.raises(ValueError, match="smaller than patch size"):
        draw_plan(OTFOptions(gt_size=512), 2, 256, 256, HostRNG(0))
    fork = draw_plan(OTFOptions(order="fork", blur_prob=1.0, gt_size=128), 2, 160, 160, HostRNG(0))
    assert fork["blur1"] is True and "resize1" not in fork and fork["resize3_mode"] in OTFOptions().resize_mode_list3


class _TorchMover:  # CPU stand-in for the slot gather/scatter kernels (test only)
    def gather(self, src, idx):
        return src[torch.tensor(idx)].clone()

    def scatter(self, dst, idx, src):
        dst[torch.tensor(idx)] = src


def test_pair_pool_indirection_equals_reference_pool():
    from trainner_redux_b200.realesrgan_feed import PairPool

    perms = [torch.randperm(12, generator=torch.Generator().manual_seed(s)) for s in range(20)]
    it = iter(perms)
    mine = PairPool(12, mover=_TorchMover(), randperm=lambda n: next(it))
    ref = O.PairPool(12)
    used = 0
    for step in range(15):
        g = torch.Generator().manual_seed(100 + step)
        lq, gt = torch.rand(4, 3, 5, 5, generator=g), torch.rand(4, 3, 20, 20, generator=g)
        perm = None
        if ref.queue_ptr == 12:
            perm = perms[used]
            used += 1
        r_lq, r_gt = ref.step(lq, gt, perm)
        m_lq, m_gt = mine.step(lq, gt)
        assert torch.equal(r_lq, m_lq) and torch.equal(r_gt, m_gt), f"step {step}"
    assert used == 12
    with pytest.raises(AssertionError, match="divisible by batch size"):
        PairPool(10, mover=_TorchMover()).step(torch.zeros(4, 1, 1, 1), torch.zeros(4, 1, 1, 1))


def test_shard_range_partitions():
    from trainner_redux_b200.realesrgan_feed import shard_range

    for n in (64, 32, 7, 1):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["OTF_ROOT"])
from trainner_redux_b200.realesrgan_feed import HostRNG, OTFOptions, draw_plan, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
opt = OTFOptions(blur_prob=0.5, gaussian_noise_prob=0.5, noise_range=(1, 30), poisson_scale_range=(0.05, 3), jpeg_range=(30, 95), gt_size=224)
lo, hi = shard_range(64, rank, world)
plan = draw_plan(opt, hi - lo, 256, 256, HostRNG(5, rank))
mine = torch.tensor([lo, hi, int(plan["blur1"]), plan["crop"][0], plan["crop"][1], int(plan["jpeg1"][0].item() * 1000)])
allv = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allv, mine)
t = torch.tensor([float(hi - lo)])
dist.all_reduce(t)  # what bench.py does with the per-rank pair counts
if rank == 0:
    assert allv[0][0] == 0 and allv[-1][1] == 64 and all(allv[i][1] == allv[i + 1][0] for i in range(world - 1))
    assert len({int(v[2]) for v in allv}) == 1, "stage coins must agree across ranks"
    assert len({int(v[5]) for v in allv}) == world, "per-sample draws must differ across ranks"
    assert t.item() == 64
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, OTF_ROOT=ROOT)
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29653", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
@pytest.mark.parametrize("case", [((40, 36), (17, 50)), ((96, 96), (39, 143)), ((33, 20), (33, 7))])
def test_numpy_restatement_of_aten_antialias_rules(mode, case):
    (h, w), (oh, ow) = case
    x = torch.rand(1, 1, h, w, generator=torch.Generator().manual_seed(h * w))
    want = F.interpolate(x, size=(oh, ow), mode=mode, antialias=True)[0, 0].numpy()
    got = NR.resize_aa(x[0, 0].numpy(), oh, ow, mode)
    assert np.abs(got - want).max() < 1e-6


def test_numpy_area_and_nearest_rules():
    x = torch.rand(1, 1, 37, 41)
    for out_n in (5, 37, 40, 64):
        idx = NR.nearest_exact_index(41, out_n)
        want = F.interpolate(x, size=(37, out_n), mode="nearest-exact")[0, 0]
        assert torch.equal(x[0, 0][:, idx], want)
    for out_n in (5, 13, 41):
        win = NR.area_windows(41, out_n)
        want = F.interpolate(x, size=(37, out_n), mode="area")[0, 0]
        got = torch.stack([x[0, 0][:, a:b].mean(1) for a, b in win], 1)
        assert (got - want).abs().max() < 1e-6


def test_lanczos_taps_host_side_match_oracle():
    from trainner_redux_b200.degradations import _lanczos_taps

    for ratio in (0.4, 0.75, 0.25, 0.9, 117 / 288, 0.1):
        a, b = _lanczos_taps(ratio), O._lanczos_taps(ratio).numpy()
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-7


class _Frozen:
    """Stand-in for the reference's msgspec StrictStruct options (redux_options.py): attributes can be read, never set."""

    def __init__(self, **kw):
        object.__setattr__(self, "_d", dict(kw))

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_d")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        raise AttributeError(f"{type(self).__name__} is frozen (cannot set {name!r})")


def test_redux_options_shaped_object_is_accepted_as_is():
    """A ReduxOptions-shaped object: no `order`, no top-level `gt_size` (it lives in opt.datasets["train"]), MoA fields
    under opt.train, every attribute read-only.  The feed resolves all of them without writing to the object."""
    import dataclasses

    from trainner_redux_b200.realesrgan_feed import HostRNG, OTFOptions, draw_plan, resolve_gt_size, resolve_order

    fields = {f.name: getattr(OTFOptions(), f.name) for f in dataclasses.fields(OTFOptions)}
    for k in ("gt_size", "order", "use_moa", "moa_augs", "moa_probs", "final_jpeg_first_prob", "noise_enabled", "fork_compression", "codec_fallback"):
        fields.pop(k)
    opt = _Frozen(**fields, datasets={"train": _Frozen(gt_size=128, lq_size=None)},
                  train=_Frozen(use_moa=True, moa_augs=["none", "mixup"], moa_probs=[0.5, 0.5], moa_debug=False, moa_debug_limit=0))
    with pytest.raises(AttributeError):
        opt.gt_size = 128
    assert resolve_gt_size(opt) == 128 and resolve_order(opt) == "fork"
    assert resolve_gt_size(OTFOptions(gt_size=96)) == 96 and resolve_order(OTFOptions()) == "classic"
    plan = draw_plan(opt, 4, 160, 160, HostRNG(5), gt_size=resolve_gt_size(opt), order=resolve_order(opt))
    assert plan["order"] == "fork" and plan["gt_size"] == 128 and "resize3_mode" in plan and "resize1" not in plan
    assert 0 <= plan["crop"][0] <= 160 // 4 - 32
    # the p_clean gate is drawn first even at probability 0 (realesrgan_model.py:487-489): the lens gate sees the 2nd uniform
    import numpy as np

    ref = np.random.default_rng(5)
    ref.uniform()
    opt2 = _Frozen(**{**fields, "lens_distort_prob": 1.0}, datasets={"train": _Frozen(gt_size=128)})
    plan2 = draw_plan(opt2, 4, 160, 160, HostRNG(5), gt_size=128, order="fork")
    ref.uniform()  # lens gate
    assert plan2["lens"] == float(ref.uniform(*opt2.lens_distort_strength_range))
    # explicit overrides win
    assert draw_plan(opt, 4, 160, 160, HostRNG(5), gt_size=64, order="classic")["order"] == "classic"
