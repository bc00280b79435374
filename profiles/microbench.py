"""Per-stage CUDA-event timings at the bench shapes (B=64): BASELINE.json configs[3] microbench.
    python profiles/microbench.py [--reps 20] [--json out.json]
Each row: GPU time per launch from a CUDA-graph replay of back-to-back launches whose inputs rotate
over >256 MB of distinct buffers (so every launch reads HBM, not L2), algorithmic bytes,
achieved GB/s and fraction of the measured HBM peak; filter2d rows also report TFLOP/s at the true
(zero-trimmed) tap count against the FP32 FMA roof."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trainner_redux_b200 as T  # noqa: E402
from trainner_redux_b200 import synthetic as S  # noqa: E402
from trainner_redux_b200.kernels import synthesize_kernels
from trainner_redux_b200.img_process_util import KernelAnalysis  # noqa: E402
from trainner_redux_b200 import degradations as D  # noqa: E402
from trainner_redux_b200.realesrgan_feed import clamp_round  # noqa: E402
from trainner_redux_b200.transforms import crop_pair  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--json", default=None)
ap.add_argument("--only", default=None)
args = ap.parse_args()
dev = torch.device("cuda:0")
peak = 6464.0
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
FMA_PEAK = 148 * 128 * 2 * 1.965e9 / 1e12
B = int(os.environ.get("OTF_MB_BATCH", "64"))
ROTATE_BYTES = 256 * 1024 * 1024  # inputs rotate over > 2x the 126 MB L2, so every launch reads HBM


def timeit(make_fn, in_bytes):
    """make_fn(i) -> callable for input set i.  The launches (inputs rotating over nbuf distinct
    buffers) are captured into one CUDA graph and the graph replay is timed with one event pair, so
    the number is GPU time only: no Python/ctypes launch overhead, no per-launch events, no flush
    traffic in the timed region."""
    nbuf = max(2, min(64, -(-ROTATE_BYTES // max(in_bytes, 1))))
    fns = [make_fn(i) for i in range(nbuf)]
    for f in fns[:3]:
        f()
    torch.cuda.synchronize()
    D.pin_resize_tables()  # weight tables built during warm-up are reused by the captured launches
    n = max(nbuf, args.reps)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for r in range(n):
                fns[r % nbuf]()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    del g
    return best


def img(h, w, seed=1):
    return S.synth_gt(B, h, w, "uniform", seed=seed).to(dev)


rows = []


def add(name, fn, nbytes, flops=None, x=None):
    """fn(x) runs the kernel on input tensor x; x is cloned into the rotation buffers."""
    if args.only and not any(o in name for o in args.only.split(",")):
        return
    xs = {}

    def make(i):
        if i not in xs:
            xs[i] = x.clone() if i else x
        xi = xs[i]
        return lambda: fn(xi)

    ms = timeit(make, x.numel() * 4)
    gbs = nbytes / ms / 1e6
    r = {"kernel": name, "ms": round(ms, 4), "alg_MB": round(nbytes / 1e6, 2), "GBps": round(gbs, 1), "frac_hbm": round(gbs / peak, 3)}
    if flops:
        r["TFLOPs"] = round(flops / ms / 1e9, 2)
        r["frac_fma"] = round(flops / ms / 1e9 / FMA_PEAK, 3)
    rows.append(r)
    print(r, flush=True)


x256, x192, x64 = img(256, 256), img(192, 192, 2), img(64, 64, 3)
N = lambda t: t.numel() * 4
sigma = (torch.rand(B) * 29 + 1).to(dev)
gray = (torch.rand(B) < 0.4).float().to(dev)
nogray = torch.zeros(B, device=dev)
q = (torch.rand(B) * 65 + 30).to(dev)
jp = T.DiffJPEG(differentiable=False)


def taps(k):
    return float((k != 0).flatten(1).sum(1).float().mean())


def kernels(seed, kernel_list=None, sinc_prob=0.1, which=0):
    """(B,21,21) kernels synthesised on the device from the reference's parameter distributions (sizes 7..21)."""
    return synthesize_kernels(S.synth_kernel_params(B, seed, S.bench_kernel_options(kernel_list, sinc_prob))[which], dev)


for kinds, sp in ((None, 0.1), (("iso",), 0.0), (("aniso",), 0.0), (("generalized_iso", "plateau_iso"), 0.0), (("iso",), 1.0)):
    k = kernels(10, kinds, sp)
    label = "default mix" if kinds is None else ("sinc" if sp == 1.0 else "/".join(kinds))
    add(f"filter2d 256^2 sizes 7..21 {label}", lambda t, k=k: T.filter2d(t, k), 2 * N(x256), 2 * taps(k) * x256.numel(), x=x256)
    ka = KernelAnalysis([k])  # as the chain runs it: one analysis launch per step for all kernel sets, outside this row
    add(f"filter2d 256^2 sizes 7..21 {label} (pre-analysed)", lambda t, k=k, ka=ka: T.filter2d(t, k, _analysis=(ka, 0)), 2 * N(x256),
        2 * taps(k) * x256.numel(), x=x256)
for ks in (7, 9, 13, 17, 21):
    k = torch.rand(B, ks, ks, device=dev)
    k = k / k.sum((1, 2), keepdim=True)
    add(f"filter2d 256^2 dense K={ks}", lambda t, k=k: T.filter2d(t, k), 2 * N(x256), 2 * ks * ks * x256.numel(), x=x256)
k2 = kernels(20, which=1)
add("filter2d 192^2 mixed", lambda t: T.filter2d(t, k2), 2 * N(x192), 2 * taps(k2) * x192.numel(), x=x192)
sk = kernels(30, which=2)
add("filter2d 64^2 final sinc", lambda t: T.filter2d(t, sk), 2 * N(x64), 2 * taps(sk) * x64.numel(), x=x64)
ka_s = KernelAnalysis([sk])
add("filter2d 64^2 final sinc (pre-analysed)", lambda t: T.filter2d(t, sk, _analysis=(ka_s, 0)), 2 * N(x64), 2 * taps(sk) * x64.numel(), x=x64)
k7 = torch.rand(B, 7, 7, device=dev)
k7 = torch.nn.functional.pad(k7 / k7.sum((1, 2), keepdim=True), (7, 7, 7, 7))
ka7 = KernelAnalysis([k7])
add("filter2d 256^2 dense K=7 in 21x21 (pre-analysed)", lambda t: T.filter2d(t, k7, _analysis=(ka7, 0)), 2 * N(x256), 2 * 49 * x256.numel(), x=x256)
for mode in ("bilinear", "bicubic", "area", "nearest-exact", "lanczos"):
    for s in (0.4, 0.75, 1.25, 1.5):
        oh = round(256 * s)
        add(f"resize {mode} 256->{oh}", lambda t, mode=mode, s=s: T.resize_pt(t, mode, scale_factor=s), N(x256) + B * 3 * oh * oh * 4, x=x256)
add("resize bilinear 192->64", lambda t: T.resize_pt(t, "bilinear", size=(64, 64)), N(x192) + N(x64), x=x192)
add("resize area 64->64", lambda t: T.resize_pt(t, "area", size=(64, 64)), 2 * N(x64), x=x64)
add("gaussian colour 192^2", lambda t: D.add_gaussian_noise_pt(t, sigma, nogray), 2 * N(x192), x=x192)
add("gaussian 40% gray 192^2", lambda t: D.add_gaussian_noise_pt(t, sigma, gray), 2 * N(x192), x=x192)
add("gaussian 40% gray 64^2", lambda t: D.add_gaussian_noise_pt(t, sigma, gray), 2 * N(x64), x=x64)
add("poisson colour 192^2", lambda t: D.add_poisson_noise_pt(t, sigma / 10, True, False, nogray), 3 * N(x192), x=x192)
add("poisson 40% gray 192^2", lambda t: D.add_poisson_noise_pt(t, sigma / 10, True, False, gray), 3 * N(x192), x=x192)
for qq in (30, 50, 75, 95):
    add(f"diffjpeg q={qq} 192^2", lambda t, qq=qq: jp(t, quality=float(qq)), 2 * N(x192), x=x192)
add("diffjpeg per-sample q 192^2 (+factor kernel)", lambda t: jp(t, quality=q.clone(), _clamp_in=True), 2 * N(x192), x=x192)
add("diffjpeg per-sample q 64^2 +round8", lambda t: jp(t, quality=q.clone(), _clamp_in=True, _round8=True), 2 * N(x64), x=x64)
usm = T.USMSharp().to(dev)
add("usm 256^2 (4 launches)", lambda t: usm(t), 2 * N(x256), x=x256)
add("clamp_round 64^2", lambda t: clamp_round(t), 2 * N(x64), x=x64)
add("crop_pair 256/64 -> 224/56", lambda t: crop_pair(t, x64, 224, 4, 4, 4), 2 * B * 3 * (224 * 224 + 56 * 56) * 4, x=x256)

# ---- row g1: the executor's fused launches beside the launches they replace ---------------------------------------
from trainner_redux_b200 import _lib as L  # noqa: E402


def _fused_resize_gauss(t, oh, mode_id, gp, fused=True):
    b, c, h, w = t.shape
    tab = D.pinned_resize_table(dev, h, w, oh, oh, mode_id)
    nb = L.load().otf_resize_workspace_bytes(h, w, oh, oh, mode_id)
    out = torch.empty(b, c, oh, oh, device=dev)
    L.call("otf_resize_gauss_f32", L.ptr(t), b, c, h, w, L.ptr(out), oh, oh, mode_id, 1, L.ptr(tab), nb, 1, L.ptr(sigma), L.ptr(gp),
           7, 1, None, L.NOISE_CLIP, L.stream())
    return out


for gp, lab in ((None, "colour"), (gray, "40% gray")):
    add(f"g1 resize bicubic 256->192 + gaussian {lab} (1 launch)", lambda t, gp=gp: _fused_resize_gauss(t, 192, L.RESIZE_BICUBIC_AA, gp),
        N(x256) + N(x192), x=x256)
    add(f"g1 resize bilinear 192->64 + gaussian {lab} (1 launch)", lambda t, gp=gp: _fused_resize_gauss(t, 64, L.RESIZE_BILINEAR_AA, gp),
        N(x192) + N(x64), x=x192)
gt_out = torch.empty(B, 3, 224, 224, device=dev)
lq_out = torch.empty(B, 3, 56, 56, device=dev)
add("g1 diffjpeg 64^2 + lattice + LQ crop + GT crop 224 (1 launch)",
    lambda t: L.call("otf_diffjpeg_crop_pair_f32", L.ptr(t), B, 64, 64, L.ptr(q), 0.0, 1, 0, 1, L.ptr(x256), 256, 256, 4, 4, None, 56, 4,
                     L.ptr(gt_out), L.ptr(lq_out), L.stream()), N(x64) + 2 * B * 3 * 224 * 224 * 4 + B * 3 * 56 * 56 * 4, x=x64)
add("g1 clamp/round + crop_pair 256/64 -> 224/56 (1 launch)",
    lambda t: L.call("otf_crop_pair_f32", L.ptr(t), B * 3, 256, 256, L.ptr(x64), 64, 64, 4, 4, None, 56, 4, 1, L.ptr(gt_out), L.ptr(lq_out),
                     L.stream()), 2 * B * 3 * (224 * 224 + 56 * 56) * 4, x=x256)

# ---- "next" rows of SURVEY.md §8f: pair pool (f1), kernel synthesis (f2), uint8 upload + MoA (f4) --------------
import random  # noqa: E402

import numpy as np  # noqa: E402

from trainner_redux_b200 import _lib, batchaug as BA  # noqa: E402
from trainner_redux_b200.realesrgan_feed import PairPool  # noqa: E402


class _Rng:
    def __init__(self, seed):
        self.py, self.np, self.torch = random.Random(seed), np.random.default_rng(seed), torch.Generator().manual_seed(seed)


gt224, lq56 = img(224, 224, 5), img(56, 56, 6)
pair_bytes = N(gt224) + N(lq56)
pool = PairPool(192, randperm=lambda n: torch.randperm(n, generator=torch.Generator().manual_seed(1)))
for _ in range(3):  # fill the pool: 192 / 64
    pool.step(lq56, gt224)
add("pair pool step, queue 192 (2 gathers + 2 scatters of 64 slots)", lambda t: pool.step(lq56, t), 4 * pair_bytes, x=gt224)
params = S.synth_kernel_params(B, seed=3)[0] if hasattr(S, "synth_kernel_params") else None
if params is not None:
    pdev = torch.as_tensor(params, dtype=torch.float64).to(dev)
    add("kernel synthesis 64 x 21x21 (fp64)", lambda t: synthesize_kernels(pdev, dev), B * (64 + 441 * 4), x=x64)
u8 = (x256 * 255).round().to(torch.uint8)
f32 = torch.empty_like(x256)
add("u8 -> f32 /255 256^2", lambda t: _lib.call("otf_u8_to_f32", _lib.ptr(u8), u8.numel(), _lib.ptr(f32), _lib.stream()),
    u8.numel() * 5, x=x64)
for name, fn, nb in (("mixup", BA.mixup, 3 * pair_bytes), ("cutmix", BA.cutmix, None), ("resizemix", BA.resizemix, None),
                     ("cutblur", BA.cutblur, None)):
    add(f"moa {name} 224^2/56^2 pair", lambda t, fn=fn: fn(t, lq56, 4, rng=_Rng(11)), nb or 2 * pair_bytes, x=gt224)
add("moa downup 56^2", lambda t: BA.downup(gt224, t, rng=_Rng(12)), 4 * N(lq56), x=lq56)
add("moa up 224^2/56^2 pair", lambda t: BA.up(t, lq56, 4, rng=_Rng(13)), 2 * pair_bytes, x=gt224)

# ---- row f3: the fork's extra stages (paragon_otf.py) at 256^2 ---------------------------------------------------
from trainner_redux_b200 import paragon_otf as PO  # noqa: E402

add("f3 lens distortion 256^2", lambda t: PO.lens_distortion(t, 0.2), 2 * N(x256), x=x256)
add("f3 rolling shutter 256^2", lambda t: PO.rolling_shutter(t, 0.08), 2 * N(x256), x=x256)
add("f3 chromatic aberration 256^2", lambda t: PO.chromatic_aberration(t), 2 * N(x256), x=x256)
add("f3 motion blur K=15 256^2", lambda t: PO.motion_blur(t, 15, 30.0), 2 * N(x256), x=x256)
add("f3 oversharpen (5x5 box) 256^2", lambda t: PO.oversharpen(t, 1.5), 2 * N(x256), x=x256)
add("f3 exposure 256^2", lambda t: PO.exposure(t, 1.3), 2 * N(x256), x=x256)
add("f3 colour temperature 256^2", lambda t: PO.color_temperature(t, 0.1), 2 * N(x256), x=x256)
add("f3 sensor noise 256^2", lambda t: PO.sensor_noise(t, 0.05), 2 * N(x256), x=x256)
add("f3 aliasing x0.75 256^2 (2 launches)", lambda t: PO.aliasing(t, 0.75), 2 * N(x256) + 2 * N(x192), x=x256)
add("f3 jpeg round 64^2 (libjpeg round trip, bit-exact PIL: 2 launches)", lambda t: PO.compress_with_format(t, "jpeg", 77.0), 2 * N(x64), x=x64)
add("f3 jpeg round 192^2 (libjpeg round trip, bit-exact PIL: 2 launches)", lambda t: PO.compress_with_format(t, "jpeg", 77.0), 2 * N(x192), x=x192)

if args.json:
    json.dump({"hbm_peak_gbs": peak, "fma_peak_tflops": FMA_PEAK, "rows": rows}, open(args.json, "w"), indent=1)
