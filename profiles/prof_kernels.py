"""Launch each kernel of the path a few times on the bench shapes (B=64) — the command ncu wraps.
    python profiles/prof_kernels.py [stage ...]     stages: blur1 resize1 resize_up noise1 jpeg1 libjpeg blur2 sinc poisson usm all
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trainner_redux_b200 as T  # noqa: E402
from trainner_redux_b200 import synthetic as S  # noqa: E402
from trainner_redux_b200.kernels import synthesize_kernels  # noqa: E402
from trainner_redux_b200 import degradations as D  # noqa: E402

stages = sys.argv[1:] or ["all"]
want = lambda s: "all" in stages or s in stages
dev = torch.device("cuda:0")
B = 64
gt = S.synth_gt(B, 256, 256, "uniform", seed=1).to(dev)
k1, k2, sk = (synthesize_kernels(p, dev) for p in S.synth_kernel_params(B, 10))
x192 = S.synth_gt(B, 192, 192, "uniform", seed=2).to(dev)
x64 = S.synth_gt(B, 64, 64, "uniform", seed=3).to(dev)
sigma = (torch.rand(B) * 29 + 1).to(dev)
gray = (torch.rand(B) < 0.4).float().to(dev)
q = (torch.rand(B) * 65 + 30).to(dev)
jp = T.DiffJPEG(differentiable=False)
for it in range(3):
    if want("blur1"):
        T.filter2d(gt, k1)
    if want("resize1"):
        T.resize_pt(gt, "bicubic", scale_factor=0.75)
        T.resize_pt(x192, "bilinear", size=(64, 64))
        T.resize_pt(x64, "area", size=(64, 64))
    if want("resize_ne"):
        T.resize_pt(gt, "nearest-exact", scale_factor=0.75)
    if want("resize_up"):
        T.resize_pt(gt, "bilinear", scale_factor=1.5)
    if want("noise1"):
        D.add_gaussian_noise_pt(x192, sigma, gray)
    if want("jpeg1"):
        jp(x192, quality=q.clone(), _clamp_in=True)
        jp(x64, quality=q.clone(), _clamp_in=True, _round8=True)
    if want("libjpeg"):
        from trainner_redux_b200 import paragon_otf as PO
        PO.compress_with_format(x192, "jpeg", 77.0)
        PO.compress_with_format(x64, "jpeg", 77.0)
    if want("blur2"):
        T.filter2d(x192, k2)
    if want("sinc"):
        T.filter2d(x64, sk)
    if want("poisson"):
        D.add_poisson_noise_pt(x192, sigma / 10, True, False, gray)
    if want("g1"):
        from trainner_redux_b200 import _lib as L
        for (x, oh, mode_id) in ((gt, 192, L.RESIZE_BICUBIC_AA), (x192, 64, L.RESIZE_BILINEAR_AA)):
            b, c, h, w = x.shape
            tab = D.pinned_resize_table(dev, h, w, oh, oh, mode_id)
            nb = L.load().otf_resize_workspace_bytes(h, w, oh, oh, mode_id)
            o = torch.empty(b, c, oh, oh, device=dev)
            L.call("otf_resize_gauss_f32", L.ptr(x), b, c, h, w, L.ptr(o), oh, oh, mode_id, 1, L.ptr(tab), nb, 1, L.ptr(sigma), L.ptr(gray),
                   7, 1, None, L.NOISE_CLIP, L.stream())
        go, lo = torch.empty(B, 3, 224, 224, device=dev), torch.empty(B, 3, 56, 56, device=dev)
        L.call("otf_diffjpeg_crop_pair_f32", L.ptr(x64), B, 64, 64, L.ptr(q), 0.0, 1, 0, 1, L.ptr(gt), 256, 256, 4, 4, None, 56, 4,
               L.ptr(go), L.ptr(lo), L.stream())
    if want("usm"):
        T.USMSharp().to(dev)(gt)
torch.cuda.synchronize()
print("ok")
