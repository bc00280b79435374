"""Shared builder for the full-size chain parity cases (tests/test_chain_fullsize_gpu.py, bench.py's parity leg):
BASELINE.json configs[1] (B=64 x 256^2 GT x4, crop 224/56) and configs[2] (B=32 x 512^2 GT x2, crop 480/240), every
stage of the classical chain on (SURVEY.md §8d "Config 2").  The oracle pass draws the random fields on the CPU
(Gaussian fields up front, Poisson counts from the oracle's own lambda) and records them so that the CUDA path can be
fed the very same fields — the north_star protocol: same inputs, same parameters, same injected noise."""

from __future__ import annotations

import torch

from oracle import otf_oracle as O

CONFIGS = {"c2": (64, 256, 4, 224), "c3": (32, 512, 2, 480)}


def make_case(config: str, noise_kind: str, final_order: str, gt_kind: str, seed: int = 0, batch: int | None = None) -> dict:
    b, size, scale, crop = CONFIGS[config]
    b = batch or b
    g = torch.Generator().manual_seed(1000 + seed)
    gt = O.synth_gt(b, size, size, gt_kind, seed=1234 + seed)
    k1, k2 = O.synth_blur_kernels(b, seed=seed), O.synth_blur_kernels(b, seed=50 + seed)
    sk = O.synth_sinc_or_pulse(b, seed=seed)

    def noise_stage(hi_sigma: float, hi_scale: float) -> dict:
        gray = (torch.rand(b, generator=g) < 0.4).float()
        if noise_kind == "gaussian":
            return {"kind": "gaussian", "sigma": torch.rand(b, generator=g) * (hi_sigma - 1) + 1, "gray": gray}
        return {"kind": "poisson", "scale": torch.rand(b, generator=g) * (hi_scale - 0.05) + 0.05, "gray": gray}

    plan = {
        "scale": scale, "gt_size": crop, "order": "classic", "blur1": True, "resize1": {"scale": 0.75, "mode": "bicubic"},
        "noise1": noise_stage(30, 3.0), "jpeg1": torch.rand(b, generator=g) * 65 + 30, "blur2": True,
        "resize2": {"scale": 1.0, "mode": "bilinear"}, "noise2": noise_stage(25, 2.5), "final_order": final_order,
        "resize3_mode": "area", "jpeg2": torch.rand(b, generator=g) * 65 + 30, "crop": (3, 5),
    }
    return {"b": b, "size": size, "scale": scale, "crop": crop, "gt": gt, "kernel1": k1, "kernel2": k2, "sinc_kernel": sk, "plan": plan,
            "seed": seed}


def run_oracle(case: dict, taps: dict | None = None, fields: dict | None = None) -> tuple[torch.Tensor, torch.Tensor, dict]:
    """Oracle pass: returns (gt_crop, lq_crop, noise) with ``noise`` holding every random field it consumed, keyed as
    RealESRGANFeed's ``inject`` expects; ``fields`` collects the finished noise fields (``noise1_field`` ...), the
    other thing ``inject`` accepts."""
    b, size, scale, plan = case["b"], case["size"], case["scale"], case["plan"]
    gen = torch.Generator().manual_seed(77 + case["seed"])
    noise: dict = {}
    h1 = round(size * plan["resize1"]["scale"])
    h2 = int(size / scale * plan["resize2"]["scale"])
    for key, hh in (("noise1", h1), ("noise2", h2)):
        if plan[key]["kind"] == "gaussian":
            noise[f"{key}_color"] = torch.randn(b, 3, hh, hh, generator=gen)
            noise[f"{key}_gray"] = torch.randn(hh, hh, generator=gen)
    calls: list[torch.Tensor] = []

    def rec(lam: torch.Tensor) -> torch.Tensor:
        c = torch.poisson(lam, generator=gen)
        calls.append(c)
        return c

    gt_c, lq_c = O.run_chain_b(case["gt"], case["kernel1"], case["kernel2"], case["sinc_kernel"], plan, noise, poisson_fn=rec, taps=taps,
                                 fields=fields)
    it = iter(calls)
    for key in ("noise1", "noise2"):
        st = plan[key]
        if st["kind"] == "poisson":  # gray first if any flag is set, then colour (degradations.py:787-806)
            if st["gray"].sum() > 0:
                noise[f"{key}_counts_gray"] = next(it)
            noise[f"{key}_counts_color"] = next(it)
    return gt_c, lq_c, noise


def lsb_fraction(got: torch.Tensor, want: torch.Tensor) -> tuple[float, float]:
    """(fraction of pixels within 1 LSB, max difference in LSB)."""
    diff = (got.detach().cpu().float() - want.detach().cpu().float()).abs()
    return (diff <= 1.0 / 255.0 + 1e-6).float().mean().item(), diff.max().item() * 255.0
