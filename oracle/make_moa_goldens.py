"""Generate tests/golden/moa_goldens.npz by running the REFERENCE's MoA batch augment.

Run in the build container only (needs /root/reference):

    python -m oracle.make_moa_goldens

For every augmentation of traiNNer/ops/batchaug.py (mixup, cutmix, resizemix, cutblur, downup, up)
and a few shapes/seeds this script seeds the reference's three global generators (`random`,
`RNG._rng`, torch's default generator), runs the reference's `batch_aug`, runs the oracle
(oracle/batchaug_oracle.py) from identically seeded explicit generators, asserts the two agree
bit for bit — outputs AND the state the generators are left in — and stores inputs, the plan
(ratio, permutation, box, samplers) and the reference outputs.
"""

from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import batchaug_oracle as BO  # noqa: E402
from oracle import otf_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "moa_goldens.npz")

# (tag, batch, gt_h, gt_w, scale); `up` and resizemix need square inputs (see the oracle's header)
SHAPES = [("sq4", 3, 48, 48, 4), ("sq2", 4, 40, 40, 2), ("rect2", 3, 32, 48, 2)]
CASES = {
    "mixup": [("sq4", 1), ("rect2", 2)],
    "cutmix": [("sq4", 3), ("sq2", 4), ("rect2", 5)],
    "resizemix": [("sq4", 6), ("sq2", 7)],
    "cutblur": [("sq4", 8), ("sq2", 9), ("rect2", 10)],
    "downup": [("sq4", 11), ("sq2", 12), ("rect2", 13), ("sq2", 14), ("sq2", 15), ("sq2", 20), ("sq4", 37)],  # 20, 37: both draws nearest -> redraw branch
    "up": [("sq4", 16), ("sq2", 17), ("sq2", 18)],
}
SAMPLER_ID = {"bicubic": 0, "bilinear": 1, "nearest-exact": 2}


def main() -> None:
    torch.set_num_threads(1)
    R = load_reference()
    from traiNNer.ops import batchaug as ref_ba  # noqa: PLC0415

    G: dict[str, np.ndarray] = {}
    inputs = {}
    for tag, b, h, w, s in SHAPES:
        gt = O.synth_gt(b, h, w, "natural", seed=100 + h + w)
        lq = torch.nn.functional.interpolate(gt, size=(h // s, w // s), mode="area")
        lq = (lq + 0.05 * torch.randn(lq.shape, generator=torch.Generator().manual_seed(h * w))).clamp(0, 1)
        inputs[tag] = (gt, lq, s)
        G[f"in_{tag}_gt"], G[f"in_{tag}_lq"], G[f"in_{tag}_scale"] = gt.numpy(), lq.numpy(), np.array(s)

    names = []
    for aug, cases in CASES.items():
        for tag, seed in cases:
            gt, lq, s = inputs[tag]
            # the reference, from its global generators
            random.seed(seed)
            R.RNG._rng = np.random.default_rng(seed)
            torch.manual_seed(seed)
            r_gt, r_lq = ref_ba.batch_aug(gt.clone(), lq.clone(), s, [aug, "none"], [1.0, 0.0], False, 0)
            ref_state = (random.random(), float(R.RNG._rng.random()), float(torch.rand(1)))
            # the oracle, from explicit generators in the same state
            py, nprng, tgen = random.Random(seed), np.random.default_rng(seed), torch.Generator().manual_seed(seed)
            o_gt, o_lq, plan = BO.batch_aug(gt.clone(), lq.clone(), s, [aug, "none"], [1.0, 0.0], py, nprng, tgen)
            ora_state = (py.random(), float(nprng.random()), float(torch.rand(1, generator=tgen)))
            key = f"{aug}_{tag}_{seed}"
            if not (torch.equal(r_gt.contiguous(), o_gt.contiguous()) and torch.equal(r_lq.contiguous(), o_lq.contiguous())):
                raise SystemExit(f"oracle != reference for {key}")
            if ref_state != ora_state:
                raise SystemExit(f"oracle consumed different random numbers than the reference for {key}")
            names.append(key)
            if aug in ("mixup", "cutmix", "resizemix", "up"):  # the others return gt untouched
                G[f"{key}_gt"] = r_gt.contiguous().numpy()
            else:
                assert torch.equal(r_gt, gt)
            G[f"{key}_lq"] = r_lq.contiguous().numpy()
            G[f"{key}_lam"] = np.array(plan.get("lam", plan.get("factor", 0.0)), dtype=np.float64)
            G[f"{key}_perm"] = plan["perm"].numpy() if "perm" in plan else np.zeros(0, dtype=np.int64)
            G[f"{key}_box"] = np.array(plan.get("box", (0, 0, 0, 0)), dtype=np.int64)
            G[f"{key}_samplers"] = np.array([SAMPLER_ID[plan[k][0]] if k in plan else -1 for k in ("down", "up", "lq_up")])
            print(f"{key:24s} ok  plan: " + ", ".join(f"{k}={v}" for k, v in plan.items() if k not in ("aug", "scale", "perm")))
    G["cases"] = np.array(names)
    # error behaviour the reference defines (batchaug.py:84-89, :109)
    for bad, exc in ((lambda: ref_ba.batch_aug(gt[:1], lq[:1], 2, ["none"], [1.0], False, 0), ValueError),
                     (lambda: ref_ba.batch_aug(gt, lq, 2, ["none", "mixup"], [1.0], False, 0), ValueError),
                     (lambda: ref_ba.batch_aug(gt, lq, 2, ["bogus"], [1.0], False, 0), ValueError),
                     (lambda: ref_ba.cutmix(gt, lq, 3), ValueError)):
        try:
            bad()
        except exc:
            continue
        raise SystemExit("reference did not raise where the oracle expects it to")
    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
