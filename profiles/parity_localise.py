#!/usr/bin/env python
"""Where do end-to-end LQ mismatches come from?  Runs the full-size chain cases of tests/chain_cases.py through the
oracle (with taps) and through the CUDA path stage by stage (RealESRGANFeed.collect_taps), and prints for every stage
(a) max-abs / fraction within 1 LSB of the FREE-RUNNING CUDA chain against the oracle and (b) the same with the stage
fed the ORACLE's input (isolates the stage's own error from propagated rounding flips).  Test infrastructure: imports
the oracle as the checker.

    python profiles/parity_localise.py [c2|c3] [gaussian|poisson] [resize_first|jpeg_first] [natural|uniform]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from chain_cases import lsb_fraction, make_case, run_oracle  # noqa: E402
from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed  # noqa: E402


def main() -> None:
    args = sys.argv[1:] + ["c2", "gaussian", "resize_first", "uniform"][len(sys.argv) - 1:]
    dev = torch.device("cuda:0")
    case = make_case(*args[:4], seed=int(os.environ.get("SEED", "0")))
    taps: dict = {}
    want_gt, want_lq, noise = run_oracle(case, taps)
    feed = RealESRGANFeed(OTFOptions(scale=case["scale"], gt_size=case["crop"]), device=dev, use_pool=False)
    feed.collect_taps = {}
    inject = {k: v.to(dev) for k, v in noise.items()}
    feed.feed_data({k: case[k] for k in ("gt", "kernel1", "kernel2", "sinc_kernel")}, plan=case["plan"], inject=inject)
    torch.cuda.synchronize()
    rows = []
    names = {"jpeg2+round": "lq_full", "round": "lq_full"}
    for name, got in feed.collect_taps.items():
        want = taps[names.get(name, name)]
        d = (got.cpu() - want).abs()
        frac, worst = lsb_fraction(got, want)
        rows.append({"stage": name, "free_running_maxabs": d.max().item(), "free_running_frac_1lsb": frac,
                     "n_gt_1e-5": int((d > 1e-5).sum().item()), "numel": d.numel()})
    frac, worst = lsb_fraction(feed.lq, want_lq)
    # the reference against ITSELF: the same chain of ATen calls on the CUDA device (what the reference executes in
    # training) vs on the CPU, same inputs, same injected fields / counts — how far two builds of the reference's own
    # arithmetic drift apart through the rounding cliffs
    self_frac = self_worst = None
    try:
        from oracle import otf_oracle as O

        def mv(v):
            if torch.is_tensor(v):
                return v.to(dev)
            if isinstance(v, dict):
                return {k: mv(x) for k, x in v.items()}
            return v

        _, lq_cuda = O.run_chain_b(case["gt"].to(dev), case["kernel1"].to(dev), case["kernel2"].to(dev), case["sinc_kernel"].to(dev),
                                   mv(case["plan"]), mv(noise))
        self_frac, self_worst = lsb_fraction(lq_cuda, want_lq)
        mine_vs_cuda = lsb_fraction(feed.lq, lq_cuda)
    except Exception as e:  # noqa: BLE001
        mine_vs_cuda = (None, None)
        self_frac = f"failed: {type(e).__name__}: {e}"
    out = {"case": args[:4], "final_frac_1lsb": frac, "final_max_lsb": worst,
           "reference_cuda_vs_reference_cpu_frac_1lsb": self_frac, "reference_cuda_vs_reference_cpu_max_lsb": self_worst,
           "this_repo_vs_reference_cuda_frac_1lsb": mine_vs_cuda[0], "stages": rows}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
