"""Host-side jitter of the end-to-end loop: per-step host timestamps over a long run, then the distribution of 20-step
window durations (what a --steps 20 run samples once), with and without the nvidia-smi clock sampler running beside it.
    python profiles/e2e_jitter.py [steps]"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from trainner_redux_b200.prefetch import CUDAPrefetcher, CUDAReadback  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
wl = bench.Workload("c2", 1)
arm = bench.Arm(wl, torch.device("cuda:0"), 0, 1, None)
dev = arm.dev
batches = [{"gt": (d["gt"] * 255.0).round().clamp(0, 255).to(torch.uint8).pin_memory(),
            "kernel_params": torch.stack([torch.as_tensor(p, dtype=torch.float64) for p in d["kernel_params"]]).pin_memory()} for d in arm.host]
rb = CUDAReadback(dev)
pf = CUDAPrefetcher((batches[i % 4] for i in range(10 ** 9)), device=dev, slots=2)


def run(n):
    ts = []
    for _ in range(n):
        b = pf.next()
        arm.feed.feed_data(b, plan=arm.plan())
        rb.read(arm.feed.lq)
        if len(ts) % 8 == 7:
            rb.wait()  # bound the host's run-ahead to a few steps, as a short timed region does
        ts.append(time.perf_counter())
    rb.wait()
    torch.cuda.synchronize()
    return ts


run(50)
for label, sampler in (("no sampler", False), ("nvidia-smi -lms 50", True), ("no sampler", False), ("nvidia-smi -lms 50", True)):
    s = bench.ClockSampler(0) if sampler else None
    time.sleep(0.3)
    ts = run(steps)
    if s:
        s.stop()
    w = sorted((ts[i + 20] - ts[i]) * 1e3 for i in range(0, len(ts) - 20, 5))
    d = sorted((ts[i + 1] - ts[i]) * 1e3 for i in range(len(ts) - 1))
    print(f"{label:20s} 20-step windows ms: median {statistics.median(w):.2f}  p90 {w[int(.9 * len(w))]:.2f}  p99 {w[int(.99 * len(w))]:.2f}  max {w[-1]:.2f}"
          f" | single steps ms: median {statistics.median(d):.3f} p99 {d[int(.99 * len(d))]:.3f} max {d[-1]:.3f}  steps > 1 ms: {sum(x > 1 for x in d)}", flush=True)
