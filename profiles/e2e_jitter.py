"""Is the spread of bench.py's 20-step e2e number a start-up effect or random hiccups?  The e2e loop of bench.py with ONE
prefetcher, timed in consecutive 20-step runs (each bracketed by a read-back wait + device synchronise, like the timed
region): the first runs after set-up against the steady distribution.    python profiles/e2e_jitter.py [lanes]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from trainner_redux_b200.prefetch import CUDAPrefetcher, CUDAReadback  # noqa: E402

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wl = bench.Workload("c2", 1)
arm = bench.Arm(wl, torch.device("cuda:0"), 0, 1, None)
dev = arm.dev
for i in range(12):
    arm.step(i)
torch.cuda.synchronize()
batches = [{"gt": (d["gt"] * 255.0).round().clamp(0, 255).to(torch.uint8).pin_memory(),
            "kernel_params": torch.stack([torch.as_tensor(p, dtype=torch.float64) for p in d["kernel_params"]]).pin_memory()} for d in arm.host]
rb = CUDAReadback(dev)
n_slots = max(2, lanes)
pf = CUDAPrefetcher((batches[i % 4] for i in range(10 ** 9)), device=dev, slots=n_slots)
main = torch.cuda.current_stream()
lane = [torch.cuda.Stream(dev) for _ in range(lanes)] if lanes > 1 else [main]
t = 0


def run(n):
    global t
    torch.cuda.set_stream(lane[t % lanes])
    for _ in range(n):
        b = pf.next()
        arm.feed.feed_data(b, plan=arm.plan())
        rb.read(arm.feed.lq)
        t += 1
        if lanes > 1:
            torch.cuda.set_stream(lane[t % lanes])
    torch.cuda.set_stream(main)
    rb.wait()
    torch.cuda.synchronize()


run(2 * n_slots)  # what bench.py does before its timed region
sampler = bench.ClockSampler(0) if os.environ.get("OTF_JITTER_SAMPLER") else None  # nvidia-smi -lms 50 beside the loop, as in bench.py
ms = []
for _ in range(60):
    t0 = time.perf_counter()
    run(20)
    ms.append((time.perf_counter() - t0) * 1e3)
if sampler:
    print("clock sampler:", sampler.stop())
print(f"lanes {lanes}: first five 20-step runs after set-up: {[round(x, 2) for x in ms[:5]]} ms")
s = sorted(ms[5:])
print(f"next 55 runs: min {s[0]:.2f}  median {s[len(s) // 2]:.2f}  p90 {s[int(.9 * len(s))]:.2f}  max {s[-1]:.2f} ms"
      f"  (= {64 * 20 / s[len(s) // 2]:.0f} k pairs/s at the median, {64 * 20 / s[-1]:.0f} k at the slowest)")
