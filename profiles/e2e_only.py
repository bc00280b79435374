"""The end-to-end loop of bench.py alone (uint8-GT feed and fp32 feed), plus a variant whose host batches already live
on the device (the prefetcher passes them through): the same host work and GPU work without the PCIe upload — tells a
PCIe-bound loop from a host-bound one.    python profiles/e2e_only.py [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
wl = bench.Workload("c2", 1)
arm = bench.Arm(wl, torch.device("cuda:0"), 0, 1, None)
for u8, lanes in ((True, 1), (True, 2), (True, 4), (False, 1)):
    v, h2d, d2h = arm.e2e(steps, 8, u8, lanes)
    print("slots", bench.E2E_SLOTS, "lanes", lanes, "u8" if u8 else "f32", round(v), "pairs/s", round(64e3 / v, 4), "ms/step", flush=True)

# no-upload variant of the uint8 loop
from trainner_redux_b200.prefetch import CUDAPrefetcher, CUDAReadback  # noqa: E402

dev = arm.dev
batches = [{"gt": (d["gt"] * 255.0).round().clamp(0, 255).to(torch.uint8).to(dev),
            "kernel_params": torch.stack([torch.as_tensor(p, dtype=torch.float64) for p in d["kernel_params"]]).to(dev)} for d in arm.host]
rb = CUDAReadback(dev)


def run(n, read=True, feed=True):
    pf = CUDAPrefetcher((batches[i % len(batches)] for i in range(n)), device=dev, slots=2)
    b = pf.next()
    while b is not None:
        if feed:
            arm.feed.feed_data(b, plan=arm.plan())
        else:
            arm.plan()
        if read:
            rb.read(arm.feed.lq)
        b = pf.next()
    rb.wait()
    torch.cuda.synchronize()


run(16)
for name, kw in (("device-resident u8 + readback", {}), ("device-resident u8, no readback", {"read": False}),
                 ("host only: prefetcher + draw_plan (no feed_data, no readback)", {"read": False, "feed": False})):
    t0 = time.perf_counter()
    run(steps, **kw)
    dt = time.perf_counter() - t0
    print(name, round(64 * steps / dt), "pairs/s", round(dt / steps * 1e3, 4), "ms/step", flush=True)
