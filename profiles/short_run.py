import sys, time
sys.path.insert(0, "/root/repo")
import torch, bench
wl = bench.Workload("c2", 1)
arm = bench.Arm(wl, torch.device("cuda:0"), 0, 1, None)
for i in range(12): arm.step(i)
arm.barrier()
arm.timed(8, 4)
for K, lanes in ((20, 4), (20, 4), (20, 1), (20, 1), (200, 4), (2000, 4), (20, 4), (20,4)):
    t0 = time.perf_counter()
    ms = arm.timed(K, lanes)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"K={K} lanes={lanes}: event {ms:.3f} ms ({64e3*K/ms/1e3:.0f} k pairs/s), wall incl. sync {wall:.3f} ms", flush=True)
# host issue time only (no GPU wait): time the python loop
import statistics
ts = []
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20): arm.step(i)
    ts.append((time.perf_counter() - t0) / 20 * 1e3)
    torch.cuda.synchronize()
print("host issue ms/step over 20-step bursts:", [round(t, 4) for t in ts])
