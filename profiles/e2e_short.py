"""End-to-end loop at short K (the driver runs bench.py with --steps 20): throughput against the number of untimed
warm-up steps in front of the timed region — separates start-up transients (clock ramp after the idle set-up phase,
first-use costs) from the steady state.    python profiles/e2e_short.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

wl = bench.Workload("c2", 1)
arm = bench.Arm(wl, torch.device("cuda:0"), 0, 1, None)
for i in range(12):
    arm.step(i)
for warm in (5, 50, 200, 800):
    for K in (20, 100):
        for lanes in (1, 4):
            v, _, _ = arm.e2e(K, warm, True, lanes)
            print(f"warm-up {warm:4d}  K={K:4d} lanes={lanes}  e2e {v / 1e3:6.1f} k pairs/s  ({64e3 / v:.4f} ms/step)", flush=True)
