"""Where the HOST time of the end-to-end loop goes: cProfile around bench.py's e2e loop (uint8-GT feed or fp32 feed).
    python profiles/e2e_host_profile.py [u8|f32] [steps]"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

u8 = (sys.argv[1] if len(sys.argv) > 1 else "u8") == "u8"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
wl = bench.Workload("c2", 1)
arm = bench.Arm(wl, torch.device("cuda:0"), 0, 1, None)
pr = cProfile.Profile()
pr.enable()
v, h2d, d2h = arm.e2e(steps, 8, u8)
pr.disable()
print(f"e2e {'u8' if u8 else 'f32'}: {v:.0f} pairs/s ({64e3 / v:.3f} ms per step), h2d {h2d} B/step")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
