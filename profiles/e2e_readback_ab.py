"""The end-to-end loop of bench.py on every rank with the LQ read back as fp32 (the bench's form), as uint8 (CUDAReadback.read(as_u8=True)), or not at all:
which part of the N = 8 figure is the read-back sharing the host path with the uploads.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 profiles/e2e_readback_ab.py [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
wl = bench.Workload("c2", world)
arm = bench.Arm(wl, dev, rank, world, dist)
out = {}
for name in ("f32", "none", "u8", "f32", "u8"):
    v, h2d, d2h = arm.e2e(steps, 8, True, 4, readback=name)
    out.setdefault(name, []).append(round(v))
if rank == 0:
    print(json.dumps({"n_gpus": world, "steps": steps, "e2e_pairs_per_s_by_readback": out}))
if world > 1:
    dist.destroy_process_group()
