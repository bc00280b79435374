"""Host->device bandwidth per rank with N ranks uploading at once: default pinned memory vs write-combined pinned memory
(cudaHostAllocWriteCombined: the GPU's reads do not snoop the CPU caches).  The end-to-end number of bench.py at N = 8 is
bound by this figure (no collective).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/pcie_wc.py
"""
import ctypes as C
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
torch.zeros(1, device="cuda")
rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaFreeHost.argtypes = [C.c_void_p]

NB = 12_582_912  # one step's uint8 GT: 64 x 3 x 256 x 256
REPS, SLOTS = 200, 4
dev = [torch.empty(NB, dtype=torch.uint8, device="cuda") for _ in range(SLOTS)]
out = {}
for name, flags in (("pinned", 0), ("write_combined", 4), ("pinned_again", 0)):
    host = []
    for _ in range(SLOTS):
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), NB, flags) == 0
        C.memset(p, 7, NB)  # first touch on this rank's CPUs
        host.append(p)
    st = torch.cuda.current_stream().cuda_stream
    for i in range(8):
        rt.cudaMemcpyAsync(dev[i % SLOTS].data_ptr(), host[i % SLOTS], NB, 1, st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REPS):
        rt.cudaMemcpyAsync(dev[i % SLOTS].data_ptr(), host[i % SLOTS], NB, 1, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    gbs = NB * REPS / ms / 1e6
    if world > 1:
        t = torch.tensor([gbs], dtype=torch.float64)
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        per = [round(float(x), 2) for x in lst]
    else:
        per = [round(gbs, 2)]
    out[name] = {"per_rank_gbs": per, "aggregate_gbs": round(sum(per), 1), "min": min(per)}
    for p in host:
        rt.cudaFreeHost(p)
    if world > 1:
        dist.barrier()
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_copy": NB, "copies": REPS, "h2d": out}))
if world > 1:
    dist.destroy_process_group()
