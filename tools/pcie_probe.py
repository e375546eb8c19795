"""Pinned host<->device bandwidth on this box, for 1..N ranks at once (context for the `e2e` numbers of bench.py).

    python tools/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

Every rank copies 1 GiB between its own pinned host buffers and its GPU: H2D only, D2H only and both directions at
once (two streams), all ranks started together behind a barrier; rank 0 prints one JSON line with the per-rank and
aggregate GB/s.  It also times the `cudaHostAllocPortable | WriteCombined` staging variant for the H2D side."""
import ctypes
import json
import os
import time

import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)

n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def t(fn, reps=3):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return dt


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


res = {"h2d_GBps": n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9,
       "d2h_GBps": n / t(lambda: h2.copy_(d2, non_blocking=True)) / 1e9}
tb = t(both)
res["duplex_GBps_each_direction"] = n / tb / 1e9

# write-combined, portable pinned staging for the H2D side (cudaHostAlloc flags 1 | 4)
try:
    rt = ctypes.CDLL("libcudart.so")
    ptr = ctypes.c_void_p()
    if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n), ctypes.c_uint(1 | 4)) == 0:
        rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        st = torch.cuda.current_stream().cuda_stream
        res["h2d_write_combined_GBps"] = n / t(lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ptr,
                                                                         ctypes.c_size_t(n), 1, ctypes.c_void_p(st))) / 1e9
        rt.cudaFreeHost(ptr)
except OSError as e:  # no libcudart on the loader path
    res["h2d_write_combined_GBps"] = None

if world > 1:
    keys = sorted(k for k, v in res.items() if v is not None)
    x = torch.tensor([res[k] for k in keys], device=dev, dtype=torch.float64)
    allx = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(allx, x)
    if rank == 0:
        per_rank = {k: [float(a[i]) for a in allx] for i, k in enumerate(keys)}
        print(json.dumps({"ranks": world, "aggregate_GBps": {k: sum(v) for k, v in per_rank.items()},
                          "per_rank_GBps": per_rank, "cpus": len(os.sched_getaffinity(0))}))
    dist.destroy_process_group()
else:
    res["cpus"] = len(os.sched_getaffinity(0))
    print(json.dumps(res))
