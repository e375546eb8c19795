"""Break down MaskRCNNPostBackbone.run_host: upload only / upload+kernels / full, per chunk size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from detectron2_tensorflow_b200 import engine as E

dev = torch.device("cuda", 0)
host = bench.make_host_inputs(16)
hx = bench.to_torch(host, dev=None, pin=True)
eng = bench.make_engine()
print("pinned:", hx["feats"][0].is_pinned(), hx["feats"][0][2:4].is_pinned(), hx["scores"][100:200].is_pinned())

def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

for chunk in (1, 2, 4, 8, 16):
    full = timed(lambda: eng.run_host(hx, dev, chunk_images=chunk))
    print({"chunk": chunk, "run_host_ms": full})

# upload only, same chunking, two streams
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
def upload_only(chunk=2):
    keep = []
    for ci, b in enumerate(range(0, 16, chunk)):
        e = b + chunk
        with torch.cuda.stream(streams[ci % 2]):
            for k in E.PER_IMAGE_KEYS:
                v = hx[k]
                keep.append([t[b:e].to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v[b:e].to(dev, non_blocking=True))
            for k in E.PER_ROI_KEYS:
                keep.append(hx[k][b * 1000:e * 1000].to(dev, non_blocking=True))
    torch.cuda.synchronize()
print({"upload_only_ms": timed(upload_only)})
t0 = time.perf_counter(); upload_only(); cpu = (time.perf_counter() - t0) * 1e3
print({"upload_only_incl_sync_single_ms": cpu})
# device step on 2-image chunk, kernels only
xd = bench.to_torch({k: ([a[:2] for a in v] if isinstance(v, list) and k != "anchors" else (v if k == "anchors" else v[:2] if k in ("shapes",) else v[:2000])) for k, v in host.items()}, dev=dev)
print({"kernels_2img_ms": timed(lambda: eng(xd))})
t0 = time.perf_counter()
for _ in range(20): eng(xd)
print({"python_enqueue_2img_ms": (time.perf_counter() - t0) / 20 * 1e3})
torch.cuda.synchronize()
