"""Quick device-side timing of d2b_roi_align_multilevel at the Mask R-CNN batch-16 config."""
import argparse
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from detectron2_tensorflow_b200.modeling import ROIPooler
from detectron2_tensorflow_b200.structures import BoxList, SparseBoxList
from detectron2_tensorflow_b200.utils import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=16)
ap.add_argument("--rois", type=int, default=1000)
ap.add_argument("--out", type=int, default=7)
ap.add_argument("--sr", type=int, default=0)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--bf16", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda", 0)
N, R, C = args.images, args.rois, 256
g = torch.Generator(device=dev).manual_seed(0)
feats = []
for s in syn.FPN_STRIDES:
    h, w = syn.level_hw(s)
    f = torch.randn((N, h, w, C), device=dev, generator=g)
    feats.append(f.to(torch.bfloat16) if args.bf16 else f)
boxes, idx = syn.rois(N, R, seed=1)
inst = SparseBoxList(torch.from_numpy(idx).to(dev), BoxList(torch.from_numpy(boxes).to(dev)), (N, R))
pooler = ROIPooler(args.out, [1 / 4., 1 / 8., 1 / 16., 1 / 32.], args.sr, "ROIAlignV2")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    out = pooler(feats, inst)
torch.cuda.synchronize()
times = []
for _ in range(args.iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pooler(feats, inst)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = float(np.median(times))
M = N * R
esz = 2 if args.bf16 else 4
feat_bytes = sum(f.numel() for f in feats) * esz
gathered = M * args.out * args.out * max(args.sr, 1) ** 2 * 4 * C * esz
alg = M * args.out * args.out * C * esz + min(feat_bytes, gathered) + M * 24
print(json.dumps({"rois": M, "out": args.out, "sr": args.sr, "bf16": args.bf16, "ms": ms, "min_ms": min(times),
                  "rois_per_s": M / ms * 1e3, "alg_GB": alg / 1e9, "alg_GBps": alg / ms / 1e6,
                  "frac_of_6525": alg / ms / 1e6 / 6525.2, "level_counts": pooler.last_level_counts.tolist()}))
