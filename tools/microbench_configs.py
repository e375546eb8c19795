"""Device-side timing of the other BASELINE.json configs (they are parity-test cases, not bench lines):
  config 3  RetinaNet R50-FPN inference post-processing, batch 32
  config 4  SOLOv2 Matrix-NMS, 500 masks at 200x336, batch 16
  config 5  ROIAlign / NMS sweep, 256 .. 65536 ROIs / boxes
Prints one JSON line per case (CUDA events, median of `--iters`, 256 MB L2 flush between iterations)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from detectron2_tensorflow_b200.layers import batch_nms, matrix_nms
from detectron2_tensorflow_b200.modeling import ROIPooler, RetinaNetInference
from detectron2_tensorflow_b200.structures import BoxList, SparseBoxList
from detectron2_tensorflow_b200.utils import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--which", default="3,4,5")
args = ap.parse_args()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
HBM = 6525.2


def timeit(fn, iters=args.iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


which = set(args.which.split(","))
g = torch.Generator(device=dev).manual_seed(0)

if "3" in which:
    N, K = 32, 80
    anchors = [torch.from_numpy(a).to(dev) for a in syn.retinanet_anchors()]
    cls = [torch.randn((N, a.shape[0], K), device=dev, generator=g) * 1.5 - 4.6 for a in anchors]
    dl = [torch.randn((N, a.shape[0], 4), device=dev, generator=g) * 0.3 for a in anchors]
    head = RetinaNetInference(num_classes=K)
    ms = timeit(lambda: head.inference(cls, dl, anchors))
    nscores = sum(c.numel() for c in cls)
    print(json.dumps({"config": 3, "case": "RetinaNet post-processing N=32", "ms": ms, "images_per_s": N / ms * 1e3,
                      "class_scores_per_s": nscores / ms * 1e3, "score_bytes_GB": nscores * 4 / 1e9,
                      "single_read_GBps": nscores * 4 / ms / 1e6, "frac_hbm_single_read": nscores * 4 / ms / 1e6 / HBM}))

if "4" in which:
    B, n, H, W = 16, 500, 200, 336
    m, c, s = syn.solo_masks(n, hw=(H, W), seed=7)
    masks = torch.from_numpy(m).to(dev)[None].repeat(B, 1, 1, 1).contiguous()
    classes = torch.from_numpy(c).to(dev)[None].repeat(B, 1).contiguous()
    scores = torch.from_numpy(s).to(dev)[None].repeat(B, 1).contiguous()
    ms = timeit(lambda: matrix_nms(masks, classes, scores))
    by = masks.numel() * 4
    print(json.dumps({"config": 4, "case": "SOLOv2 matrix-NMS B=16 n=500 200x336", "ms": ms, "images_per_s": B / ms * 1e3,
                      "masks_per_s": B * n / ms * 1e3, "mask_read_GB": by / 1e9, "GBps": by / ms / 1e6,
                      "frac_hbm": by / ms / 1e6 / HBM}))

if "5" in which:
    N, C = 16, 256
    feats = [torch.randn((N,) + syn.level_hw(s) + (C,), device=dev, generator=g) for s in syn.FPN_STRIDES]
    fbytes = sum(f.numel() for f in feats) * 4
    pooler = ROIPooler(7, [1 / 4., 1 / 8., 1 / 16., 1 / 32.], 0, "ROIAlignV2")
    for M in (256, 1024, 4096, 16384, 65536):
        boxes, idx = syn.rois(N, M // N, seed=1)
        inst = SparseBoxList(torch.from_numpy(idx).to(dev), BoxList(torch.from_numpy(boxes).to(dev)), (N, M // N))
        ms = timeit(lambda: pooler(feats, inst))
        alg = M * 49 * C * 4 + min(fbytes, M * 49 * 4 * C * 4) + M * 24
        print(json.dumps({"config": 5, "case": f"ROIAlign 7x7 sweep M={M}", "ms": ms, "rois_per_s": M / ms * 1e3,
                          "alg_GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / HBM}))
    rng = np.random.default_rng(5)
    for n in (256, 1024, 4096, 16384, 65536):
        cy, cx = rng.uniform(0, 800, n), rng.uniform(0, 1333, n)
        h, w = rng.uniform(16, 300, n), rng.uniform(16, 300, n)
        b = torch.from_numpy(np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1).astype(np.float32)).to(dev)[None]
        sc = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(dev)[None]
        out = {}
        def run():
            out["k"], out["n"] = batch_nms(b, sc, n, axis=1, iou_threshold=0.7)
        ms = timeit(run, iters=5)
        print(json.dumps({"config": 5, "case": f"NMS sweep n={n} (one segment, thr 0.7, uncapped)", "ms": ms,
                          "boxes_per_s": n / ms * 1e3, "kept": int(out["n"][0])}))
