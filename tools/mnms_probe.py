"""Times Matrix-NMS at BASELINE config 4 (16 x 500 masks 200x336), from fp32 masks and from packed masks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from detectron2_tensorflow_b200.layers import matrix_nms
from detectron2_tensorflow_b200.modeling import solo_mask_encode
from detectron2_tensorflow_b200.utils import synthetic as syn
dev = torch.device("cuda", 0)
B, n, H, W = 16, 500, 200, 336
m, c, s = syn.solo_masks(n, hw=(H, W), seed=7)
masks = torch.from_numpy(m).to(dev)[None].repeat(B, 1, 1, 1).contiguous()
classes = torch.from_numpy(c).to(dev)[None].repeat(B, 1).contiguous()
scores = torch.from_numpy(s).to(dev)[None].repeat(B, 1).contiguous()


def med(fn, k=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(k):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = {"fp32_masks_ms": med(lambda: matrix_nms(masks, classes, scores))}
packed, sums, _ = solo_mask_encode(masks * 8 - 4, 0.5)
out["packed_masks_ms"] = med(lambda: matrix_nms(None, classes, scores, sum_masks=sums, packed_masks=packed, mask_hw=H * W))
out["fp32_masks_ms_again"] = med(lambda: matrix_nms(masks, classes, scores))
out["mask_read_frac_of_6525GBps"] = B * n * H * W * 4 / (out["fp32_masks_ms_again"] * 1e-3) / 6525.2e9
print(json.dumps(out))
