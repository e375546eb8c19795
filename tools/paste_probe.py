"""Times the mask paste-back (d2b_paste_masks) at config-2 size: 16 images x 100 detections, 28x28 masks -> 800x1333."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from detectron2_tensorflow_b200.structures.mask_ops import reframe_box_masks_to_image_masks
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
M, H, W = 1600, 800, 1333
masks = torch.from_numpy(rng.random((M, 28, 28)).astype(np.float32)).to(dev)
cy, cx = rng.uniform(0, H, M), rng.uniform(0, W, M)
h = np.exp(rng.uniform(np.log(20), np.log(500), M)); w = np.exp(rng.uniform(np.log(20), np.log(500), M))
boxes = np.stack([np.clip(cy - h / 2, 0, H), np.clip(cx - w / 2, 0, W), np.clip(cy + h / 2, 0, H), np.clip(cx + w / 2, 0, W)], 1)
boxes = torch.from_numpy(boxes.astype(np.float32)).to(dev)
once = len(sys.argv) > 1 and sys.argv[1] == "--once"
fn = lambda: reframe_box_masks_to_image_masks(masks, boxes, (H, W))
for _ in range(3):
    out = fn()
torch.cuda.synchronize()
if once:
    sys.exit(0)
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = fn(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
print(json.dumps({"paste_masks_ms": ms, "written_GBps": M * H * W / ms / 1e6, "coverage": float(out.float().mean().item())}))
