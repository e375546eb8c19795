"""Stage timings of SOLOv2Inference.inference at BASELINE config 4 scale (16 images, ~500 candidates)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectron2_tensorflow_b200.modeling import SOLOv2Inference, solo_upsample_masks
from detectron2_tensorflow_b200 import _native as nv
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
B, H, W, E = 16, 200, 336, 256
grids = (40, 36, 24, 16, 12)
probs = [torch.where(torch.rand((B, g_, g_, 80), device=dev, generator=g) < 0.0016,
                     torch.rand((B, g_, g_, 80), device=dev, generator=g) * 0.8 + 0.15, torch.zeros((), device=dev)) for g_ in grids]
kerns = [torch.randn((B, g_, g_, E), device=dev, generator=g) / 16 for g_ in grids]
yy = torch.arange(H, device=dev, dtype=torch.float32)[None, :, None, None]
xx = torch.arange(W, device=dev, dtype=torch.float32)[None, None, :, None]
cy = torch.rand((B, 1, 1, E), device=dev, generator=g) * H
cx = torch.rand((B, 1, 1, E), device=dev, generator=g) * W
sg = torch.rand((B, 1, 1, E), device=dev, generator=g) * 35 + 5
feat = torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sg * sg)).contiguous()  # bumps: one blob per candidate
feat[..., 0] = 1.0
for k_ in kerns:
    k_.mul_(0.01)
    pick = torch.randint(1, E, k_.shape[:-1] + (1,), device=dev, generator=g)
    k_.scatter_(-1, pick, 8.0)
    k_[..., 0] = -4.0
head = SOLOv2Inference(0.5, 500, "gaussian", 2.0, 0.05, 100, score_threshold=0.1, num_grids=grids, strides=(8, 8, 16, 32, 32),
                       max_candidates=int(sys.argv[1]) if len(sys.argv) > 1 else 1024)


def med(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


sc = torch.cat([p_.reshape(B, -1, 80) for p_ in probs], 1).contiguous()
kn = torch.cat([k_.reshape(B, -1, E) for k_ in kerns], 1).contiguous()
out = {"concat_ms": med(lambda: (torch.cat([p_.reshape(B, -1, 80) for p_ in probs], 1), torch.cat([k_.reshape(B, -1, E) for k_ in kerns], 1)))}
out["select_ms"] = med(lambda: head.select_candidates(sc, kn))
cand = head.select_candidates(sc, kn)
out["tail_ms"] = med(lambda: head.postprocess(None, cand["scores"], cand["classes"], cand["strides"], cand["counts"], return_masks=False,
                                              mask_features=feat, mask_kernels=cand["kernels"]))
tail = head.postprocess(None, cand["scores"], cand["classes"], cand["strides"], cand["counts"], return_masks=False,
                        mask_features=feat, mask_kernels=cand["kernels"])
out["upsample_ms"] = med(lambda: solo_upsample_masks(tail["packed_masks"], (H, W), (800, 1333), 0.5, False))
out["inference_ms"] = med(lambda: head.inference(probs, kerns, feat, (800, 1333)))
out["candidates"] = float(cand["counts"].float().mean())
res = head.inference(probs, kerns, feat, (800, 1333))
out["detections"] = float(res["num"].float().mean())
out["coverage"] = float(res["pred_masks"].float().mean())
print(json.dumps(out))
