"""SOLOv2 dynamic conv + mask stage at BASELINE config 4 shapes (16 images x 500 candidates x 200x336, E=256):
the fused tcgen05 kernel (d2b_solo_dynamic_masks) next to the library route it replaces
(cuBLAS fp32 bmm -> 2.15 GB of logits -> d2b_solo_mask_encode).  One JSON line per variant."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detectron2_tensorflow_b200.modeling import (SOLOv2Inference, solo_dynamic_masks, solo_mask_encode,  # noqa: E402
                                                 solo_upsample_masks)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--n", type=int, default=500)
ap.add_argument("--hw", type=int, nargs=2, default=[200, 336])
ap.add_argument("--channels", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, n, (H, W), E = args.batch, args.n, args.hw, args.channels
g = torch.Generator().manual_seed(0)
feat = torch.randn((B, H, W, E), generator=g).to(dev)
kern = (torch.randn((B, n, E), generator=g) / 16).to(dev)
flops = 2.0 * B * n * H * W * E


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def library_route(tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    logits = torch.bmm(kern, feat.reshape(B, H * W, E).transpose(1, 2)).reshape(B, n, H, W)
    return solo_mask_encode(logits)


for name, fn in (("fused_tcgen05_3xtf32", lambda: solo_dynamic_masks(feat, kern)),
                 ("cublas_fp32_bmm_plus_encode", lambda: library_route(False)),
                 ("cublas_tf32_bmm_plus_encode", lambda: library_route(True))):
    med, best = timed(fn)
    print(json.dumps({"variant": name, "batch": B, "n": n, "hw": [H, W], "channels": E, "ms": med, "min_ms": best,
                      "useful_tflops": flops / med / 1e9, "tensor_tflops_issued": (3 if "fused" in name else 1) * flops / med / 1e9,
                      "logits_bytes_avoided": 4 * B * n * H * W if "fused" in name else 0}))
# sustained: 300 launches back to back with the SM clock sampled during the run (a 1 kW part under tensor load may sit
# below its maximum clock: the tensor roofline in cycles is what the kernel is judged on)
from bench import ClockSampler  # noqa: E402
cs = ClockSampler(0)
cs.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(300):
    solo_dynamic_masks(feat, kern)
e1.record()
torch.cuda.synchronize()
clk = cs.summary()
ms = e0.elapsed_time(e1) / 300
tiles_per_sm = B * ((n + 127) // 128) * ((H * W + 255) // 256) / 148.0
mma_cycles = tiles_per_sm * (E // 8) * 3 * 128  # 3 tf32 MMAs of 128 x 256 x 8 per k-step, 128 cycles each at 4096 flop/clk/SM
print(json.dumps({"variant": "fused_sustained_300_launches", "ms": ms, "issued_tflops": 3 * flops / ms / 1e9, "clocks": clk,
                  "mma_cycles_per_sm": mma_cycles,
                  "tensor_pipe_fraction_at_sampled_clock": mma_cycles / (ms * 1e-3 * clk["sm_mhz"] * 1e6) if clk["sm_mhz"] else None}))

# the whole inference tail (solo_v2.py:499-558): conv -> mask stage -> filter -> scoring -> top-k -> Matrix-NMS -> pad
head = SOLOv2Inference(0.5, 500, "gaussian", 2.0, 0.05, 100)
sc = torch.rand((B, n), generator=g).to(dev) * 0.9 + 0.1
cl = torch.randint(0, 80, (B, n), generator=g).to(dev)
stv = torch.full((B, n), 8.0, device=dev)


def tail_library():
    torch.backends.cuda.matmul.allow_tf32 = False
    logits = torch.bmm(kern, feat.reshape(B, H * W, E).transpose(1, 2)).reshape(B, n, H, W)
    return head.postprocess(logits, sc, cl, stv, return_masks=False)


for name, fn in (("tail_from_features_fused", lambda: head.postprocess(None, sc, cl, stv, return_masks=False, mask_features=feat,
                                                                        mask_kernels=kern)),
                 ("tail_cublas_fp32_bmm_then_postprocess", tail_library)):
    med, best = timed(fn)
    print(json.dumps({"variant": name, "batch": B, "n": n, "hw": [H, W], "channels": E, "ms": med, "min_ms": best,
                      "images_per_s": B / med * 1e3}))
# the last stage (solo_v2.py:599-627): kept masks -> 800x1333 image masks + boxes.  Two inputs: object-like masks
# (ellipses, utils/synthetic.solo_masks: what a trained head emits) and the worst case (the noise masks the random
# features above produce: every 16-pixel run has to be sampled)
import numpy as np  # noqa: E402
from detectron2_tensorflow_b200.utils import synthetic as syn  # noqa: E402
IH, IW = 800, 1333
obj = np.stack([syn.solo_masks(100, hw=(H, W), seed=70 + i)[0] for i in range(B)])  # [B, 100, H, W] fp32 0/1
flat = obj.reshape(B, 100, -1).astype(np.uint8)
flat = np.concatenate([flat, np.zeros((B, 100, (-flat.shape[-1]) % 64), np.uint8)], -1)
kept_obj = torch.from_numpy(np.packbits(flat, axis=-1, bitorder="little").view(np.int64)).to(dev)
kept_noise = head.postprocess(None, sc, cl, stv, return_masks=False, mask_features=feat, mask_kernels=kern)["packed_masks"]
for inp, kept in (("object_masks", kept_obj), ("noise_masks_worst_case", kept_noise)):
    for name, kw in (("upsample_uint8_masks_and_boxes", dict(return_masks=True, return_packed=False)),
                     ("upsample_packed_masks_and_boxes", dict(return_masks=False, return_packed=True))):
        med, best = timed(lambda: solo_upsample_masks(kept, (H, W), (IH, IW), 0.5, False, **kw))
        out_bytes = B * 100 * IH * IW * (1 if kw["return_masks"] else 1 / 8)
        print(json.dumps({"variant": name, "input": inp, "batch": B, "dets": 100, "from": [H, W], "to": [IH, IW], "ms": med,
                          "min_ms": best, "out_GB": out_bytes / 1e9, "write_GBps": out_bytes / med / 1e6,
                          "coverage": float(obj.mean()) if inp == "object_masks" else None,
                          "reference_fp32_bytes_per_pass_GB": B * 100 * IH * IW * 4 / 1e9}))
a = solo_dynamic_masks(feat, kern)
b = library_route(False)
mism = int((a[0] != b[0]).sum())
print(json.dumps({"check": "packed words differing from the cuBLAS-fp32 route", "words": mism, "of": a[0].numel(),
                  "sum_masks_max_abs_diff": float((a[1] - b[1]).abs().max())}))
