"""Times RetinaNet post-processing at BASELINE config 3 (N=32) -- one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from detectron2_tensorflow_b200.modeling import RetinaNetInference
from detectron2_tensorflow_b200.utils import synthetic as syn
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
N, K = 32, 80
anchors = [torch.from_numpy(a).to(dev) for a in syn.retinanet_anchors()]
cls = [torch.randn((N, a.shape[0], K), device=dev, generator=g) * 1.5 - 4.6 for a in anchors]
dl = [torch.randn((N, a.shape[0], 4), device=dev, generator=g) * 0.3 for a in anchors]
head = RetinaNetInference(num_classes=K)
for _ in range(3):
    head.inference(cls, dl, anchors)
torch.cuda.synchronize()
ts = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); head.inference(cls, dl, anchors); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(json.dumps({"config": 3, "ms": float(np.median(ts)), "min_ms": min(ts)}))
