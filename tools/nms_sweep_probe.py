"""bench.py's configs[4] NMS sweep (one uncapped segment / 16 capped segments of 256..65,536 boxes), GPU only."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device('cuda', 0)
r = bench.config4_sweep(dev, 1, 0, cpu=False)
print(json.dumps(r['nms_sweep_thr0.7']))
