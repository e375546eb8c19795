import json, sys
sys.path.insert(0, '/root/repo')
import torch, bench
dev = torch.device('cuda', 0)
r = bench.config4_sweep(dev, 1, 0, cpu=False)
print(json.dumps(r['nms_sweep_thr0.7']))
