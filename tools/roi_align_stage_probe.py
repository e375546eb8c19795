"""Stage times of the box / mask poolers inside the 16-image step (CUDA events); used with D2B_LIB for same-box A/B of
build variants."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
dev = torch.device('cuda', 0)
hp = bench.make_engine()
x = bench.to_torch(bench.make_host_inputs(16), dev=dev)
for _ in range(3): hp(x)
torch.cuda.synchronize()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(30)]
for i in range(30): hp(x, events=ev[i])
torch.cuda.synchronize()
st = [float(np.median([ev[i][s].elapsed_time(ev[i][s + 1]) for i in range(30)])) for s in range(4)]
print(json.dumps({"box_roi_align": st[1], "mask_roi_align": st[3]}))
