"""Pure-write (fill) and copy bandwidth of the GPU on a 1.7 GB buffer: the ceiling of the write-bound kernels."""
import json, os, sys
import torch, numpy as np
dev=torch.device('cuda',0)
n=1600*800*1333
x=torch.empty(n, dtype=torch.uint8, device=dev)
def med(fn,k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(k):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
r={}
r['zero_u8_ms']=med(lambda: x.zero_())
xi=x[:n//8*8].view(torch.int64)
r['fill_i64_ms']=med(lambda: xi.fill_(0))
y=torch.empty_like(x)
r['copy_ms']=med(lambda: y.copy_(x))
r['GBps_fill_i64']=n/r['fill_i64_ms']/1e6; r['GBps_copy_rw']=2*n/r['copy_ms']/1e6
print(json.dumps(r))
