"""Probe: does software-pipelining the device-resident step over image chunks on several streams (the RPN stage of
chunk i+1 overlapping the HBM-bound ROIAlign of chunk i) beat the single-stream step?  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench

dev = torch.device("cuda", 0)
host = bench.make_host_inputs(16)
x = bench.to_torch(host, dev=dev)
eng = bench.make_engine()
R = bench.ROIS_PER_IMAGE


def cut(b, e):
    d = {"anchors": x["anchors"], "shapes": x["shapes"][b:e], "scores": x["scores"][b * R:e * R],
         "cls_deltas": x["cls_deltas"][b * R:e * R]}
    for k in ("logits", "deltas", "feats"):
        d[k] = [t[b:e] for t in x[k]]
    return d


def run_chunks(nchunks, streams, stagger):
    cur = torch.cuda.current_stream(dev)
    start = torch.cuda.Event()
    start.record(cur)
    per = 16 // nchunks
    prev_rpn = None
    outs = []
    for c in range(nchunks):
        s = streams[c % len(streams)]
        s.wait_event(start)
        if stagger and prev_rpn is not None:
            s.wait_event(prev_rpn)
        with torch.cuda.stream(s):
            evs = [torch.cuda.Event() for _ in range(5)]
            outs.append(eng(cut(c * per, (c + 1) * per), evs))
            prev_rpn = evs[1]
    for s in streams:
        cur.wait_stream(s)
    return outs


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {"single_stream_16": timeit(lambda: eng(x))}


def graphed(fn):
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


try:
    g1 = graphed(lambda: eng(x))
    res["graph_single_stream_16"] = timeit(g1.replay)
    for nch, ns, stg in ((4, 4, False),):
        streams = [torch.cuda.Stream(dev) for _ in range(ns)]
        gg = graphed(lambda: run_chunks(nch, streams, stg))
        res[f"graph_chunks{nch}_streams{ns}_{'staggered' if stg else 'free'}"] = timeit(gg.replay)
except Exception as ex:  # noqa
    res["graph_error"] = repr(ex)[:300]

ref = bench.MaskRCNNPostBackbone.flatten_outputs(eng(x)) if hasattr(bench, "MaskRCNNPostBackbone") else None
for nch, ns, stg in ((2, 2, False), (2, 2, True)):
    streams = [torch.cuda.Stream(dev) for _ in range(ns)]
    res[f"chunks{nch}_streams{ns}_{'staggered' if stg else 'free'}"] = timeit(lambda: run_chunks(nch, streams, stg))
# two (or three) graph instances replayed on alternating streams: consecutive steps overlap
for depth in (2, 3):
    gs = [eng.capture(x, chunks=4) for _ in range(depth)]
    ss = [torch.cuda.Stream(dev) for _ in range(depth)]

    def run_k(K=60):
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        for s_ in ss:
            s_.wait_event(ev)
        for k in range(K):
            with torch.cuda.stream(ss[k % depth]):
                gs[k % depth].replay()
        for s_ in ss:
            cur.wait_stream(s_)
    run_k(6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_k(60)
    e1.record()
    torch.cuda.synchronize()
    res[f"steps_in_flight_{depth}_ms_per_step"] = e0.elapsed_time(e1) / 60
print(json.dumps(res))
