"""Generates the committed golden fixtures under tests/golden/ (run in the BUILD container only).

    python tests/golden/make_golden.py

Sources of truth, in decreasing authority:
  1. nms_reference_numpy.npz -- outputs of the REFERENCE's own TF-free numpy NMS
     (/root/reference/lib/structures/np_box_list_ops.py:146-216 with np_box_ops.py:48-64), imported
     by path under a stub package because lib/structures/__init__.py imports TensorFlow.  That code
     computes IoU partly in float64, so inputs are rejection-sampled until no pair's IoU lies within
     1e-4 of the threshold: then the fp32 TF rule and the fp64 numpy rule must select the same boxes.
  2. roi_align_torchvision.npz / topk_torch.npz / nms_torchvision.npz -- independent
     implementations available offline (torchvision.ops.roi_align(aligned=True), torch.topk,
     torchvision.ops.nms).  ROIAlign agrees to ~4e-5 abs (the reference's normalise/denormalise
     coordinate round trip differs from the textbook formula in fp32 rounding only).
  3. oracle_pins.npz -- regression pins of the oracle itself on adversarial inputs (ties, NaN, zero-area
     boxes, class-offset near-threshold pairs).  NOT independent evidence; they freeze today's behaviour.

/root/reference does not exist on the GPU box, so nothing at test time reads it: only this script does.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/lib/structures"


def load_reference_numpy_nms():
    pkg = types.ModuleType("refstructs")
    pkg.__path__ = [REF]
    sys.modules["refstructs"] = pkg
    mods = {}
    for name in ("np_box_ops", "np_box_list", "np_box_list_ops"):
        spec = importlib.util.spec_from_file_location(f"refstructs.{name}", os.path.join(REF, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refstructs.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def iou64(b):
    b = b.astype(np.float64)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    ih = np.maximum(0, np.minimum(b[:, None, 2], b[None, :, 2]) - np.maximum(b[:, None, 0], b[None, :, 0]))
    iw = np.maximum(0, np.minimum(b[:, None, 3], b[None, :, 3]) - np.maximum(b[:, None, 1], b[None, :, 1]))
    inter = ih * iw
    return inter / (area[:, None] + area[None, :] - inter)


def boxes_for_nms(rng, n, thr, clustered):
    while True:
        if clustered:
            k = max(n // 12, 2)
            cy, cx = rng.uniform(50, 750, k), rng.uniform(50, 1280, k)
            h, w = rng.uniform(30, 300, k), rng.uniform(30, 300, k)
            c = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1)
            b = c[rng.integers(0, k, n)] + rng.normal(0, 8, (n, 4))
        else:
            cy, cx = rng.uniform(0, 800, n), rng.uniform(0, 1333, n)
            h, w = rng.uniform(10, 400, n), rng.uniform(10, 400, n)
            b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1)
        b = b.astype(np.float32)
        if np.any(b[:, 2] <= b[:, 0]) or np.any(b[:, 3] <= b[:, 1]):
            continue
        io = iou64(b)
        if np.abs(io - thr).min() > 1e-4:
            return b


def make_nms_reference():
    mods = load_reference_numpy_nms()
    BoxList = mods["np_box_list"].BoxList
    ops = mods["np_box_list_ops"]
    rng = np.random.default_rng(100)
    out = {}
    cases = [(200, 0.7, 200, False), (300, 0.5, 300, True), (400, 0.7, 37, True), (150, 0.3, 150, True)]
    for ci, (n, thr, max_out, clustered) in enumerate(cases):
        b = boxes_for_nms(rng, n, thr, clustered)
        s = rng.permutation(n).astype(np.float32) / n  # unique scores: tie order is not exercised here
        bl = BoxList(b.copy())
        bl.add_field("scores", s.copy())
        res = ops.non_max_suppression(bl, max_output_size=max_out, iou_threshold=thr, score_threshold=-10.0)
        kept_scores = res.get_field("scores")
        order = {float(v): i for i, v in enumerate(s)}
        keep = np.array([order[float(v)] for v in kept_scores], np.int32)
        out[f"c{ci}_boxes"], out[f"c{ci}_scores"], out[f"c{ci}_keep"] = b, s, keep
        out[f"c{ci}_thr"], out[f"c{ci}_max_out"] = np.float32(thr), np.int32(max_out)
    out["num_cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(HERE, "nms_reference_numpy.npz"), **out)
    print("nms_reference_numpy.npz:", [len(out[f"c{i}_keep"]) for i in range(len(cases))])


def make_independent():
    import torch
    import torchvision
    rng = np.random.default_rng(200)
    # ROIAlign (single level) vs torchvision aligned=True
    N, H, W, C = 2, 24, 36, 8
    img = rng.standard_normal((N, H, W, C)).astype(np.float32)
    M = 48
    cy, cx = rng.uniform(0, H * 16, M), rng.uniform(0, W * 16, M)
    s = np.exp(rng.uniform(np.log(16), np.log(400), M))
    a = np.exp(rng.uniform(np.log(.5), np.log(2), M))
    hh, ww = s * np.sqrt(a), s / np.sqrt(a)
    boxes = np.stack([cy - hh / 2, cx - ww / 2, cy + hh / 2, cx + ww / 2], 1).astype(np.float32)
    boxes[:6] += 30  # partly outside
    bi = rng.integers(0, N, M).astype(np.int32)
    out = {"img": img, "boxes": boxes, "box_ind": bi, "scale": np.float32(1 / 16.)}
    tb = torch.from_numpy(np.concatenate([bi[:, None].astype(np.float32), boxes[:, [1, 0, 3, 2]]], 1))
    for (o, sr) in ((7, 0), (7, 2), (14, 0)):
        tv = torchvision.ops.roi_align(torch.from_numpy(img).permute(0, 3, 1, 2), tb, (o, o), 1 / 16., max(sr, 1), True)
        out[f"out_{o}_{sr}"] = tv.permute(0, 2, 3, 1).contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "roi_align_torchvision.npz"), **out)
    # top-k vs torch.topk (tie-free)
    x = rng.permutation(20000).astype(np.float32) * 1e-3 - 7.0
    v, i = torch.topk(torch.from_numpy(x), 1000)
    np.savez_compressed(os.path.join(HERE, "topk_torch.npz"), x=x, values=v.numpy(), indices=i.numpy().astype(np.int32))
    # NMS vs torchvision (xyxy there, yxyx here), tie-free
    b = boxes_for_nms(rng, 500, 0.7, True)
    s = rng.permutation(500).astype(np.float32)
    keep = torchvision.ops.nms(torch.from_numpy(b[:, [1, 0, 3, 2]].copy()), torch.from_numpy(s), 0.7).numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "nms_torchvision.npz"), boxes=b, scores=s, keep=keep, thr=np.float32(0.7))
    print("independent fixtures written")


def make_pins():
    import oracle
    from detectron2_tensorflow_b200.utils import synthetic as syn
    rng = np.random.default_rng(300)
    out = {}
    # math pins
    xs = np.concatenate([rng.uniform(-90, 90, 64), [0.0, -0.0, 1.0, -1.0, 88.5, -88.5, 4.135166556742356]]).astype(np.float32)
    out["exp_x"], out["exp_y"] = xs, oracle.expf(xs)
    ls = np.concatenate([np.exp(rng.uniform(-30, 30, 64)), [0.5, 1.0, 2.0, 2.220446e-16, 1e-38]]).astype(np.float32)
    out["log_x"], out["log_y"] = ls, oracle.logf(ls)
    # level assignment at the bin edges
    eb = np.array([[0, 0, 112, 112], [0, 0, 224, 224], [0, 0, 448, 448], [0, 0, 0, 0], [5, 5, 3, 9],
                   [0, 0, 111.99999, 112], [0, 0, 224.00002, 224], [0, 0, 56, 224], [0, 0, 1e4, 1e4]], np.float32)
    out["lvl_boxes"], out["lvl"] = eb, oracle.assign_boxes_to_levels(eb, 2, 5, 224, 4)
    # NMS with ties / degenerate boxes
    n = 96
    b = np.repeat(np.array([[10, 10, 60, 60], [12, 12, 58, 64], [200, 200, 260, 240], [0, 0, 0, 50]], np.float32), n // 4, 0)
    b = b + rng.integers(0, 3, (n, 4)).astype(np.float32)
    s = np.round(rng.standard_normal(n) * 2).astype(np.float32)
    s[::17] = -np.inf
    out["tie_boxes"], out["tie_scores"] = b, s
    out["tie_keep"] = oracle.nms(b, s, 20, 0.5)
    # top-k ties + NaN
    x = np.round(rng.standard_normal(500)).astype(np.float32)
    x[::41] = np.nan
    out["tk_x"] = x
    out["tk_v"], out["tk_i"] = oracle.top_k(x, 64)
    # Fast R-CNN class-offset composition
    N, R, K = 2, 40, 6
    idx = np.stack([np.repeat(np.arange(N), R), np.tile(np.arange(R), N)], 1).astype(np.int64)
    prop, _ = syn.rois(N, R, seed=3)
    deltas = (rng.standard_normal((N * R, K * 4)) * 0.3).astype(np.float32)
    boxes = oracle.apply_deltas(deltas, prop, (10., 10., 5., 5.))
    lg = rng.standard_normal((N * R, K + 1)) * 2
    e = np.exp(lg - lg.max(1, keepdims=True))
    sc = (e / e.sum(1, keepdims=True)).astype(np.float32)
    shapes = syn.image_shapes(N)
    fb, fs, fc, fv, fr, fn = oracle.fast_rcnn_inference(boxes, sc, idx, (N, R), shapes, 0.05, 0.5, 20, False)
    out.update(fr_boxes=boxes, fr_scores=sc, fr_idx=idx, fr_shapes=shapes, fr_ob=fb, fr_os=fs, fr_oc=fc, fr_ov=fv,
               fr_or=fr, fr_on=fn)
    # matrix-NMS
    m, c, s2 = syn.solo_masks(40, hw=(24, 40), num_classes=3, seed=9)
    out.update(mn_masks=np.packbits(m.astype(np.uint8), axis=None), mn_shape=np.array(m.shape), mn_classes=c,
               mn_scores=s2, mn_gauss=oracle.matrix_nms(m, c, s2, None, "gaussian", 2.0),
               mn_linear=oracle.matrix_nms(m, c, s2, None, "linear", 2.0))
    np.savez_compressed(os.path.join(HERE, "oracle_pins.npz"), **out)
    print("oracle_pins.npz written")


if __name__ == "__main__":
    make_nms_reference()
    make_independent()
    make_pins()
