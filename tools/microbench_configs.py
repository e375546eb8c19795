"""Device-side timing of the other BASELINE.json configs (they are parity-test cases, not bench lines):
  config 3  RetinaNet R50-FPN inference post-processing, batch 32
  config 4  SOLOv2 Matrix-NMS, 500 masks at 200x336, batch 16
  config 5  ROIAlign / NMS sweep, 256 .. 65536 ROIs / boxes
  2         SURVEY.md 8(d) stage numbers at config-2 sizes: NMS-only (80 segments x 2000 boxes) and the full proposal
            stage, gaussian / clustered / ties inputs, cold (L2 flushed) and warm, CPU oracle with 1 and all threads
  train     SURVEY.md 8(f) #3: fused RPN label assignment (16 x 268 K anchors), ROIPooler backward (16,000 ROIs)
Prints one JSON line per case (CUDA events, median of `--iters`, 256 MB L2 flush between iterations)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from detectron2_tensorflow_b200.layers import batch_nms, matrix_nms
from detectron2_tensorflow_b200.modeling import Box2BoxTransform, Matcher, ROIPooler, RetinaNetInference, label_boxes
from detectron2_tensorflow_b200.structures import BoxList, SparseBoxList, pairwise_iou
from detectron2_tensorflow_b200.utils import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--which", default="3,4,5")
args = ap.parse_args()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
HBM = 6525.2


def timeit(fn, iters=args.iters, warm=3, cold=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


which = set(args.which.split(","))
g = torch.Generator(device=dev).manual_seed(0)

if "2" in which:
    import time
    import oracle
    from detectron2_tensorflow_b200.modeling import find_top_rpn_proposals
    from detectron2_tensorflow_b200.structures import ImageList
    N = 16
    anchors = syn.rpn_anchors()
    shapes = syn.image_shapes(N)
    for variant in ("gaussian", "clustered", "ties"):
        logits, deltas = syn.rpn_inputs(N, seed={"gaussian": 2, "clustered": 3, "ties": 4}[variant], variant=variant,
                                        anchors=anchors)
        props = [oracle.rpn_predict_proposals(d, a) for d, a in zip(deltas, anchors)]
        tp = [torch.from_numpy(x).to(dev) for x in props]
        tl = [torch.from_numpy(x).to(dev) for x in logits]
        images = ImageList(None, torch.from_numpy(shapes).to(dev))
        run = lambda: find_top_rpn_proposals(tp, tl, images, 0.7, 2000, 1000, 0.0)
        cold, warm = timeit(run, iters=20, warm=5), timeit(run, iters=20, warm=5, cold=False)
        res = run()
        nms_in = N * sum(min(2000, a.shape[0]) for a in anchors)
        # NMS-only on exactly the boxes that enter NMS: per (image, level) top-2000, clipped (oracle prepares them)
        segs_b, segs_s = [], []
        for l in range(len(anchors)):
            for n in range(N):
                k = min(2000, logits[l].shape[1])
                v, i = oracle.top_k(logits[l][n], k)
                b = props[l][n][i].copy()
                b[:, 0::2] = np.clip(b[:, 0::2], 0, 800.0)
                b[:, 1::2] = np.clip(b[:, 1::2], 0, 1333.0)
                pad = 2000 - k
                segs_b.append(np.concatenate([b, np.zeros((pad, 4), np.float32)]))
                segs_s.append(np.concatenate([v, np.full(pad, -np.inf, np.float32)]))
        sb, ss = np.stack(segs_b), np.stack(segs_s)
        tb, ts_ = torch.from_numpy(sb).to(dev), torch.from_numpy(ss).to(dev)
        nrun = lambda: batch_nms(tb, ts_, 1000, axis=1, iou_threshold=0.7)
        ncold, nwarm = timeit(nrun, iters=20, warm=5), timeit(nrun, iters=20, warm=5, cold=False)
        kept = int(nrun()[1].sum())
        cpu = {}
        for thr in (1, oracle.max_threads()):
            oracle.set_num_threads(thr)
            t0 = time.perf_counter()
            oracle.batch_nms(sb[:16], ss[:16], 1000, 0.7)  # 16 of the 80 segments (bounded sample)
            t_nms = (time.perf_counter() - t0) * 5
            t0 = time.perf_counter()
            oracle.find_top_rpn_proposals([x[:2] for x in props], [x[:2] for x in logits], shapes[:2], 0.7, 2000, 1000, 0.0)
            t_stage = (time.perf_counter() - t0) * 8  # 2 of the 16 images
            cpu[f"threads_{thr}"] = {"nms_only_boxes_per_s": nms_in / t_nms, "proposal_stage_boxes_per_s": nms_in / t_stage}
        oracle.set_num_threads(0)
        print(json.dumps({"config": 2, "case": f"RPN proposal stage, {variant} inputs, N=16 (80 segments, {nms_in} boxes enter NMS)",
                          "proposal_stage_ms_cold": cold, "proposal_stage_ms_warm": warm,
                          "proposal_stage_boxes_per_s": nms_in / cold * 1e3,
                          "nms_only_ms_cold": ncold, "nms_only_ms_warm": nwarm, "nms_only_boxes_per_s": nms_in / ncold * 1e3,
                          "kept_by_nms": kept, "valid_proposals": int(res.get_field("is_valid").sum()),
                          "cpu_oracle": cpu,
                          "speedup_nms_only_vs_all_threads": (nms_in / ncold * 1e3) /
                          cpu[f"threads_{oracle.max_threads()}"]["nms_only_boxes_per_s"]}))

if "3" in which:
    N, K = 32, 80
    anchors = [torch.from_numpy(a).to(dev) for a in syn.retinanet_anchors()]
    cls = [torch.randn((N, a.shape[0], K), device=dev, generator=g) * 1.5 - 4.6 for a in anchors]
    dl = [torch.randn((N, a.shape[0], 4), device=dev, generator=g) * 0.3 for a in anchors]
    head = RetinaNetInference(num_classes=K)
    ms = timeit(lambda: head.inference(cls, dl, anchors))
    nscores = sum(c.numel() for c in cls)
    print(json.dumps({"config": 3, "case": "RetinaNet post-processing N=32", "ms": ms, "images_per_s": N / ms * 1e3,
                      "class_scores_per_s": nscores / ms * 1e3, "score_bytes_GB": nscores * 4 / 1e9,
                      "single_read_GBps": nscores * 4 / ms / 1e6, "frac_hbm_single_read": nscores * 4 / ms / 1e6 / HBM}))

if "4" in which:
    B, n, H, W = 16, 500, 200, 336
    m, c, s = syn.solo_masks(n, hw=(H, W), seed=7)
    masks = torch.from_numpy(m).to(dev)[None].repeat(B, 1, 1, 1).contiguous()
    classes = torch.from_numpy(c).to(dev)[None].repeat(B, 1).contiguous()
    scores = torch.from_numpy(s).to(dev)[None].repeat(B, 1).contiguous()
    ms = timeit(lambda: matrix_nms(masks, classes, scores))
    by = masks.numel() * 4
    print(json.dumps({"config": 4, "case": "SOLOv2 matrix-NMS B=16 n=500 200x336", "ms": ms, "images_per_s": B / ms * 1e3,
                      "masks_per_s": B * n / ms * 1e3, "mask_read_GB": by / 1e9, "GBps": by / ms / 1e6,
                      "frac_hbm": by / ms / 1e6 / HBM}))

    # 8(f) #4: the mask stage feeding Matrix-NMS with bit-packed masks (logits -> words; no fp32 0/1 masks)
    from detectron2_tensorflow_b200.modeling import solo_mask_encode
    logits = (masks * 8.0 - 4.0) + torch.randn(masks.shape, device=dev, generator=g) * 0.5
    ms_e = timeit(lambda: solo_mask_encode(logits, 0.5))
    packed, sums, _ = solo_mask_encode(logits, 0.5)
    ms_p = timeit(lambda: matrix_nms(None, classes, scores, sum_masks=sums, packed_masks=packed, mask_hw=H * W))
    print(json.dumps({"config": 4, "case": "SOLOv2 mask stage: sigmoid+threshold+bit-pack+sums of the logits", "ms": ms_e,
                      "GBps": by / ms_e / 1e6, "frac_hbm": by / ms_e / 1e6 / HBM}))
    from detectron2_tensorflow_b200.modeling import SOLOv2Inference
    head = SOLOv2Inference(0.5, 500, "gaussian", 2.0, 0.05, 100)
    strides = torch.full((B, n), 8.0, device=dev)
    ms_t = timeit(lambda: head.postprocess(logits, scores, classes, strides, return_masks=False))
    ms_tm = timeit(lambda: head.postprocess(logits, scores, classes, strides, return_masks=True))
    print(json.dumps({"config": 4, "case": "SOLOv2 inference tail after the conv (encode + filter + scoring + top-k + Matrix-NMS + emit), B=16 n=500",
                      "ms_packed_masks_out": ms_t, "ms_fp32_masks_out": ms_tm, "images_per_s": B / ms_t * 1e3,
                      "logit_GBps": by / ms_t / 1e6, "frac_hbm": by / ms_t / 1e6 / HBM}))
    print(json.dumps({"config": 4, "case": "Matrix-NMS on packed masks (no fp32 mask read)", "ms": ms_p,
                      "images_per_s": B / ms_p * 1e3, "packed_GB": packed.numel() * 8 / 1e9}))

if "5" in which:
    N, C = 16, 256
    feats = [torch.randn((N,) + syn.level_hw(s) + (C,), device=dev, generator=g) for s in syn.FPN_STRIDES]
    fbytes = sum(f.numel() for f in feats) * 4
    pooler = ROIPooler(7, [1 / 4., 1 / 8., 1 / 16., 1 / 32.], 0, "ROIAlignV2")
    for M in (256, 1024, 4096, 16384, 65536):
        boxes, idx = syn.rois(N, M // N, seed=1)
        inst = SparseBoxList(torch.from_numpy(idx).to(dev), BoxList(torch.from_numpy(boxes).to(dev)), (N, M // N))
        ms = timeit(lambda: pooler(feats, inst))
        alg = M * 49 * C * 4 + min(fbytes, M * 49 * 4 * C * 4) + M * 24
        print(json.dumps({"config": 5, "case": f"ROIAlign 7x7 sweep M={M}", "ms": ms, "rois_per_s": M / ms * 1e3,
                          "alg_GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / HBM}))
    # sampling_ratio = 2 (the keypoint head's setting, defaults.py:513): 4 samples per bin, same algorithmic bytes
    for (osz, M) in ((7, 16384), (14, 1600)):
        pooler2 = ROIPooler(osz, [1 / 4., 1 / 8., 1 / 16., 1 / 32.], 2, "ROIAlignV2")
        boxes, idx = syn.rois(N, M // N, seed=1)
        inst = SparseBoxList(torch.from_numpy(idx).to(dev), BoxList(torch.from_numpy(boxes).to(dev)), (N, M // N))
        ms = timeit(lambda: pooler2(feats, inst))
        alg = M * osz * osz * C * 4 + min(fbytes, M * osz * osz * 16 * C * 4) + M * 24
        print(json.dumps({"config": 5, "case": f"ROIAlign {osz}x{osz} sampling_ratio=2 M={M}", "ms": ms,
                          "rois_per_s": M / ms * 1e3, "alg_GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / HBM}))
    rng = np.random.default_rng(5)
    for n in (256, 1024, 4096, 16384, 65536):
        cy, cx = rng.uniform(0, 800, n), rng.uniform(0, 1333, n)
        h, w = rng.uniform(16, 300, n), rng.uniform(16, 300, n)
        b = torch.from_numpy(np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1).astype(np.float32)).to(dev)[None]
        sc = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(dev)[None]
        out = {}
        def run():
            out["k"], out["n"] = batch_nms(b, sc, n, axis=1, iou_threshold=0.7)
        ms = timeit(run, iters=5)
        print(json.dumps({"config": 5, "case": f"NMS sweep n={n} (one segment, thr 0.7, uncapped)", "ms": ms,
                          "boxes_per_s": n / ms * 1e3, "kept": int(out["n"][0])}))

if "train" in which:
    # ---- RPN ground truth: pairwise_iou + Matcher + get_deltas fused (rpn_outputs.py:245-304)
    N = 16
    anchors = torch.from_numpy(np.concatenate(syn.rpn_anchors(), 0)).to(dev)
    P = anchors.shape[0]
    rng = np.random.default_rng(11)
    for G in (20, 100):
        cy, cx = rng.uniform(0, 800, (N, G)), rng.uniform(0, 1333, (N, G))
        h, w = rng.uniform(16, 500, (N, G)), rng.uniform(16, 500, (N, G))
        gt = torch.from_numpy(np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 2).astype(np.float32)).to(dev)
        valid = torch.ones((N, G), dtype=torch.bool, device=dev)
        crowd = torch.zeros((N, G), dtype=torch.bool, device=dev)
        m = Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
        bt = Box2BoxTransform((1., 1., 1., 1.))
        ms = timeit(lambda: label_boxes(anchors, gt, valid, m, gt_crowd=crowd, box2box_transform=bt))
        out_bytes = N * P * 32 + P * 16
        print(json.dumps({"config": "train", "case": f"RPN label assignment N={N} anchors={P} G={G} (fused iou+matcher+deltas)",
                          "ms": ms, "anchors_per_s": N * P / ms * 1e3, "pairs_per_s": 2 * N * P * G / ms * 1e3,
                          "alg_GBps": out_bytes / ms / 1e6, "frac_hbm": out_bytes / ms / 1e6 / HBM,
                          "unfused_matrix_bytes_GB": N * P * G * 4 * 2 / 1e9}))
    q_ms = timeit(lambda: pairwise_iou(gt[0], anchors))
    print(json.dumps({"config": "train", "case": f"pairwise_iou [{G}, {P}] materialised", "ms": q_ms,
                      "GBps": G * P * 4 / q_ms / 1e6, "frac_hbm": G * P * 4 / q_ms / 1e6 / HBM}))
    # ---- ROIPooler backward (box head, 16,000 ROIs 7x7x256)
    C = 256
    shapes = [(N,) + syn.level_hw(s) + (C,) for s in syn.FPN_STRIDES]
    grads = [torch.zeros(sh, device=dev) for sh in shapes]
    pooler = ROIPooler(7, [1 / 4., 1 / 8., 1 / 16., 1 / 32.], 0, "ROIAlignV2")
    for M in (16000, 8192):
        boxes, idx = syn.rois(N, M // N, seed=1)
        inst = SparseBoxList(torch.from_numpy(idx).to(dev), BoxList(torch.from_numpy(boxes).to(dev)), (N, M // N))
        go = torch.randn((M, 7, 7, C), device=dev, generator=g)
        ms = timeit(lambda: pooler.backward(go, shapes, inst, grad_x=grads))
        rd = M * 49 * C * 4
        print(json.dumps({"config": "train", "case": f"ROIPooler backward M={M} 7x7x{C} (accumulate into resident grads)",
                          "ms": ms, "rois_per_s": M / ms * 1e3, "grad_read_GBps": rd / ms / 1e6,
                          "red_GBps": 4 * rd / ms / 1e6, "frac_hbm_on_grad_read": rd / ms / 1e6 / HBM}))
    zero_ms = timeit(lambda: [t_.zero_() for t_ in grads])
    print(json.dumps({"config": "train", "case": "zero the 1.46 GB of feature gradients (torch memset)", "ms": zero_ms}))
