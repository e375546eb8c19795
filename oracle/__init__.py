"""CPU oracle for the post-backbone hot path -- TEST INFRASTRUCTURE ONLY.

numpy-in / numpy-out ctypes bindings over ``oracle/liboracle.so`` (built from
``oracle/d2b_oracle.c`` by ``oracle/Makefile``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package; the product package never does.

Parity status: the COMPOSITION is pinned on the reference's own Python, executed unmodified on a numpy
stand-in for the TF ops it calls (``tests/golden/make_reference_golden.py`` -> ``reference_python.npz``); the
arithmetic inside the stock TF kernels stays *unpinned* (TensorFlow cannot be installed here) and is
cross-checked against the reference's numpy NMS and torchvision/torch -- see ``tests/golden/make_golden.py``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

SCALE_CLAMP = float(np.log(1000.0 / 16))  # lib/modeling/box_regression.py:10


def build(force=False):
    src = os.path.join(_HERE, "d2b_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_expf.restype = C.c_float
        _lib.orc_expf.argtypes = [C.c_float]
        _lib.orc_logf.restype = C.c_float
        _lib.orc_logf.argtypes = [C.c_float]
        _lib.orc_sigmoidf.restype = C.c_float
        _lib.orc_sigmoidf.argtypes = [C.c_float]
        _lib.orc_nms.restype = C.c_int32
        _lib.orc_get_max_threads.restype = C.c_int
    return _lib


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))


def max_threads():
    return int(lib().orc_get_max_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def expf(x):
    x = _f32(x)
    return np.array([lib().orc_expf(float(v)) for v in x.ravel()], np.float32).reshape(x.shape)


def logf(x):
    x = _f32(x)
    return np.array([lib().orc_logf(float(v)) for v in x.ravel()], np.float32).reshape(x.shape)


def sigmoidf(x):
    x = _f32(x)
    return np.array([lib().orc_sigmoidf(float(v)) for v in x.ravel()], np.float32).reshape(x.shape)


def assign_boxes_to_levels(boxes, min_level, max_level, canonical_box_size=224, canonical_level=4):
    boxes = _f32(boxes).reshape(-1, 4)
    out = np.empty(boxes.shape[0], np.int64)
    lib().orc_assign_boxes_to_levels(_p(boxes), C.c_int64(boxes.shape[0]), int(min_level), int(max_level),
                                     int(canonical_box_size), int(canonical_level), _p(out))
    return out


def crop_and_resize(image, boxes, box_ind, crop_size, aligned=True, pad_border=True):
    image = _f32(image)
    N, H, W, Cc = image.shape
    boxes = _f32(boxes).reshape(-1, 4)
    box_ind = np.ascontiguousarray(box_ind, np.int32)
    M = boxes.shape[0]
    ch, cw = int(crop_size[0]), int(crop_size[1])
    out = np.zeros((M, ch, cw, Cc), np.float32)
    rc = lib().orc_crop_and_resize(_p(image), N, H, W, Cc, _p(boxes), _p(box_ind), C.c_int64(M), ch, cw,
                                   int(bool(aligned)), int(bool(pad_border)), _p(out))
    assert rc == 0
    return out


def roi_align(image, boxes, box_ind, output_size, spatial_scale, sampling_ratio, aligned=True):
    image = _f32(image)
    N, H, W, Cc = image.shape
    boxes = _f32(boxes).reshape(-1, 4)
    box_ind = np.ascontiguousarray(box_ind, np.int32)
    M = boxes.shape[0]
    oh, ow = int(output_size[0]), int(output_size[1])
    out = np.zeros((M, oh, ow, Cc), np.float32)
    rc = lib().orc_roi_align(_p(image), N, H, W, Cc, _p(boxes), _p(box_ind), C.c_int64(M), oh, ow,
                             C.c_float(spatial_scale), int(sampling_ratio), int(bool(aligned)), _p(out))
    assert rc == 0
    return out


def roi_pooler(feats, scales, boxes, batch_idx, output_size, sampling_ratio, aligned=True,
               canonical_box_size=224, canonical_level=4):
    feats = [_f32(f) for f in feats]
    L = len(feats)
    N, _, _, Cc = feats[0].shape
    Hs = np.array([f.shape[1] for f in feats], np.int32)
    Ws = np.array([f.shape[2] for f in feats], np.int32)
    sc = _f32(scales)
    boxes = _f32(boxes).reshape(-1, 4)
    batch_idx = np.ascontiguousarray(batch_idx, np.int64)
    M = boxes.shape[0]
    oh, ow = int(output_size[0]), int(output_size[1])
    out = np.zeros((M, oh, ow, Cc), np.float32)
    counts = np.zeros(L, np.int32)
    rc = lib().orc_roi_pooler(_ptr_array(feats), _p(Hs), _p(Ws), L, N, Cc, _p(sc), _p(boxes), _p(batch_idx),
                              C.c_int64(M), oh, ow, int(sampling_ratio), int(bool(aligned)),
                              int(canonical_box_size), int(canonical_level), _p(out), _p(counts))
    assert rc == 0
    return out, counts


def apply_deltas(deltas, boxes, weights, scale_clamp=SCALE_CLAMP):
    boxes = _f32(boxes).reshape(-1, 4)
    n = boxes.shape[0]
    deltas = _f32(deltas).reshape(n, -1)
    k = deltas.shape[1] // 4
    w = _f32(weights)
    out = np.empty_like(deltas)
    lib().orc_apply_deltas(_p(deltas), _p(boxes), C.c_int64(n), k, _p(w), C.c_float(scale_clamp), _p(out))
    return out


def top_k(x, k):
    x = _f32(x).ravel()
    k = min(int(k), x.size)
    vals = np.empty(k, np.float32)
    idx = np.empty(k, np.int32)
    lib().orc_top_k(_p(x), C.c_int64(x.size), C.c_int64(k), _p(vals), _p(idx))
    return vals, idx


def box_iou(a, b):
    a = _f32(a)
    b = _f32(b)
    f = lib().orc_box_iou
    f.restype = C.c_float
    return float(f(_p(a), _p(b)))


def nms(boxes, scores, max_output_size, iou_threshold):
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores).ravel()
    keep = np.empty(max(int(max_output_size), 1), np.int32)
    c = lib().orc_nms(_p(boxes), _p(scores), C.c_int64(boxes.shape[0]), C.c_int32(int(max_output_size)),
                      C.c_float(iou_threshold), _p(keep))
    return keep[:c].copy()


def batch_nms(boxes, scores, max_output_size, iou_threshold=0.5):
    boxes = _f32(boxes)
    scores = _f32(scores)
    B, n = scores.shape
    keep = np.empty((B, max_output_size), np.int32)
    num = np.empty(B, np.int32)
    lib().orc_batch_nms(_p(boxes), _p(scores), B, C.c_int64(n), C.c_int32(max_output_size),
                        C.c_float(iou_threshold), _p(keep), _p(num))
    return keep, num


def rpn_predict_proposals(deltas, anchors, weights=(1.0, 1.0, 1.0, 1.0), scale_clamp=SCALE_CLAMP):
    """deltas [N, HWA, 4], anchors [HWA, 4] -> [N, HWA, 4] (rpn_outputs.py:403-426)."""
    deltas = _f32(deltas)
    anchors = _f32(anchors)
    N, hwa, _ = deltas.shape
    out = np.empty_like(deltas)
    w = _f32(weights)
    lib().orc_rpn_predict_proposals(_p(deltas), _p(anchors), N, C.c_int64(hwa), _p(w), C.c_float(scale_clamp),
                                    _p(out))
    return out


def find_top_rpn_proposals(proposals, logits, image_shapes, nms_thresh, pre_nms_topk, post_nms_topk,
                           min_box_side_len):
    proposals = [_f32(p) for p in proposals]
    logits = [_f32(x) for x in logits]
    L = len(proposals)
    N = logits[0].shape[0]
    hwa = np.array([x.shape[1] for x in logits], np.int64)
    shapes = np.ascontiguousarray(image_shapes, np.int32).reshape(N, 2)
    ob = np.zeros((N, post_nms_topk, 4), np.float32)
    ol = np.zeros((N, post_nms_topk), np.float32)
    ov = np.zeros((N, post_nms_topk), np.uint8)
    on = np.zeros(N, np.int32)
    lib().orc_find_top_rpn_proposals(_ptr_array(proposals), _ptr_array(logits), _p(hwa), L, N, _p(shapes),
                                     C.c_float(nms_thresh), int(pre_nms_topk), int(post_nms_topk),
                                     C.c_float(min_box_side_len), _p(ob), _p(ol), _p(ov), _p(on))
    return ob, ol, ov.astype(bool), on


def fast_rcnn_inference(boxes, scores, indices, dense_shape, image_shapes, score_thresh, nms_thresh,
                        topk_per_image, nms_cls_agnostic=False):
    scores = _f32(scores)
    M, K1 = scores.shape
    K = K1 - 1
    boxes = _f32(boxes).reshape(M, -1)
    Kb = boxes.shape[1] // 4
    indices = np.ascontiguousarray(indices, np.int64).reshape(M, 2)
    N, Rmax = int(dense_shape[0]), int(dense_shape[1])
    shapes = np.ascontiguousarray(image_shapes, np.int32).reshape(N, 2)
    T = int(topk_per_image)
    ob = np.zeros((N, T, 4), np.float32)
    os_ = np.zeros((N, T), np.float32)
    oc = np.zeros((N, T), np.int64)
    ov = np.zeros((N, T), np.uint8)
    orr = np.zeros((N, T), np.int32)
    on = np.zeros(N, np.int32)
    lib().orc_fast_rcnn_inference(_p(boxes), _p(scores), _p(indices), C.c_int64(M), N, Rmax, Kb, K, _p(shapes),
                                  C.c_float(score_thresh), C.c_float(nms_thresh), T, int(bool(nms_cls_agnostic)),
                                  _p(ob), _p(os_), _p(oc), _p(ov), _p(orr), _p(on))
    return ob, os_, oc, ov.astype(bool), orr, on


def retinanet_inference(box_cls, box_delta, anchors, num_classes, topk_candidates, score_thresh, nms_thresh,
                        max_det, weights=(1.0, 1.0, 1.0, 1.0), scale_clamp=SCALE_CLAMP):
    box_cls = [_f32(x) for x in box_cls]
    box_delta = [_f32(x) for x in box_delta]
    anchors = [_f32(x) for x in anchors]
    L = len(box_cls)
    N = box_cls[0].shape[0]
    hwa = np.array([a.shape[0] for a in anchors], np.int64)
    T = int(max_det)
    ob = np.zeros((N, T, 4), np.float32)
    os_ = np.zeros((N, T), np.float32)
    oc = np.zeros((N, T), np.int32)
    ov = np.zeros((N, T), np.uint8)
    on = np.zeros(N, np.int32)
    w = _f32(weights)
    lib().orc_retinanet_inference(_ptr_array(box_cls), _ptr_array(box_delta), _ptr_array(anchors), _p(hwa), L, N,
                                  int(num_classes), int(topk_candidates), C.c_float(score_thresh),
                                  C.c_float(nms_thresh), T, _p(w), C.c_float(scale_clamp), _p(ob), _p(os_), _p(oc),
                                  _p(ov), _p(on))
    return ob, os_, oc, ov.astype(bool), on


def matrix_nms(masks, classes, scores, sum_masks=None, kernel="gaussian", sigma=2.0):
    masks = _f32(masks)
    n = masks.shape[0]
    hw = int(np.prod(masks.shape[1:]))
    classes = np.ascontiguousarray(classes, np.int64)
    scores = _f32(scores)
    kid = {"gaussian": 0, "linear": 1}.get(kernel, -1)
    if kid < 0:
        raise NotImplementedError(f"NMS kernel {kernel} not implemented yet.")
    sm = None if sum_masks is None else _f32(sum_masks)
    out = np.zeros(n, np.float32)
    rc = lib().orc_matrix_nms(_p(masks), _p(classes), _p(scores), None if sm is None else _p(sm), n,
                              C.c_int64(hw), kid, C.c_float(sigma), _p(out))
    assert rc == 0
    return out


def reframe_box_masks_to_image_masks(box_masks, boxes, image_shape, mask_threshold=0.5):
    """lib/structures/mask_ops.py:7-56 -> uint8 [M, H, W]."""
    box_masks = _f32(box_masks)
    M, mh, mw = box_masks.shape
    boxes = _f32(boxes).reshape(M, 4)
    H, W = int(image_shape[0]), int(image_shape[1])
    out = np.zeros((M, H, W), np.uint8)
    lib().orc_reframe_box_masks(_p(box_masks), _p(boxes), C.c_int64(M), mh, mw, H, W, C.c_float(mask_threshold), _p(out))
    return out


def pairwise_iou(boxes1, boxes2, iou_type="iou"):
    """lib/structures/box_list_ops.py:295-371 -> [n1, n2]; iou_type in iou / giou / diou / ciou."""
    b1 = _f32(boxes1).reshape(-1, 4)
    b2 = _f32(boxes2).reshape(-1, 4)
    out = np.zeros((b1.shape[0], b2.shape[0]), np.float32)
    if iou_type == "iou":
        lib().orc_pairwise_iou(_p(b1), C.c_int64(b1.shape[0]), _p(b2), C.c_int64(b2.shape[0]), _p(out))
    else:
        lib().orc_pairwise_iou_variant(_p(b1), C.c_int64(b1.shape[0]), _p(b2), C.c_int64(b2.shape[0]),
                                       {"giou": 1, "diou": 2, "ciou": 3}[iou_type], _p(out))
    return out


def matcher(q, thresholds, labels, allow_low_quality_matches=False, crowd=None, difficult=None):
    """lib/modeling/matcher.py:8-174 -> (matches int64 [N], match_labels int64 [N])."""
    q = _f32(q)
    M, N = q.shape
    th = _f32(thresholds)
    lab = np.ascontiguousarray(labels, np.int32)
    cr = None if crowd is None else _f32(crowd)
    df = None if difficult is None else _f32(difficult)
    matches = np.zeros(N, np.int64)
    ml = np.zeros(N, np.int64)
    lib().orc_matcher(_p(q), C.c_int64(M), C.c_int64(N), None if cr is None else _p(cr),
                      C.c_int64(0 if cr is None else cr.shape[0]), None if df is None else _p(df),
                      C.c_int64(0 if df is None else df.shape[0]), _p(th), len(th), _p(lab),
                      int(bool(allow_low_quality_matches)), _p(matches), _p(ml))
    return matches, ml


def get_deltas(src_boxes, target_boxes, weights):
    """lib/modeling/box_regression.py:38-74."""
    s = _f32(src_boxes).reshape(-1, 4)
    t = _f32(target_boxes).reshape(-1, 4)
    out = np.zeros_like(s)
    lib().orc_get_deltas(_p(s), _p(t), C.c_int64(s.shape[0]), _p(_f32(weights)), _p(out))
    return out


def label_boxes(pred_boxes, gt_boxes, gt_valid, thresholds, labels, allow_low_quality_matches=False, gt_crowd=None,
                gt_difficult=None, pred_counts=None, boundary_threshold=-1.0, image_shapes=None, weights=None):
    """rpn_outputs.py:245-304 / roi_heads.py:100-165: pairwise_iou + Matcher (+ inside_window, get_deltas).

    pred_boxes [P,4] (shared, anchors) or [N,P,4]; gt_boxes [N,G,4]; flags [N,G].
    -> (matches [N,P] int64, labels [N,P] int64, deltas [N,P,4] or None)."""
    pb = _f32(pred_boxes)
    shared = pb.ndim == 2
    gt = _f32(gt_boxes)
    N, G = gt.shape[:2]
    P = pb.shape[-2]
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    v, c, d = u8(gt_valid), u8(gt_crowd), u8(gt_difficult)
    pc = None if pred_counts is None else np.ascontiguousarray(pred_counts, np.int32)
    sh = None if image_shapes is None else np.ascontiguousarray(image_shapes, np.int32)
    th = _f32(thresholds)
    lab = np.ascontiguousarray(labels, np.int32)
    w = None if weights is None else _f32(weights)
    matches = np.zeros((N, P), np.int64)
    out_labels = np.zeros((N, P), np.int64)
    deltas = None if w is None else np.zeros((N, P, 4), np.float32)
    q = lambda a: None if a is None else _p(a)
    lib().orc_label_boxes(_p(pb), int(shared), q(pc), N, P, _p(gt), _p(v), q(c), q(d), G, _p(th), len(th), _p(lab),
                          int(bool(allow_low_quality_matches)), C.c_float(boundary_threshold), q(sh), q(w),
                          _p(matches), _p(out_labels), q(deltas))
    return matches, out_labels, deltas


def roi_align_backward(grad_out, image_shape, boxes, box_ind, spatial_scale, sampling_ratio, aligned=True):
    """Gradient of lib/layers/roi_align.py:45-66 w.r.t. the NHWC feature map `image_shape` = (N,H,W,C)."""
    g = _f32(grad_out)
    M, oh, ow, Cc = g.shape
    N, H, W, C2 = image_shape
    assert C2 == Cc
    b = _f32(boxes).reshape(-1, 4)
    bi = np.ascontiguousarray(box_ind, np.int32)
    out = np.zeros(image_shape, np.float32)
    rc = lib().orc_roi_align_backward(_p(g), N, H, W, Cc, _p(b), _p(bi), C.c_int64(M), oh, ow,
                                      C.c_float(spatial_scale), int(sampling_ratio), int(aligned), _p(out))
    assert rc == 0
    return out


def roi_pooler_backward(grad_out, feat_shapes, scales, boxes, batch_idx, sampling_ratio, aligned=True,
                        canonical_box_size=224, canonical_level=4):
    """Gradient of ROIPooler.call (lib/modeling/poolers.py:134-180) w.r.t. every level's feature map."""
    g = _f32(grad_out)
    M, oh, ow, Cc = g.shape
    L = len(feat_shapes)
    N = feat_shapes[0][0]
    outs = [np.zeros(s, np.float32) for s in feat_shapes]
    Hs = (C.c_int * L)(*[s[1] for s in feat_shapes])
    Ws = (C.c_int * L)(*[s[2] for s in feat_shapes])
    b = _f32(boxes).reshape(-1, 4)
    bi = np.ascontiguousarray(batch_idx, np.int64)
    rc = lib().orc_roi_pooler_backward(_p(g), _ptr_array(outs), Hs, Ws, L, N, Cc, _p(_f32(scales)), _p(b), _p(bi),
                                       C.c_int64(M), oh, ow, int(sampling_ratio), int(aligned),
                                       int(canonical_box_size), int(canonical_level))
    assert rc == 0
    return outs


def yolo_inference(boxes, probs, score_thresh, nms_thresh, post_nms_topk):
    """lib/modeling/single_stage_heads/yolov4_outputs.py:331-390 -> (boxes, scores, classes int64, valid, num)."""
    boxes = _f32(boxes)
    probs = _f32(probs)
    N, n, K = probs.shape
    T = int(post_nms_topk)
    ob = np.zeros((N, T, 4), np.float32)
    os_ = np.zeros((N, T), np.float32)
    oc = np.zeros((N, T), np.int64)
    ov = np.zeros((N, T), np.uint8)
    on = np.zeros(N, np.int32)
    lib().orc_yolo_inference(_p(boxes), _p(probs), N, C.c_int64(n), K, C.c_float(score_thresh), C.c_float(nms_thresh),
                             T, _p(ob), _p(os_), _p(oc), _p(ov), _p(on))
    return ob, os_, oc, ov.astype(bool), on


def point_nms(x):
    """lib/modeling/single_stage_heads/solo_v2.py:29-40 on NHWC scores."""
    x = _f32(x)
    N, H, W, Cc = x.shape
    out = np.empty_like(x)
    lib().orc_point_nms(_p(x), N, H, W, Cc, _p(out))
    return out


def solo_mask_stage(logits, mask_threshold=0.5):
    """solo_v2.py:513-517,530-533: logits [n,H,W] -> (masks fp32 0/1 [n,H,W], sum_masks [n], score_sums [n])."""
    x = _f32(logits)
    n = x.shape[0]
    hw = int(np.prod(x.shape[1:]))
    masks = np.empty_like(x)
    sm = np.empty(n, np.float32)
    ss = np.empty(n, np.float32)
    lib().orc_solo_mask_stage(_p(x), C.c_int64(n), C.c_int64(hw), C.c_float(mask_threshold), _p(masks), _p(sm), _p(ss))
    return masks, sm, ss


def solo_dynamic_conv(features, kernels):
    """solo_v2.py:499-511 for one image: features [H,W,E] (or [hw,E]), kernels [n,E] -> (logits [n,hw],
    absum [n,hw] = sum_k |kernel_k * feature_k|, the scale of the GPU tolerance)."""
    f = _f32(features)
    E = f.shape[-1]
    f = f.reshape(-1, E)
    k = _f32(kernels).reshape(-1, E)
    n, hw = k.shape[0], f.shape[0]
    out = np.empty((n, hw), np.float32)
    ab = np.empty((n, hw), np.float32)
    lib().orc_solo_dynamic_conv(_p(f), _p(k), C.c_int64(n), C.c_int64(hw), C.c_int64(E), _p(out), _p(ab))
    return out, ab


def sigmoid_array(x):
    """Elementwise orc_sigmoidf over an array (same function as `sigmoidf`, without the Python loop)."""
    x = _f32(x)
    out = np.empty_like(x)
    lib().orc_sigmoid_array(_p(x), C.c_int64(x.size), _p(out))
    return out


def solo_postprocess(mask_logits, scores, classes, strides, mask_threshold=0.5, pre_nms_topk=500, kernel="gaussian",
                     sigma=2.0, update_score_threshold=0.05, max_detections=100):
    """solo_v2.py:507-558 for one image -> (masks [D,H,W] fp32 0/1, classes int64 [D], scores [D], valid [D], n)."""
    x = _f32(mask_logits)
    n, H, W = x.shape
    D = int(max_detections)
    om = np.zeros((D, H, W), np.float32)
    oc = np.zeros(D, np.int64)
    os_ = np.zeros(D, np.float32)
    ov = np.zeros(D, np.uint8)
    f = lib().orc_solo_postprocess
    f.restype = C.c_int
    nv = f(_p(x), _p(_f32(scores)), _p(np.ascontiguousarray(classes, np.int64)), _p(_f32(strides)), n, C.c_int64(H * W),
           C.c_float(mask_threshold), int(pre_nms_topk), {"gaussian": 0, "linear": 1}[kernel], C.c_float(sigma),
           C.c_float(update_score_threshold), D, _p(om), _p(oc), _p(os_), _p(ov))
    return om, oc, os_, ov.astype(bool), int(nv)


def solo_upsample_boxes(masks, image_hw, align_corners=False, mask_threshold=0.5):
    """solo_v2.py:599-627 for one image: masks [D,h,w] fp32 0/1 -> (image masks uint8 [D,H,W], boxes [D,4] yxyx).
    align_corners=False: tf.compat.v2.image.resize (half-pixel centres); True: tf.image.resize_images(align_corners=True)."""
    m = _f32(masks)
    D, h, w = m.shape
    H, W = int(image_hw[0]), int(image_hw[1])
    out = np.empty((D, H, W), np.uint8)
    boxes = np.empty((D, 4), np.float32)
    lib().orc_solo_upsample_boxes(_p(m), C.c_int64(D), h, w, H, W, int(bool(align_corners)), C.c_float(mask_threshold),
                                  _p(out), _p(boxes))
    return out, boxes


def solo_select(pred_scores, pred_kernels, num_grids, strides, score_threshold):
    """solo_v2.py:481-497 for one image: pred_scores [G, K], pred_kernels [G, E] ->
    (scores [n], classes int64 [n], kernels [n, E], strides [n]) in tf.where (row-major) order."""
    sc = _f32(pred_scores)
    keep = np.argwhere(sc > np.float32(score_threshold))
    cell = np.concatenate([np.full(g * g, s_, np.float32) for g, s_ in zip(num_grids, strides)])
    return sc[keep[:, 0], keep[:, 1]], keep[:, 1].astype(np.int64), _f32(pred_kernels)[keep[:, 0]], cell[keep[:, 0]]


def mask_rcnn_inference(pred_mask_logits, pred_classes):
    """lib/modeling/roi_heads/mask_head.py:71-103: logits [M,Hm,Wm,C] NHWC, classes [M] -> sigmoid of the class channel
    [M,Hm,Wm] (class-agnostic C == 1: channel 0)."""
    x = _f32(pred_mask_logits)
    M, Hm, Wm, Cc = x.shape
    c = np.zeros(M, np.int64) if Cc == 1 else np.asarray(pred_classes, np.int64)
    sel = x[np.arange(M), :, :, np.clip(c, 0, Cc - 1)]
    sel = np.where(((c >= 0) & (c < Cc))[:, None, None], sel, np.float32(0))
    return sigmoid_array(sel)


def subsample_labels(labels, num_samples, positive_fraction, bg_label, seed=0, image=0):
    """lib/modeling/sampling.py:6-45 with the documented counter-based generator of csrc/sampling.cu restated in numpy
    (TF's random_shuffle has no defined bit pattern; see that file).  labels [P] -> (pos_idx, neg_idx) int64."""
    labels = np.asarray(labels, np.int64)
    P = labels.shape[0]
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def scores(stream):
        with np.errstate(over="ignore"):
            i = np.arange(P, dtype=np.uint64)
            z = (np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(image) * np.uint64(0x9E3779B97F4A7C15) +
                 i * np.uint64(0xBF58476D1CE4E5B9) + np.uint64(stream) * np.uint64(0x94D049BB133111EB)) & M
            z = (z + np.uint64(0x9E3779B97F4A7C15)) & M
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
            z = z ^ (z >> np.uint64(31))
        return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)

    is_pos = (labels != -1) & (labels != bg_label)
    is_neg = labels == bg_label
    num_pos = min(int(is_pos.sum()), int(num_samples * positive_fraction))
    num_neg = min(int(is_neg.sum()), num_samples - num_pos)
    k = min(int(num_samples), P)
    out = []
    for mask, stream, n in ((is_pos, 0, num_pos), (is_neg, 1, num_neg)):
        if P == 0 or n == 0:
            out.append(np.zeros(0, np.int64))
            continue
        sc = np.where(mask, scores(stream), np.float32(-np.inf)).astype(np.float32)
        _, idx = top_k(sc, k)
        out.append(np.asarray(idx[:n], np.int64))
    return out[0], out[1]
