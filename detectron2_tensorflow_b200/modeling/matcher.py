"""Drop-in for lib/modeling/matcher.py (`Matcher`, :8-174) plus the fused label assignment
(`label_boxes`) that replaces pairwise_iou -> Matcher -> inside_window -> get_deltas of
RPNOutputs._get_ground_truth (rpn_outputs.py:245-304) and of
ROIHeads.label_and_sample_proposals (roi_heads.py:100-165) without writing the IoU matrices."""
import torch

from .. import _native as nv

__all__ = ["Matcher", "label_boxes"]


def _fill_thresholds(p, thresholds, labels):
    assert 1 <= len(thresholds) <= nv.MATCH_MAX_THRESHOLDS
    p.num_thresholds = len(thresholds)
    for i, t in enumerate(thresholds):
        p.thresholds[i] = float(t)
    for i, l in enumerate(labels):
        p.labels[i] = int(l)


class Matcher(object):
    """
    Assigns to each predicted element a ground-truth element from the MxN match_quality_matrix
    (M ground truth x N predictions): `matches` [N] int64 = argmax over M, `match_labels` [N] int64 in
    {-1, 0, 1} by the threshold intervals, optionally promoted by the low-quality rule.
    """

    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        # matcher.py:44-54: add -inf / +inf, check ordering and label values
        thresholds = thresholds[:]
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = labels
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix, crowd_matrix=None, difficult_matrix=None):
        assert match_quality_matrix.dim() == 2
        host = not match_quality_matrix.is_cuda
        dev = nv.device_of(match_quality_matrix)
        q = nv.to_device(match_quality_matrix, dev, torch.float32)
        M, N = q.shape
        cm = None if crowd_matrix is None else nv.to_device(crowd_matrix, dev, torch.float32)
        dm = None if difficult_matrix is None else nv.to_device(difficult_matrix, dev, torch.float32)
        matches = torch.empty(N, dtype=torch.int64, device=dev)
        labels = torch.empty(N, dtype=torch.int64, device=dev)
        p = nv.MatcherParams()
        p.match_quality_matrix = q.data_ptr() if M > 0 else None
        p.crowd_matrix = cm.data_ptr() if cm is not None and cm.shape[0] > 0 else None
        p.difficult_matrix = dm.data_ptr() if dm is not None and dm.shape[0] > 0 else None
        p.num_gt = M
        p.num_crowd = 0 if cm is None else cm.shape[0]
        p.num_difficult = 0 if dm is None else dm.shape[0]
        p.use_crowd, p.use_difficult = int(cm is not None), int(dm is not None)
        p.num_preds = N
        _fill_thresholds(p, self.thresholds[1:-1], self.labels)
        p.allow_low_quality_matches = int(bool(self.allow_low_quality_matches))
        p.out_matches, p.out_labels = matches.data_ptr(), labels.data_ptr()
        nv.call("matcher", p, dev)
        if host:
            return nv.to_host(matches), nv.to_host(labels)
        return matches, labels


def label_boxes(pred_boxes, gt_boxes, gt_valid, matcher, gt_crowd=None, gt_difficult=None, pred_counts=None,
                boundary_threshold=-1, image_shapes=None, box2box_transform=None):
    """Fused label assignment for a whole batch.

    pred_boxes: [P,4] (anchors shared by all images) or [N,P,4] (+ optional pred_counts [N] valid prefix)
    gt_boxes [N,G,4], gt_valid / gt_crowd / gt_difficult [N,G] bool
    -> (matches [N,P] int64 into the image's valid-GT list, labels [N,P] int64,
        gt deltas [N,P,4] when `box2box_transform` is given, else None)
    """
    host = not pred_boxes.is_cuda
    dev = nv.device_of(pred_boxes, gt_boxes)
    pb = nv.to_device(pred_boxes, dev, torch.float32)
    shared = pb.dim() == 2
    gt = nv.to_device(gt_boxes, dev, torch.float32)
    N, G = gt.shape[0], gt.shape[1]
    P = pb.shape[-2]
    u8 = lambda t: None if t is None else nv.to_device(t, dev).to(torch.uint8).contiguous()
    v, c, d = u8(gt_valid), u8(gt_crowd), u8(gt_difficult)
    pc = None if pred_counts is None else nv.to_device(pred_counts, dev, torch.int32)
    sh = None if image_shapes is None else nv.to_device(image_shapes, dev, torch.int32)
    matches = torch.empty((N, P), dtype=torch.int64, device=dev)
    labels = torch.empty((N, P), dtype=torch.int64, device=dev)
    deltas = torch.empty((N, P, 4), dtype=torch.float32, device=dev) if box2box_transform is not None else None
    p = nv.LabelBoxesParams()
    p.pred_boxes, p.pred_shared, p.pred_counts = pb.data_ptr(), int(shared), nv.ptr(pc)
    p.num_images, p.num_preds = N, P
    p.gt_boxes, p.gt_valid, p.gt_crowd, p.gt_difficult = gt.data_ptr(), nv.ptr(v), nv.ptr(c), nv.ptr(d)
    p.max_gt = G
    _fill_thresholds(p, matcher.thresholds[1:-1], matcher.labels)
    p.allow_low_quality_matches = int(bool(matcher.allow_low_quality_matches))
    p.boundary_threshold = float(boundary_threshold)
    p.image_shapes = nv.ptr(sh)
    p.compute_deltas = int(deltas is not None)
    if box2box_transform is not None:
        for i in range(4):
            p.weights[i] = float(box2box_transform.weights[i])
    p.out_matches, p.out_labels, p.out_deltas = matches.data_ptr(), labels.data_ptr(), nv.ptr(deltas)
    nv.call("label_boxes", p, dev)
    if host:
        return nv.to_host(matches), nv.to_host(labels), None if deltas is None else nv.to_host(deltas)
    return matches, labels, deltas
