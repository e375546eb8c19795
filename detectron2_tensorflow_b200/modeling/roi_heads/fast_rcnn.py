"""Drop-in for the inference half of lib/modeling/roi_heads/fast_rcnn.py:
`fast_rcnn_inference` (:28-187) and `FastRCNNOutputs.predict_boxes/predict_probs/inference` (:359-395)."""
import torch

from ... import _native as nv
from ...structures import BoxList


def fast_rcnn_inference(boxes, scores, proposals, score_thresh, nms_thresh, topk_per_image, nms_cls_agnostic,
                        pred_proposal_deltas=None, box2box_transform=None):
    """
    Postprocess predicted boxes: clip, score threshold, class-offset NMS, keep top-k, zero-pad.

    Args:
        boxes (Tensor): (M, K*4) class-specific or (M, 4) class-agnostic predicted boxes -- or None with
            `pred_proposal_deltas` (M, K*4 | 4) and `box2box_transform` given: the boxes are then
            `box2box_transform.apply_deltas(pred_proposal_deltas, proposals.data.boxes)` decoded inside the kernels
            (what `FastRCNNOutputs.inference` computes, fast_rcnn.py:381-395; identical values, no [M, K*4] tensor).
        scores (Tensor): (M, K+1) class probabilities (last column = background).
        proposals (SparseBoxList): `.indices` [M,2] int64, `.dense_shape` (N, Rmax), tracking 'image_shape'.
    Returns:
        (BoxList, kept_indices): dense BoxList boxes [N,topk,4], scores [N,topk], pred_classes int64 [N,topk],
        is_valid [N,topk]; kept_indices int32 [N,topk] = ROI slot of each detection (-1 padding).
    """
    fused = boxes is None
    if fused and (pred_proposal_deltas is None or box2box_transform is None):
        raise ValueError("fast_rcnn_inference: boxes=None needs pred_proposal_deltas and box2box_transform")
    host = not scores.is_cuda
    dev = nv.device_of(scores) if fused else nv.device_of(boxes, scores)
    s = nv.to_device(scores, dev, torch.float32)
    M, K1 = s.shape
    K = K1 - 1
    b = nv.to_device(pred_proposal_deltas if fused else boxes, dev, torch.float32).reshape(M, -1)
    Kb = b.shape[1] // 4
    pb = nv.to_device(proposals.data.boxes, dev, torch.float32).reshape(M, 4) if fused else None
    idx = nv.to_device(proposals.indices, dev, torch.int64).reshape(M, 2)
    N, Rmax = int(proposals.dense_shape[0]), int(proposals.dense_shape[1])
    image_shapes = proposals.get_tracking('image_shape')
    shapes = nv.to_device(image_shapes, dev, torch.int32).reshape(N, 2)
    T = int(topk_per_image)
    ob = torch.empty((N, T, 4), dtype=torch.float32, device=dev)
    os_ = torch.empty((N, T), dtype=torch.float32, device=dev)
    oc = torch.empty((N, T), dtype=torch.int64, device=dev)
    ov = torch.empty((N, T), dtype=torch.bool, device=dev)
    oroi = torch.empty((N, T), dtype=torch.int32, device=dev)
    p = nv.FastRcnnParams()
    p.boxes, p.scores, p.indices = (None if fused else b.data_ptr()), s.data_ptr(), idx.data_ptr()
    if fused:
        p.deltas, p.proposal_boxes = b.data_ptr(), pb.data_ptr()
        for i in range(4):
            p.weights[i] = float(box2box_transform.weights[i])
        p.scale_clamp = float(box2box_transform.scale_clamp)
    p.num_preds, p.num_images, p.rmax = M, N, Rmax
    p.num_bbox_reg_classes, p.num_classes = Kb, K
    p.image_shapes = shapes.data_ptr()
    p.score_thresh, p.nms_thresh = float(score_thresh), float(nms_thresh)
    p.topk_per_image = T
    p.nms_cls_agnostic = int(bool(nms_cls_agnostic))
    p.out_boxes, p.out_scores, p.out_classes = ob.data_ptr(), os_.data_ptr(), oc.data_ptr()
    p.out_valid, p.out_roi_index = ov.data_ptr(), oroi.data_ptr()
    p.out_num = None
    p.out_nms_boxes_in = None
    nv.call("fast_rcnn_postprocess", p, dev)
    if host:
        ob, os_, oc, ov, oroi = nv.to_host(ob), nv.to_host(os_), nv.to_host(oc), nv.to_host(ov), nv.to_host(oroi)
    result = BoxList(ob)
    result.add_field('scores', os_)
    result.add_field('pred_classes', oc)
    result.add_field('is_valid', ov)
    result.set_tracking('image_shape', image_shapes)
    return result, oroi


class FastRCNNOutputs(object):
    """Inference-side subset of `FastRCNNOutputs` (fast_rcnn.py:190-395)."""

    def __init__(self, box2box_transform, pred_class_logits, pred_proposal_deltas, proposals):
        self.box2box_transform = box2box_transform
        self.pred_class_logits = pred_class_logits
        self.pred_proposal_deltas = pred_proposal_deltas
        self.proposals = proposals

    def predict_boxes(self):
        return self.box2box_transform.apply_deltas(self.pred_proposal_deltas, self.proposals.data.boxes)

    def predict_probs(self):
        return torch.softmax(self.pred_class_logits, dim=-1)

    def inference(self, score_thresh, nms_thresh, topk_per_image, nms_cls_agnostic):
        """fast_rcnn.py:381-395; the box decode of `predict_boxes` runs inside the post-processing kernels."""
        return fast_rcnn_inference(None, self.predict_probs(), self.proposals, score_thresh, nms_thresh,
                                   topk_per_image, nms_cls_agnostic, pred_proposal_deltas=self.pred_proposal_deltas,
                                   box2box_transform=self.box2box_transform)
