"""Drop-in for the inference half of lib/modeling/proposal_generator/rpn_outputs.py:
`find_top_rpn_proposals` (:29-132) and `RPNOutputs.predict_proposals/predict_objectness_logits`
(:403-440).  The loss half (:135-401) builds training labels and is out of scope.
"""
import torch

from ... import _native as nv
from ...structures import BoxList
from ..box_regression import _DEFAULT_SCALE_CLAMP


def _rpn_call(logits, proposals, deltas, anchors, image_shapes, nms_thresh, pre_nms_topk, post_nms_topk,
              min_box_side_len, weights=(1.0, 1.0, 1.0, 1.0), scale_clamp=_DEFAULT_SCALE_CLAMP, count_nms_in=False):
    dev = nv.device_of(*logits)
    host = not logits[0].is_cuda
    L = len(logits)
    lg = [nv.to_device(x, dev, torch.float32) for x in logits]
    N = lg[0].shape[0]
    lg = [x.reshape(N, -1) for x in lg]
    pr = None if proposals is None else [nv.to_device(x, dev, torch.float32).reshape(N, -1, 4) for x in proposals]
    dl = None if deltas is None else [nv.to_device(x, dev, torch.float32).reshape(N, -1, 4) for x in deltas]
    from ..anchor_generator import GridAnchors
    an = None
    if anchors is not None:
        an = [a if isinstance(a, GridAnchors) else nv.to_device(a, dev, torch.float32).reshape(-1, 4) for a in anchors]
    keep_alive = []
    shapes = nv.to_device(image_shapes, dev, torch.int32).reshape(N, 2)
    post = int(post_nms_topk)
    out_boxes = torch.empty((N, post, 4), dtype=torch.float32, device=dev)
    out_logits = torch.empty((N, post), dtype=torch.float32, device=dev)
    out_valid = torch.empty((N, post), dtype=torch.bool, device=dev)
    nms_in = torch.zeros(1, dtype=torch.int64, device=dev) if count_nms_in else None
    p = nv.RpnProposalsParams()
    for l in range(L):
        p.logits[l] = lg[l].data_ptr()
        p.hwa[l] = lg[l].shape[1]
        if pr is not None:
            assert pr[l].shape[1] == lg[l].shape[1]
            p.proposals[l] = pr[l].data_ptr()
        if dl is not None:
            assert dl[l].shape[1] == lg[l].shape[1]
            p.deltas[l] = dl[l].data_ptr()
            if isinstance(an[l], GridAnchors):  # synthesised in-kernel (anchor_generator.py:92-109)
                assert an[l].num_anchors == lg[l].shape[1]
                cell = nv.to_device(an[l].cell_anchors, dev, torch.float32).reshape(-1, 4)
                keep_alive.append(cell)
                p.cell_anchors[l] = cell.data_ptr()
                p.num_cell_anchors[l] = cell.shape[0]
                p.grid_w[l] = an[l].grid_hw[1]
                p.stride[l] = an[l].stride
            else:
                assert an[l].shape[0] == lg[l].shape[1]
                p.anchors[l] = an[l].data_ptr()
    p.num_levels, p.num_images = L, N
    p.image_shapes = shapes.data_ptr()
    p.nms_thresh = float(nms_thresh)
    p.pre_nms_topk, p.post_nms_topk = int(pre_nms_topk), post
    p.min_box_side_len = float(min_box_side_len)
    for i in range(4):
        p.weights[i] = float(weights[i])
    p.scale_clamp = float(scale_clamp)
    p.out_boxes, p.out_logits, p.out_valid = out_boxes.data_ptr(), out_logits.data_ptr(), out_valid.data_ptr()
    p.out_num_valid = None
    p.out_nms_boxes_in = nv.ptr(nms_in)
    nv.call("rpn_proposals", p, dev)
    if host:
        out_boxes, out_logits, out_valid = nv.to_host(out_boxes), nv.to_host(out_logits), nv.to_host(out_valid)
    results = BoxList(out_boxes)
    results.add_field("objectness_logits", out_logits)
    results.add_field("is_valid", out_valid)
    results.set_tracking("image_shape", image_shapes)
    if count_nms_in:
        results.set_tracking("nms_boxes_in", nms_in)
    return results


def find_top_rpn_proposals(
    proposals,
    pred_objectness_logits,
    images,
    nms_thresh,
    pre_nms_topk,
    post_nms_topk,
    min_box_side_len,
):
    """
    For each feature map, select the `pre_nms_topk` highest scoring proposals, clip them, remove
    small boxes, apply NMS, then keep the `post_nms_topk` best per image (rpn_outputs.py:29-132).

    Args:
        proposals (list[Tensor]): L tensors (N, Hi*Wi*A, 4).
        pred_objectness_logits (list[Tensor]): L tensors (N, Hi*Wi*A).
        images (ImageList): `.image_shapes` [N,2] in (h, w) order.
    Returns:
        BoxList: boxes [N,post,4], objectness_logits [N,post], is_valid [N,post]; tracking image_shape.
    """
    return _rpn_call(pred_objectness_logits, proposals, None, None, images.image_shapes, nms_thresh, pre_nms_topk,
                     post_nms_topk, min_box_side_len)


class RPNOutputs(object):
    """Inference-side subset of `RPNOutputs` (rpn_outputs.py:229-440)."""

    def __init__(self, box2box_transform, images, pred_objectness_logits, pred_anchor_deltas, anchors):
        """
        pred_objectness_logits: L tensors (N, Hi, Wi, A); pred_anchor_deltas: L tensors (N, Hi, Wi, A*4);
        anchors: L BoxLists / tensors (Hi*Wi*A, 4).
        """
        self.box2box_transform = box2box_transform
        self.images = images
        self.pred_objectness_logits = pred_objectness_logits
        self.pred_anchor_deltas = pred_anchor_deltas
        # BoxLists / tensors of materialised anchors, or GridAnchors descriptors (synthesised in-kernel)
        self.anchors = [a.boxes if hasattr(a, "boxes") else a for a in anchors]
        self.num_images = pred_objectness_logits[0].shape[0]

    def predict_proposals(self):
        """Decode ALL anchors (rpn_outputs.py:403-426) -> L tensors (N, Hi*Wi*A, 4)."""
        out = []
        for anchors_i, deltas_i in zip(self.anchors, self.pred_anchor_deltas):
            if hasattr(anchors_i, "materialize"):
                anchors_i = anchors_i.materialize().to(deltas_i.device)
            N = deltas_i.shape[0]
            d = deltas_i.reshape(-1, 4)
            a = anchors_i.unsqueeze(0).expand(N, -1, -1).reshape(-1, 4)
            out.append(self.box2box_transform.apply_deltas(d, a).reshape(N, -1, 4))
        return out

    def predict_objectness_logits(self):
        """(N, Hi, Wi, A) -> (N, Hi*Wi*A) (rpn_outputs.py:428-440)."""
        return [s.reshape(s.shape[0], -1) for s in self.pred_objectness_logits]

    def find_top_proposals(self, nms_thresh, pre_nms_topk, post_nms_topk, min_box_side_len):
        """Fused predict_proposals + find_top_rpn_proposals: only the top-k winners are decoded
        (value-identical to decode-all because decoding is elementwise)."""
        return _rpn_call(self.predict_objectness_logits(), None,
                         [d.reshape(d.shape[0], -1, 4) for d in self.pred_anchor_deltas], self.anchors,
                         self.images.image_shapes, nms_thresh, pre_nms_topk, post_nms_topk, min_box_side_len,
                         weights=self.box2box_transform.weights, scale_clamp=self.box2box_transform.scale_clamp)
