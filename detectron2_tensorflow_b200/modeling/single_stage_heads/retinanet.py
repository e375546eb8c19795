"""Drop-in for `RetinaNetHead.inference` (lib/modeling/single_stage_heads/retinanet.py:285-387)."""
import torch

from ... import _native as nv
from ...structures import BoxList
from ..box_regression import Box2BoxTransform


class RetinaNetInference(object):
    """Holds the inference attributes `RetinaNetHead.__init__` reads from cfg (retinanet.py:70-95)."""

    def __init__(self, num_classes=80, topk_candidates=1000, score_threshold=0.05, nms_threshold=0.5,
                 max_detections_per_image=100, bbox_reg_weights=(1.0, 1.0, 1.0, 1.0)):
        self.num_classes = num_classes
        self.topk_candidates = topk_candidates
        self.score_threshold = score_threshold
        self.nms_threshold = nms_threshold
        self.max_detections_per_image = max_detections_per_image
        self.box2box_transform = Box2BoxTransform(weights=bbox_reg_weights)

    def inference(self, box_cls, box_delta, anchors):
        """
        Arguments:
            box_cls: L tensors (N, Hi, Wi, A*K) (or (N, Hi*Wi*A, K)) class logits.
            box_delta: L tensors (N, Hi, Wi, A*4) (or (N, Hi*Wi*A, 4)).
            anchors: L BoxLists / tensors (Hi*Wi*A, 4).
        Returns:
            BoxList: boxes [N,100,4], scores [N,100], pred_classes int32 [N,100], is_valid [N,100].
        """
        K = self.num_classes
        dev = nv.device_of(*box_cls)
        host = not box_cls[0].is_cuda
        N = box_cls[0].shape[0]
        cls = [nv.to_device(x, dev, torch.float32).reshape(N, -1, K) for x in box_cls]   # reshape_to_N_HWA_K :371
        dl = [nv.to_device(x, dev, torch.float32).reshape(N, -1, 4) for x in box_delta]  # :372
        from ..anchor_generator import GridAnchors
        an = [a if isinstance(a, GridAnchors) else
              nv.to_device(a.boxes if hasattr(a, "boxes") else a, dev, torch.float32).reshape(-1, 4) for a in anchors]
        keep_alive = []
        L = len(cls)
        T = int(self.max_detections_per_image)
        ob = torch.empty((N, T, 4), dtype=torch.float32, device=dev)
        os_ = torch.empty((N, T), dtype=torch.float32, device=dev)
        oc = torch.empty((N, T), dtype=torch.int32, device=dev)
        ov = torch.empty((N, T), dtype=torch.bool, device=dev)
        p = nv.RetinanetParams()
        for l in range(L):
            assert cls[l].shape[1] == dl[l].shape[1]
            p.box_cls[l], p.box_delta[l] = cls[l].data_ptr(), dl[l].data_ptr()
            if isinstance(an[l], GridAnchors):
                assert an[l].num_anchors == cls[l].shape[1]
                cell = nv.to_device(an[l].cell_anchors, dev, torch.float32).reshape(-1, 4)
                keep_alive.append(cell)
                p.cell_anchors[l], p.num_cell_anchors[l] = cell.data_ptr(), cell.shape[0]
                p.grid_w[l], p.stride[l] = an[l].grid_hw[1], an[l].stride
            else:
                assert an[l].shape[0] == cls[l].shape[1]
                p.anchors[l] = an[l].data_ptr()
            p.hwa[l] = cls[l].shape[1]
        p.num_levels, p.num_images, p.num_classes = L, N, K
        p.topk_candidates = int(self.topk_candidates)
        p.score_thresh, p.nms_thresh = float(self.score_threshold), float(self.nms_threshold)
        p.max_detections = T
        for i in range(4):
            p.weights[i] = float(self.box2box_transform.weights[i])
        p.scale_clamp = float(self.box2box_transform.scale_clamp)
        p.out_boxes, p.out_scores, p.out_classes, p.out_valid = ob.data_ptr(), os_.data_ptr(), oc.data_ptr(), ov.data_ptr()
        p.out_num = None
        p.out_nms_boxes_in = None
        nv.call("retinanet_postprocess", p, dev)
        if host:
            ob, os_, oc, ov = nv.to_host(ob), nv.to_host(os_), nv.to_host(oc), nv.to_host(ov)
        result = BoxList(ob)
        result.add_field('scores', os_)
        result.add_field('pred_classes', oc)
        result.add_field('is_valid', ov)
        return result
