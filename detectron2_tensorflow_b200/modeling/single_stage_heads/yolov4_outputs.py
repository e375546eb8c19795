"""Drop-in for the inference tail of lib/modeling/single_stage_heads/yolov4_outputs.py
(`YOLOv4Outputs.inference`, :331-390): per image max/argmax over classes, score threshold, one
class-agnostic NMS, zero-padded outputs -- one C-ABI call for the whole batch."""
import torch

from ... import _native as nv
from ...structures import BoxList

__all__ = ["YOLOv4Inference"]


class YOLOv4Inference(object):
    """Holds the attributes `YOLOv4Outputs.inference` reads (score_threshold, nms_threshold, post_nms_topk)."""

    def __init__(self, score_threshold=0.05, nms_threshold=0.5, post_nms_topk=100):
        self.score_threshold = score_threshold
        self.nms_threshold = nms_threshold
        self.post_nms_topk = post_nms_topk

    def inference(self, predicted_boxes, predicted_prob):
        """
        Args:
            predicted_boxes: (N, n, 4) decoded boxes (`_get_predictions()[0]`)
            predicted_prob: (N, n, K) class probabilities (`_get_predictions()[2]`)
        Returns:
            BoxList: boxes [N,post,4], scores [N,post], pred_classes int64 [N,post], is_valid [N,post].
        """
        host = not predicted_boxes.is_cuda
        dev = nv.device_of(predicted_boxes, predicted_prob)
        b = nv.to_device(predicted_boxes, dev, torch.float32)
        pr = nv.to_device(predicted_prob, dev, torch.float32)
        assert b.dim() == 3 and pr.dim() == 3 and b.shape[:2] == pr.shape[:2] and b.shape[2] == 4
        N, n, K = pr.shape
        T = int(self.post_nms_topk)
        ob = torch.empty((N, T, 4), dtype=torch.float32, device=dev)
        os_ = torch.empty((N, T), dtype=torch.float32, device=dev)
        oc = torch.empty((N, T), dtype=torch.int64, device=dev)
        ov = torch.empty((N, T), dtype=torch.bool, device=dev)
        p = nv.YoloParams()
        p.boxes, p.probs = b.data_ptr(), pr.data_ptr()
        p.num_images, p.num_boxes, p.num_classes = N, n, K
        p.score_thresh, p.nms_thresh, p.post_nms_topk = float(self.score_threshold), float(self.nms_threshold), T
        p.out_boxes, p.out_scores, p.out_classes, p.out_valid = ob.data_ptr(), os_.data_ptr(), oc.data_ptr(), ov.data_ptr()
        p.out_num = None
        p.out_nms_boxes_in = None
        nv.call("yolo_postprocess", p, dev)
        if host:
            ob, os_, oc, ov = nv.to_host(ob), nv.to_host(os_), nv.to_host(oc), nv.to_host(ov)
        result = BoxList(ob)
        result.add_field('scores', os_)
        result.add_field('pred_classes', oc)
        result.add_field('is_valid', ov)
        return result
