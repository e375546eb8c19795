"""Drop-in for `mask_rcnn_inference` of lib/modeling/roi_heads/mask_head.py (:71-103)."""
import torch

from ... import _native as nv


def mask_rcnn_inference(pred_mask_logits, pred_instances):
    """Per detection, sigmoid of the mask-head logits of ITS predicted class: logits [M, Hmask, Wmask, C] (NHWC; C = 1
    for a class-agnostic head) + `pred_instances.data['pred_classes']` -> new field `pred_masks` [M, Hmask, Wmask] on
    `pred_instances` (a SparseBoxList); returns None, as the reference function does.  One gather + sigmoid kernel instead of the
    reference's full NHWC -> NCHW transpose."""
    dev = nv.device_of(pred_mask_logits)
    x = nv.to_device(pred_mask_logits, dev, torch.float32)
    assert x.dim() == 4
    M, Hm, Wm, Cc = x.shape
    cls = None
    if Cc != 1:
        cls = nv.to_device(pred_instances.data.get_field('pred_classes'), dev, torch.int64).reshape(-1)
        assert cls.shape[0] == M
    out = torch.empty((M, Hm, Wm), dtype=torch.float32, device=dev)
    p = nv.MaskRcnnInferenceParams()
    p.mask_logits, p.pred_classes, p.num_masks = x.data_ptr(), nv.ptr(cls), M
    p.mask_h, p.mask_w, p.num_classes = Hm, Wm, Cc
    p.out = out.data_ptr()
    nv.call("mask_rcnn_inference", p, dev)
    pred_instances.data.add_field('pred_masks', out if pred_mask_logits.is_cuda else nv.to_host(out))
