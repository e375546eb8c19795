"""Drop-in for the hot part of lib/structures/box_list_ops.py: `pairwise_iou` (:295-371, every iou_type)."""
import torch

from .. import _native as nv

__all__ = ["pairwise_iou"]


def pairwise_iou(boxlist1, boxlist2, iou_type='iou', scope=None):
    """[N, M] IoU between two box collections (BoxList or [n,4] tensors); iou_type 'iou' (the matching path) or
    'giou' / 'diou' / 'ciou' (the YOLOv4 losses, :335-371; anything else falls through to plain IoU in the
    reference -- here it is an error)."""
    if iou_type not in nv.IOU_TYPES:
        raise ValueError(f"iou_type '{iou_type}' is not recognized.")
    b1 = boxlist1.boxes if hasattr(boxlist1, "boxes") else boxlist1
    b2 = boxlist2.boxes if hasattr(boxlist2, "boxes") else boxlist2
    host = not b1.is_cuda
    dev = nv.device_of(b1, b2)
    b1 = nv.to_device(b1, dev, torch.float32).reshape(-1, 4)
    b2 = nv.to_device(b2, dev, torch.float32).reshape(-1, 4)
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=dev)
    p = nv.PairwiseIouParams()
    p.boxes1, p.boxes2, p.n1, p.n2, p.out = b1.data_ptr(), b2.data_ptr(), b1.shape[0], b2.shape[0], out.data_ptr()
    p.iou_type = nv.IOU_TYPES[iou_type]
    nv.call("pairwise_iou", p, dev)
    return nv.to_host(out) if host else out
