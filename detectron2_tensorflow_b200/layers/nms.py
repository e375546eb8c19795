"""Drop-in for lib/layers/nms.py: `batch_nms` (:6-26) and `matrix_nms` (:29-83)."""
import torch

from .. import _native as nv


def batch_nms(boxes, scores, max_output_size, axis=0, iou_threshold=0.5, scope=None):
    """tf.image.non_max_suppression per batch row.

    boxes [B, n, 4] / scores [B, n] when axis=1... the reference transposes when ``axis == 0``
    (inputs given as [n, B, 4] / [n, B]).  The reference's ``tf.map_fn`` cannot stack ragged keep
    lists; this version returns ``(keep [B, max_output_size] int32 padded with -1,
    num_keep [B] int32)`` -- the documented deviation.
    """
    assert boxes.dim() == 3
    assert scores.dim() == 2
    assert axis in [0, 1]
    host = not boxes.is_cuda
    dev = nv.device_of(boxes, scores)
    b = nv.to_device(boxes, dev, torch.float32)
    s = nv.to_device(scores, dev, torch.float32)
    if axis == 0:
        b = b.permute(1, 0, 2).contiguous()
        s = s.permute(1, 0).contiguous()
    B, n = s.shape
    keep = torch.empty((B, max_output_size), dtype=torch.int32, device=dev)
    num = torch.empty(B, dtype=torch.int32, device=dev)
    p = nv.BatchedNmsParams()
    p.boxes, p.scores, p.counts = b.data_ptr(), s.data_ptr(), None
    p.num_segments, p.n, p.max_output_size = B, n, int(max_output_size)
    p.iou_threshold = float(iou_threshold)
    p.keep, p.num_keep = keep.data_ptr(), num.data_ptr()
    nv.call("batched_nms", p, dev)
    if host:
        return nv.to_host(keep), nv.to_host(num)
    return keep, num


def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5):
    """Single-segment convenience with tf.image.non_max_suppression's return value (kept indices)."""
    keep, num = batch_nms(boxes[None], scores[None], max_output_size, axis=1, iou_threshold=iou_threshold)
    return keep[0, :int(num[0])]


def matrix_nms(masks, classes, scores, sum_masks=None, kernel="gaussian", sigma=2.0, scope=None, packed_masks=None,
               mask_hw=None):
    """Matrix NMS of SOLOv2 (lib/layers/nms.py:29-83).

    masks [n, H, W] binary fp32 (sorted by score desc, solo_v2.py:536), classes [n] int64, scores [n]
    -> updated scores [n].  A leading batch dimension on all inputs ([B, n, H, W] ...) runs B images
    in one launch (the reference loops with tf.map_fn, solo_v2.py:587).

    `packed_masks` ([n, ceil(hw/64)] int64 words from `solo_mask_encode`, with `mask_hw` = H*W) replaces
    `masks` (pass masks=None): the fp32 0/1 masks are then never read.
    """
    if kernel == "gaussian":
        kid = nv.MNMS_GAUSSIAN
    elif kernel == "linear":
        kid = nv.MNMS_LINEAR
    else:
        raise NotImplementedError(f"NMS kernel {kernel} not implemented yet.")
    if packed_masks is not None:
        assert masks is None and mask_hw is not None
        batched = packed_masks.dim() == 3
        host = not packed_masks.is_cuda
        dev = nv.device_of(packed_masks, scores)
        m = None
        pk = nv.to_device(packed_masks, dev, torch.int64)
    else:
        batched = masks.dim() == 4
        assert masks.dim() in (3, 4)
        host = not masks.is_cuda
        dev = nv.device_of(masks, scores)
        m = nv.to_device(masks, dev, torch.float32)
        pk = None
    assert classes.dim() == (2 if batched else 1)
    assert scores.dim() == (2 if batched else 1)
    c = nv.to_device(classes, dev, torch.int64)
    s = nv.to_device(scores, dev, torch.float32)
    sm = None if sum_masks is None else nv.to_device(sum_masks, dev, torch.float32)
    if not batched:
        m, c, s = (None if m is None else m[None]), c[None], s[None]
        pk = None if pk is None else pk[None]
        sm = None if sm is None else sm[None]
    B, n = s.shape
    out = torch.empty((B, n), dtype=torch.float32, device=dev)
    p = nv.MatrixNmsParams()
    p.masks, p.classes, p.scores = nv.ptr(m), c.data_ptr(), s.data_ptr()
    p.packed_masks = nv.ptr(pk)
    p.sum_masks = nv.ptr(sm)
    p.counts = None
    p.batch, p.n = B, n
    p.hw = int(mask_hw) if m is None else int(m.shape[2] * m.shape[3])
    p.kernel, p.sigma = kid, float(sigma)
    p.out = out.data_ptr()
    nv.call("matrix_nms", p, dev)
    if not batched:
        out = out[0]
    return nv.to_host(out) if host else out
