"""Drop-in for `point_nms` of lib/modeling/single_stage_heads/solo_v2.py (:29-40)."""
import torch

from ... import _native as nv

__all__ = ["point_nms", "solo_mask_encode"]


def point_nms(inputs, kernel_size=2, scope=None):
    """NHWC category scores: a value survives iff it equals the max of its 2x2 (self, up, left, up-left) window."""
    assert kernel_size == 2
    host = not inputs.is_cuda
    dev = nv.device_of(inputs)
    x = nv.to_device(inputs, dev, torch.float32)
    assert x.dim() == 4
    out = torch.empty_like(x)
    p = nv.PointNmsParams()
    p.scores, p.out = x.data_ptr(), out.data_ptr()
    p.num_images, p.height, p.width, p.channels = x.shape
    nv.call("point_nms", p, dev)
    return nv.to_host(out) if host else out


def solo_mask_encode(mask_logits, mask_threshold=0.5, counts=None):
    """The mask stage of SOLOv2Head.inference (solo_v2.py:513-517, 530-533) in one streaming pass.

    mask_logits [n, H, W] (or [B, n, H, W]): the dynamic-conv output BEFORE the sigmoid.
    Returns (packed_masks int64 [.., n, ceil(H*W/64)], sum_masks fp32 [.., n], score_sums fp32 [.., n]) where
    packed bit p of word w is `sigmoid(logit) > mask_threshold` of pixel 64*w+p, sum_masks = reduce_sum(pred_masks)
    and score_sums = reduce_sum(pred_mask_scores * pred_masks); mask_scoring (:531-533) = score_sums / sum_masks.
    Feed `packed_masks` to `layers.matrix_nms(None, ..., packed_masks=..., mask_hw=H*W)`.
    """
    host = not mask_logits.is_cuda
    dev = nv.device_of(mask_logits)
    x = nv.to_device(mask_logits, dev, torch.float32)
    batched = x.dim() == 4
    if not batched:
        x = x[None]
    B, n, H, W = x.shape
    hw = H * W
    Wd = (hw + 63) // 64
    packed = torch.empty((B, n, Wd), dtype=torch.int64, device=dev)
    sums = torch.empty((B, n), dtype=torch.float32, device=dev)
    ssum = torch.empty((B, n), dtype=torch.float32, device=dev)
    cnt = None if counts is None else nv.to_device(counts, dev, torch.int32)
    p = nv.SoloMaskEncodeParams()
    p.mask_logits, p.counts = x.data_ptr(), nv.ptr(cnt)
    p.batch, p.n, p.hw = B, n, hw
    p.mask_threshold = float(mask_threshold)
    p.packed_masks, p.sum_masks, p.score_sums = packed.data_ptr(), sums.data_ptr(), ssum.data_ptr()
    nv.call("solo_mask_encode", p, dev)
    outs = (packed, sums, ssum)
    if not batched:
        outs = tuple(o[0] for o in outs)
    return tuple(nv.to_host(o) for o in outs) if host else outs
