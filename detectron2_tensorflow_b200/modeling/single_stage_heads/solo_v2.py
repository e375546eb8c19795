"""Drop-in for `point_nms` of lib/modeling/single_stage_heads/solo_v2.py (:29-40)."""
import torch

from ... import _native as nv

__all__ = ["point_nms"]


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
