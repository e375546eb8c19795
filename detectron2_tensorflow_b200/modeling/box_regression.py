"""Drop-in for lib/modeling/box_regression.py: `Box2BoxTransform.apply_deltas` (:76-123)."""
import math

import torch

from .. import _native as nv

# Value for clamping large dw and dh predictions (box_regression.py:10)
_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)

__all__ = ["Box2BoxTransform"]


class Box2BoxTransform(object):
    """
    The box-to-box transform defined in R-CNN, parameterized by 4 deltas (dy, dx, dh, dw).
    `apply_deltas` is on the inference hot path; `get_deltas` builds training targets (:38-74).
    """

    def __init__(self, weights, scale_clamp=_DEFAULT_SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        """Deltas (dy, dx, dh, dw) that transform `src_boxes` into `target_boxes` (both (N, 4))."""
        host = not src_boxes.is_cuda
        dev = nv.device_of(src_boxes, target_boxes)
        s = nv.to_device(src_boxes, dev, torch.float32).reshape(-1, 4)
        t = nv.to_device(target_boxes, dev, torch.float32).reshape(-1, 4)
        assert s.shape == t.shape
        out = torch.empty_like(s)
        p = nv.GetDeltasParams()
        p.src_boxes, p.target_boxes, p.n, p.out = s.data_ptr(), t.data_ptr(), s.shape[0], out.data_ptr()
        for i in range(4):
            p.weights[i] = float(self.weights[i])
        nv.call("get_deltas", p, dev)
        return nv.to_host(out) if host else out

    def apply_deltas(self, deltas, boxes):
        """
        Args:
            deltas (Tensor): (N, k*4); deltas[i] holds k class-specific transforms for boxes[i].
            boxes (Tensor): (N, 4)
        """
        host = not deltas.is_cuda
        dev = nv.device_of(deltas, boxes)
        d = nv.to_device(deltas, dev, torch.float32)
        b = nv.to_device(boxes, dev, torch.float32).reshape(-1, 4)
        n = b.shape[0]
        d2 = d.reshape(n, -1)
        assert d2.shape[1] % 4 == 0
        out = torch.empty_like(d2)
        p = nv.ApplyDeltasParams()
        p.deltas, p.boxes, p.n, p.k = d2.data_ptr(), b.data_ptr(), n, d2.shape[1] // 4
        for i in range(4):
            p.weights[i] = float(self.weights[i])
        p.scale_clamp = float(self.scale_clamp)
        p.out = out.data_ptr()
        nv.call("apply_deltas", p, dev)
        out = out.reshape(d.shape)
        return nv.to_host(out) if host else out
