"""Drop-in for lib/modeling/poolers.py (`assign_boxes_to_levels` :11-49, `ROIPooler` :52-180).

The reference runs assign -> per-level where/gather -> pad -> crop_and_resize ->
[avg_pool] -> concat -> invert_permutation -> gather.  Here `ROIPooler.call` is ONE
kernel launch (d2b_roi_align_multilevel): the level is computed in-kernel and each
ROI is written straight to its final row.
"""
import math

import torch

from .. import _native as nv
from ..layers import Layer, ROIAlign
from ..layers.functional import _roi_align_call, _roi_align_backward_call

__all__ = ["ROIPooler", "assign_boxes_to_levels"]


def assign_boxes_to_levels(boxlist, min_level, max_level, canonical_box_size, canonical_level):
    """
    Map each box in `boxlist` to a feature map level index (offset from `min_level`), int64 [M].
    Eqn.(1) of the FPN paper in the reference's fp32 op order (poolers.py:37-49).
    """
    boxes = boxlist.boxes if hasattr(boxlist, "boxes") else boxlist
    host = not boxes.is_cuda
    dev = nv.device_of(boxes)
    b = nv.to_device(boxes, dev, torch.float32).reshape(-1, 4)
    M = b.shape[0]
    L = max_level - min_level + 1
    # the level computation lives in the ROIAlign kernel's prologue; run it on a 1x1x4 dummy map
    dummy = [torch.zeros((1, 1, 1, 4), dtype=torch.float32, device=dev) for _ in range(L)]
    bidx = torch.zeros(M, dtype=torch.int32, device=dev)
    _, _, levels = _roi_align_call(dummy, [2.0 ** -(min_level + l) for l in range(L)], b, bidx, 1, (1, 1), 0, True,
                                   True, min_level=min_level, canonical_box_size=canonical_box_size,
                                   canonical_level=canonical_level, want_levels=True)
    return nv.to_host(levels) if host else levels


class ROIPooler(Layer):
    """
    Region of interest feature map pooler that supports pooling from one or
    more feature maps.
    """

    def __init__(
        self,
        output_size,
        scales,
        sampling_ratio,
        pooler_type,
        canonical_box_size=224,
        canonical_level=4,
    ):
        super().__init__()

        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2
        assert isinstance(output_size[0], int) and isinstance(output_size[1], int)
        self.output_size = output_size

        if pooler_type == "ROIAlign":
            self.aligned = False
        elif pooler_type == "ROIAlignV2":
            self.aligned = True
        else:
            raise ValueError("Unknown pooler type: {}".format(pooler_type))
        self.scales = list(scales)
        self.sampling_ratio = sampling_ratio
        # kept for API parity with the reference (per-level poolers are not used by call())
        self.level_poolers = [
            ROIAlign(output_size, spatial_scale=scale, sampling_ratio=sampling_ratio, aligned=self.aligned)
            for scale in scales
        ]

        # Map scale (defined as 1 / stride) to its feature map level under the
        # assumption that stride is a power of 2.
        min_level = -math.log2(scales[0])
        max_level = -math.log2(scales[-1])
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
        self.min_level = int(min_level)
        self.max_level = int(max_level)
        assert 0 < self.min_level and self.min_level <= self.max_level
        assert self.min_level <= canonical_level and canonical_level <= self.max_level
        self.canonical_level = canonical_level
        assert canonical_box_size > 0
        self.canonical_box_size = canonical_box_size
        self.last_level_counts = None  # 'roi_align/num_roi_level_k' summaries (poolers.py:173)

    def call(self, x, instances):
        """
        Args:
            x (list[Tensor]): NHWC feature maps with scales matching those used to construct this module.
            instances (SparseBoxList): `.data.boxes` [M,4], `.indices` [M,2] int64 (column 0 = image)
        Returns:
            Tensor (M, output_h, output_w, C) in input ROI order.
        """
        num_level_assignments = len(self.level_poolers)

        assert len(x) == num_level_assignments, (
            "unequal value, num_level_assignments={}, but x is list of {} "
            "Tensors".format(num_level_assignments, len(x)))

        batch_idx = instances.indices[:, 0]
        if num_level_assignments == 1:
            return _roi_align_call(x, self.scales, instances.data.boxes, batch_idx, 1, self.output_size,
                                   self.sampling_ratio, self.aligned, True)
        out, counts, _ = _roi_align_call(
            x, self.scales, instances.data.boxes, batch_idx, 1, self.output_size, self.sampling_ratio,
            self.aligned, True, min_level=self.min_level, canonical_box_size=self.canonical_box_size,
            canonical_level=self.canonical_level, want_levels=True)
        self.last_level_counts = counts
        return out

    def backward(self, grad_output, x_shapes, instances, grad_x=None):
        """Gradient of `call` w.r.t. the feature maps `x` (list of (N,H,W,C) shapes) -> list of NHWC tensors.
        One kernel: each ROI scatters into the level `assign_boxes_to_levels` gave it."""
        assert len(x_shapes) == len(self.level_poolers)
        return _roi_align_backward_call(grad_output, x_shapes, self.scales, instances.data.boxes,
                                        instances.indices[:, 0].contiguous(), self.sampling_ratio, self.aligned,
                                        min_level=self.min_level, canonical_box_size=self.canonical_box_size,
                                        canonical_level=self.canonical_level, grad_features=grad_x)
