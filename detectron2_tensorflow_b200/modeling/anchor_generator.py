"""Drop-in for `DefaultAnchorGenerator` (lib/modeling/anchor_generator.py:44-162), SURVEY.md 8f "next" #2.

Anchors are a pure function of (level, y, x, a): the RPN / RetinaNet pipelines can synthesise them in-kernel
from the per-level cell anchors, the grid width and the stride (`GridAnchors`), which removes the
[HWA, 4] tables and the tile over images of `predict_proposals` (rpn_outputs.py:418-420).  `grid_anchors` still
materialises the tables for callers that want them (same values: fp32 `shift + cell_anchor`).
"""
import math

import torch

from ..structures import BoxList


class GridAnchors(object):
    """Descriptor of one level's anchors for in-kernel synthesis: index = (y * grid_w + x) * A + a."""

    def __init__(self, cell_anchors, grid_hw, stride):
        self.cell_anchors = cell_anchors  # [A, 4] fp32 yxyx
        self.grid_hw = (int(grid_hw[0]), int(grid_hw[1]))
        self.stride = int(stride)

    @property
    def num_anchors(self):
        return self.grid_hw[0] * self.grid_hw[1] * self.cell_anchors.shape[0]

    def materialize(self):
        H, W = self.grid_hw
        dev = self.cell_anchors.device
        sy = (torch.arange(0, H * self.stride, self.stride, device=dev)).to(torch.float32)
        sx = (torch.arange(0, W * self.stride, self.stride, device=dev)).to(torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack([yy.reshape(-1), xx.reshape(-1), yy.reshape(-1), xx.reshape(-1)], 1)
        return (shifts[:, None, :] + self.cell_anchors[None, :, :]).reshape(-1, 4)


class DefaultAnchorGenerator(object):
    """For a set of feature maps, computes a set of anchors (anchor_generator.py:44-162)."""

    def __init__(self, sizes, aspect_ratios, strides, device=None):
        self.strides = list(strides)
        self.num_features = len(self.strides)
        sizes = [list(s) for s in sizes]
        aspect_ratios = [list(a) for a in aspect_ratios]
        if len(sizes) == 1:
            sizes = sizes * self.num_features
        if len(aspect_ratios) == 1:
            aspect_ratios = aspect_ratios * self.num_features
        assert self.num_features == len(sizes)
        assert self.num_features == len(aspect_ratios)
        self.sizes, self.aspect_ratios = sizes, aspect_ratios
        self.device = device
        self.cell_anchors = [self.generate_cell_anchors(s, a) for s, a in zip(sizes, aspect_ratios)]

    @property
    def box_dim(self):
        return 4

    @property
    def num_cell_anchors(self):
        return [c.shape[0] for c in self.cell_anchors]

    def generate_cell_anchors(self, sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1, 2)):
        """(len(sizes) * len(aspect_ratios), 4) anchors centred on a cell, (y0, x0, y1, x1); :111-144."""
        anchors = []
        for size in sizes:
            area = size ** 2.0
            for aspect_ratio in aspect_ratios:
                w = math.sqrt(area / aspect_ratio)
                h = aspect_ratio * w
                anchors.append([-h / 2.0, -w / 2.0, h / 2.0, w / 2.0])
        t = torch.tensor(anchors, dtype=torch.float32)
        return t.to(self.device) if self.device is not None else t

    def grid_descriptors(self, grid_sizes):
        return [GridAnchors(c, hw, s) for c, hw, s in zip(self.cell_anchors, grid_sizes, self.strides)]

    def grid_anchors(self, grid_sizes):
        return [d.materialize() for d in self.grid_descriptors(grid_sizes)]

    def __call__(self, features, materialize=True):
        """features: list of NHWC maps.  Returns BoxLists of materialised anchors (reference behaviour) or, with
        materialize=False, `GridAnchors` descriptors for the in-kernel path."""
        grid_sizes = [tuple(f.shape[1:3]) for f in features]
        if not materialize:
            return self.grid_descriptors(grid_sizes)
        return [BoxList(a) for a in self.grid_anchors(grid_sizes)]
