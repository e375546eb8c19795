"""Drop-in for lib/layers/roi_align.py (`ROIAlign`, :9-75)."""
from .base import Layer
from .functional import _roi_align_call


class ROIAlign(Layer):
    def __init__(self,
                 output_size,
                 spatial_scale,
                 sampling_ratio,
                 aligned=True):
        """
        Args:
            output_size (tuple): h, w
            spatial_scale (float): scale the input boxes by this number
            sampling_ratio (int): number of inputs samples to take for each
                output sample (0: one bilinear sample per bin, roi_align.py:52-66).
            aligned (bool): half-pixel-aligned sampling (functional.py:138-152).
        """
        super(ROIAlign, self).__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        assert isinstance(sampling_ratio, int), sampling_ratio
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def call(self, inputs, boxes, box_inds):
        """
        Args:
            inputs: NHWC images
            boxes: Bx4 boxes.
            box_inds: B image indices
        """
        return _roi_align_call([inputs], [self.spatial_scale], boxes, box_inds, 1, self.output_size,
                               self.sampling_ratio, self.aligned, True)

    def __repr__(self):
        tmpstr = self.__class__.__name__ + "("
        tmpstr += "output_size=" + str(self.output_size)
        tmpstr += ", spatial_scale=" + str(self.spatial_scale)
        tmpstr += ", sampling_ratio=" + str(self.sampling_ratio)
        tmpstr += ", aligned=" + str(self.aligned)
        tmpstr += ")"
        return tmpstr
