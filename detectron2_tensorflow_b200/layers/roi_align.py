"""Drop-in for lib/layers/roi_align.py (`ROIAlign`, :9-75)."""
from .base import Layer
from .functional import _roi_align_call, _roi_align_backward_call


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

    def backward(self, grad_output, input_shape, boxes, box_inds, grad_input=None):
        """Gradient of `call` w.r.t. `inputs` (what TF autodiff runs in training; boxes get none,
        functional.py:120).  grad_output [B, oh, ow, C] -> [N, H, W, C]; accumulates into `grad_input` if given."""
        return _roi_align_backward_call(grad_output, [input_shape], [self.spatial_scale], boxes, box_inds,
                                        self.sampling_ratio, self.aligned,
                                        grad_features=None if grad_input is None else [grad_input])[0]

    def __repr__(self):
        tmpstr = self.__class__.__name__ + "("
        tmpstr += "output_size=" + str(self.output_size)
        tmpstr += ", spatial_scale=" + str(self.spatial_scale)
        tmpstr += ", sampling_ratio=" + str(self.sampling_ratio)
        tmpstr += ", aligned=" + str(self.aligned)
        tmpstr += ")"
        return tmpstr
