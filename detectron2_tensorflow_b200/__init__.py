"""detectron2_tensorflow_b200 -- B200-native (sm_100a) drop-in for the detection
post-backbone hot path of SimeonZhang/detectron2_tensorflow.

Host layer mirrors the reference's operator signatures (lib/layers, lib/modeling/poolers,
rpn_outputs.find_top_rpn_proposals, fast_rcnn_inference, RetinaNetHead.inference, matrix_nms)
over the C-ABI of libd2b200.so (include/d2b200.h).  No CPU fallback.
"""
from . import _native  # noqa: F401

__version__ = "0.1.0"
