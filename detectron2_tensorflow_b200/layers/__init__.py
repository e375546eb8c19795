from .base import Layer
from .roi_align import ROIAlign
from .functional import crop_and_resize
from .nms import batch_nms, matrix_nms, non_max_suppression
from .topk import segmented_top_k
