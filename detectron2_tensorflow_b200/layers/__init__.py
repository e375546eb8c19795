from .base import Layer
from .roi_align import ROIAlign
from .functional import crop_and_resize
