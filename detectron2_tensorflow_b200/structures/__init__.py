from .box_list import BoxList, SparseBoxList
from .image_list import ImageList
