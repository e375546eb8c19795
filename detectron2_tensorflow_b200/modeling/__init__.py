from .poolers import ROIPooler, assign_boxes_to_levels
