"""Drop-in for `detector_postprocess` of lib/modeling/postprocessing.py (:9-59): the last step of
`GeneralizedRCNN.inference` (rcnn.py:124-133) -- paste the ROI-head masks into image-size masks."""
import torch

from ..structures.box_list import SparseBoxList
from ..structures.mask_ops import reframe_box_masks_to_image_masks


def detector_postprocess(results, output_shape, mask_format, image_shapes=None, mask_threshold=0.5, scope=None):
    """Image-size masks for the detections of an R-CNN (arguments and return structure of the reference function).

    results: dense BoxList [N, R] with `pred_masks` [N, R, mh, mw], `is_valid` and the `image_shape` tracking.
    mask_format "conventional": masks are pasted at the boxes as they are; "fixed": at the boxes scaled by
    output_shape / image_shape (:36-43 -- only the pasting boxes are scaled, `results.boxes` stay, as in the reference;
    the int / int division there is TF's true division, i.e. float64, cast to fp32 by `box_list_ops.scale`);
    "raw": `pred_masks > mask_threshold` (the reference's branch reads an undefined name, :51; this is its intent).
    Returns the dense BoxList with `pred_masks` uint8 [N, R, H, W]."""
    if not results.has_field("pred_masks"):
        return results
    if mask_format in ("conventional", "fixed"):
        results = SparseBoxList.from_dense(results)
        box_masks = results.data.get_field("pred_masks")
        if mask_format == "fixed":
            assert image_shapes is not None, "Detection results should carry the true input shape."
            image_shapes = results.get_tracking('image_shape')
            out = torch.as_tensor([float(output_shape[0]), float(output_shape[1])], dtype=torch.float64,
                                  device=image_shapes.device)
            scales = (out[None] / image_shapes.to(torch.float64))[results.indices[:, 0]].to(torch.float32)
            b = results.data.boxes
            boxes = torch.stack([scales[:, 0] * b[:, 0], scales[:, 1] * b[:, 1], scales[:, 0] * b[:, 2],
                                 scales[:, 1] * b[:, 3]], 1)
        else:
            boxes = results.data.boxes
        pred_masks = reframe_box_masks_to_image_masks(box_masks, boxes, output_shape, mask_threshold)
        results.data.set_field("pred_masks", pred_masks)
        results = results.to_dense()
    elif mask_format == "raw":
        pred_masks = results.get_field("pred_masks")
        results.set_field("pred_masks", (pred_masks > mask_threshold).to(torch.uint8))
    else:
        raise ValueError(f"mask format '{mask_format}' is not recognized.")
    return results
