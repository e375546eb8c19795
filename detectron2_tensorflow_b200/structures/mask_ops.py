"""Drop-in for lib/structures/mask_ops.py: `reframe_box_masks_to_image_masks` (:7-56)."""
import torch

from .. import _native as nv


def reframe_box_masks_to_image_masks(box_masks, boxes, image_shape, mask_threshold=0.5, scope=None):
    """Transforms the box masks back to full image masks.

    Args:
        box_masks: [M, mh, mw] fp32 mask probabilities.
        boxes: [M, 4] absolute boxes (ymin, xmin, ymax, xmax) in the output image.
        image_shape: (height, width) of the output masks.
    Returns:
        uint8 [M, height, width] (1 where the pasted probability exceeds `mask_threshold`).
    """
    host = not box_masks.is_cuda
    dev = nv.device_of(box_masks, boxes)
    m = nv.to_device(box_masks, dev, torch.float32)
    b = nv.to_device(boxes, dev, torch.float32).reshape(-1, 4)
    assert m.dim() == 3 and m.shape[0] == b.shape[0]
    H, W = int(image_shape[0]), int(image_shape[1])
    out = torch.empty((m.shape[0], H, W), dtype=torch.uint8, device=dev)
    p = nv.PasteMasksParams()
    p.box_masks, p.boxes, p.num_masks = m.data_ptr(), b.data_ptr(), m.shape[0]
    p.mask_h, p.mask_w, p.image_h, p.image_w = m.shape[1], m.shape[2], H, W
    p.mask_threshold = float(mask_threshold)
    p.out = out.data_ptr()
    nv.call("paste_masks", p, dev)
    return nv.to_host(out) if host else out
