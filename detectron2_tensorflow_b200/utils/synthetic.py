"""Seeded synthetic inputs of the BASELINE.json configs.

All generators return numpy fp32/int arrays created on the host with
``numpy.random.default_rng(seed)``; tests and bench.py copy them to the device.
Shapes follow one 800x1333 image padded to 800x1344.
"""
import math

import numpy as np

IMAGE_HW = (800, 1333)
PADDED_HW = (800, 1344)
FPN_STRIDES = (4, 8, 16, 32)          # P2..P5 (ROI heads)
RPN_STRIDES = (4, 8, 16, 32, 64)      # P2..P6
RETINA_STRIDES = (8, 16, 32, 64, 128)  # P3..P7
RPN_SIZES = (32, 64, 128, 256, 512)
ASPECT_RATIOS = (0.5, 1.0, 2.0)


def level_hw(stride, padded_hw=PADDED_HW):
    return (int(math.ceil(padded_hw[0] / stride)), int(math.ceil(padded_hw[1] / stride)))


def fpn_features(n_images, channels=256, strides=FPN_STRIDES, seed=0, padded_hw=PADDED_HW, dtype=np.float32):
    """P2..P5 NHWC feature maps, i.i.d. N(0,1)."""
    rng = np.random.default_rng(seed)
    out = []
    for s in strides:
        h, w = level_hw(s, padded_hw)
        out.append(rng.standard_normal((n_images, h, w, channels), dtype=np.float32).astype(dtype, copy=False))
    return out


def rois(n_images, rois_per_image, seed=1, image_hw=IMAGE_HW, frac_outside=0.02):
    """ROIs: sqrt(area) log-uniform in [16, 900] px, aspect log-uniform in [0.5, 2], centre uniform,
    clipped to the image; `frac_outside` of them then extend up to 32 px outside.
    Returns boxes [M,4] yxyx fp32 and indices [M,2] int64 (image, slot), image-major."""
    rng = np.random.default_rng(seed)
    H, W = image_hw
    M = n_images * rois_per_image
    s = np.exp(rng.uniform(np.log(16.0), np.log(900.0), M))
    a = np.exp(rng.uniform(np.log(0.5), np.log(2.0), M))  # h / w
    h = s * np.sqrt(a)
    w = s / np.sqrt(a)
    cy = rng.uniform(0, H, M)
    cx = rng.uniform(0, W, M)
    b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1)
    b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, H)
    b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, W)
    out = rng.uniform(0, 1, M) < frac_outside
    b[out] += rng.uniform(-32, 32, (int(out.sum()), 4))
    b = b.astype(np.float32)
    # keep y2>=y1, x2>=x1
    b = np.stack([np.minimum(b[:, 0], b[:, 2]), np.minimum(b[:, 1], b[:, 3]),
                  np.maximum(b[:, 0], b[:, 2]), np.maximum(b[:, 1], b[:, 3])], 1)
    img = np.repeat(np.arange(n_images, dtype=np.int64), rois_per_image)
    slot = np.tile(np.arange(rois_per_image, dtype=np.int64), n_images)
    return np.ascontiguousarray(b), np.stack([img, slot], 1)


def cell_anchors(sizes, aspect_ratios):
    """DefaultAnchorGenerator.generate_cell_anchors (lib/modeling/anchor_generator.py:111-144), yxyx."""
    out = []
    for size in sizes:
        area = float(size) ** 2.0
        for ar in aspect_ratios:
            w = math.sqrt(area / ar)
            h = ar * w
            out.append([-h / 2.0, -w / 2.0, h / 2.0, w / 2.0])
    return np.asarray(out, np.float32)


def grid_anchors(hw, stride, cell):
    """DefaultAnchorGenerator.grid_anchors (anchor_generator.py:92-109): order (y, x, a)."""
    H, W = hw
    sy = (np.arange(H, dtype=np.float32) * np.float32(stride))
    sx = (np.arange(W, dtype=np.float32) * np.float32(stride))
    yy, xx = np.meshgrid(sy, sx, indexing="ij")
    shifts = np.stack([yy.ravel(), xx.ravel(), yy.ravel(), xx.ravel()], 1)  # [HW,4]
    return (shifts[:, None, :] + cell[None, :, :]).reshape(-1, 4).astype(np.float32)


def rpn_anchors(strides=RPN_STRIDES, sizes=RPN_SIZES, padded_hw=PADDED_HW):
    return [grid_anchors(level_hw(s, padded_hw), s, cell_anchors([sz], ASPECT_RATIOS)) for s, sz in zip(strides, sizes)]


def retinanet_anchors(strides=RETINA_STRIDES, padded_hw=PADDED_HW):
    out = []
    for s in strides:
        base = s * 4
        sizes = [base * 2 ** (i / 3.0) for i in range(3)]
        out.append(grid_anchors(level_hw(s, padded_hw), s, cell_anchors(sizes, ASPECT_RATIOS)))
    return out


def rpn_inputs(n_images, seed=2, variant="gaussian", anchors=None):
    """Per-level logits [N,HWA] and deltas [N,HWA,4].
    variant: "gaussian" (logits N(0,2^2)), "clustered" (peaked near 60 latent objects per image so NMS
    suppresses most boxes, like a trained RPN), "ties" (logits rounded to 1/64)."""
    rng = np.random.default_rng(seed)
    anchors = rpn_anchors() if anchors is None else anchors
    logits, deltas = [], []
    for a in anchors:
        hwa = a.shape[0]
        lg = (rng.standard_normal((n_images, hwa)) * 2.0).astype(np.float32)
        d = rng.standard_normal((n_images, hwa, 4)).astype(np.float32)
        d[..., :2] *= 0.5
        d[..., 2:] *= 0.25
        if variant == "ties":
            lg = (np.round(lg * 64.0) / 64.0).astype(np.float32)
        logits.append(lg)
        deltas.append(d)
    if variant == "clustered":
        H, W = IMAGE_HW
        for n in range(n_images):
            objs_c = np.stack([rng.uniform(0, H, 60), rng.uniform(0, W, 60)], 1)
            objs_s = np.exp(rng.uniform(np.log(24), np.log(500), 60))
            for l, a in enumerate(anchors):
                ac = np.stack([(a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2], 1)
                asz = np.sqrt((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))
                boost = np.zeros(a.shape[0], np.float32)
                for (oy, ox), os_ in zip(objs_c, objs_s):
                    if not (0.5 < os_ / asz[0] < 2.0):
                        continue
                    d2 = ((ac[:, 0] - oy) ** 2 + (ac[:, 1] - ox) ** 2) / (0.25 * os_) ** 2
                    boost = np.maximum(boost, 9.0 * np.exp(-0.5 * d2).astype(np.float32))
                logits[l][n] += boost
                deltas[l][n] *= 0.2
    return logits, deltas


def image_shapes(n_images, image_hw=IMAGE_HW):
    return np.tile(np.asarray(image_hw, np.int32), (n_images, 1))


def fast_rcnn_inputs(n_images, rois_per_image, num_classes=80, seed=5, proposal_boxes=None):
    """scores = softmax(N(0,3^2)) over K+1; class-specific deltas N(0,1) scaled by 1/(10,10,5,5);
    boxes = apply_deltas done by the caller.  Returns (scores [M,K+1], deltas [M,K*4])."""
    rng = np.random.default_rng(seed)
    M = n_images * rois_per_image
    lg = rng.standard_normal((M, num_classes + 1)) * 3.0
    lg -= lg.max(1, keepdims=True)
    e = np.exp(lg)
    scores = (e / e.sum(1, keepdims=True)).astype(np.float32)
    d = rng.standard_normal((M, num_classes, 4)).astype(np.float32)
    return scores, d.reshape(M, num_classes * 4)


def retinanet_inputs(n_images, num_classes=80, seed=6, anchors=None):
    """box_cls [N,HWA,K] logits N(-4.6,1.5^2) (prior 0.01, retinanet.py:425), box_delta [N,HWA,4] N(0,0.3^2)."""
    rng = np.random.default_rng(seed)
    anchors = retinanet_anchors() if anchors is None else anchors
    cls, dl = [], []
    for a in anchors:
        hwa = a.shape[0]
        c = rng.standard_normal((n_images, hwa, num_classes), dtype=np.float32)
        c *= np.float32(1.5)
        c -= np.float32(4.6)
        cls.append(c)
        dl.append((rng.standard_normal((n_images, hwa, 4), dtype=np.float32) * np.float32(0.3)))
    return cls, dl


def solo_masks(n_masks=500, hw=(200, 336), num_classes=80, seed=7, dup_frac=0.4):
    """Binary ellipse masks [n,H,W] (area log-uniform 200..20000 px, `dup_frac` jittered duplicates),
    classes int64 uniform, scores sorted descending in U(0.1, 1)."""
    rng = np.random.default_rng(seed)
    H, W = hw
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    masks = np.zeros((n_masks, H, W), np.float32)
    classes = rng.integers(0, num_classes, n_masks).astype(np.int64)
    params = []
    for i in range(n_masks):
        if i > 0 and rng.uniform() < dup_frac:
            j = int(rng.integers(0, i))
            cy, cx, ry, rx = params[j]
            cy, cx = cy + rng.normal(0, 2.0), cx + rng.normal(0, 2.0)
            ry, rx = ry * np.exp(rng.normal(0, 0.08)), rx * np.exp(rng.normal(0, 0.08))
            classes[i] = classes[j]
        else:
            area = np.exp(rng.uniform(np.log(200.0), np.log(20000.0)))
            ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0)))
            ry, rx = math.sqrt(area / math.pi * ar), math.sqrt(area / math.pi / ar)
            cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        params.append((cy, cx, ry, rx))
        masks[i] = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0).astype(np.float32)
    scores = np.sort(rng.uniform(0.1, 1.0, n_masks).astype(np.float32))[::-1].copy()
    return masks, classes, scores
