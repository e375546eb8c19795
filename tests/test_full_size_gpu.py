"""GPU: BASELINE.json's full sizes.  Oracle comparison where the oracle finishes in seconds, otherwise
size-independent properties (idempotence, prefix/cap consistency, permutation of inputs)."""
import numpy as np
import pytest
import torch

from detectron2_tensorflow_b200.layers import batch_nms, matrix_nms, segmented_top_k
from detectron2_tensorflow_b200.modeling import RetinaNetInference, ROIPooler
from detectron2_tensorflow_b200.structures import BoxList, SparseBoxList
from detectron2_tensorflow_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def test_retinanet_config3_one_image(cuda, oracle_lib):
    """Config 3 shapes (P3..P7, A=9, K=80: 16,128,000 class scores per image), one image vs the oracle."""
    anchors = syn.retinanet_anchors()
    cls, dl = syn.retinanet_inputs(1, 80, seed=6, anchors=anchors)
    wb, ws, wc, wv, wn = oracle_lib.retinanet_inference(cls, dl, anchors, 80, 1000, 0.05, 0.5, 100)
    res = RetinaNetInference(80).inference([T(x, cuda) for x in cls], [T(x, cuda) for x in dl], [T(a, cuda) for a in anchors])
    assert np.array_equal(res.get_field('is_valid').cpu().numpy(), wv)
    assert np.array_equal(res.get_field('pred_classes').cpu().numpy(), wc)
    assert np.array_equal(res.get_field('scores').cpu().numpy(), ws)
    assert np.array_equal(res.boxes.cpu().numpy(), wb)


def test_topk_long_rows_saturated_sigmoid(cuda, oracle_lib):
    """12.1 M-element row (config 3, P3): sigmoid saturation creates ties at the k-th value."""
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((1, 151200 * 80), dtype=np.float32) * 8).astype(np.float32)
    vals, idx, cnt = segmented_top_k([T(x, cuda)], 1000, sigmoid=True)
    p = 1.0 / (1.0 + np.exp(-x[0].astype(np.float64)))
    got = idx[0, 0].cpu().numpy()
    assert cnt[0, 0].item() == 1000 and len(set(got.tolist())) == 1000
    v = vals[0, 0].cpu().numpy()
    assert np.all(np.diff(v) <= 0)
    # every non-selected element is <= the k-th value; ties at the boundary resolved to lower indices
    kth = v[-1]
    sel = np.zeros(x.shape[1], bool)
    sel[got] = True
    pv = oracle_lib.sigmoidf(x[0, got])
    assert np.array_equal(pv, v)
    assert p[~sel].max() <= float(kth) + 1e-6
    tied = np.flatnonzero(np.abs(p - float(kth)) < 1e-12)
    if len(tied) > 1:
        taken = tied[sel[tied]]
        assert taken.max() <= tied[~sel[tied]].min() if (~sel[tied]).any() else True


def test_matrix_nms_config4_resolution(cuda, oracle_lib):
    """200x336 masks (config 4 resolution), 160 candidates vs the oracle; then n=500 x batch 2 properties."""
    m, c, s = syn.solo_masks(160, hw=(200, 336), seed=7)
    want = oracle_lib.matrix_nms(m, c, s, None, "gaussian", 2.0)
    got = matrix_nms(T(m, cuda), T(c, cuda), T(s, cuda)).cpu().numpy()
    assert np.array_equal(got, want)
    m5, c5, s5 = syn.solo_masks(500, hw=(200, 336), seed=11)
    g5 = matrix_nms(T(np.stack([m5, m5]), cuda), T(np.stack([c5, c5]), cuda), T(np.stack([s5, s5]), cuda)).cpu().numpy()
    assert np.array_equal(g5[0], g5[1])
    assert g5[0, 0] == s5[0]                      # first (highest) candidate is never decayed
    assert np.all(g5[0] <= s5 + 1e-7) and np.all(g5[0] >= 0)
    # a candidate whose class appears nowhere earlier keeps its score
    first_of_class = np.array([i for i in range(500) if c5[i] not in c5[:i]])
    assert np.array_equal(g5[0][first_of_class], s5[first_of_class])


@pytest.mark.parametrize("n", [16384, 65536])
def test_nms_sweep_sizes(cuda, oracle_lib, n):
    rng = np.random.default_rng(n)
    cy, cx = rng.uniform(0, 800, n), rng.uniform(0, 1333, n)
    h, w = rng.uniform(16, 300, n), rng.uniform(16, 300, n)
    b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1).astype(np.float32)
    sc = rng.standard_normal(n).astype(np.float32)
    keep, num = batch_nms(T(b[None], cuda), T(sc[None], cuda), n, axis=1, iou_threshold=0.7)
    k = keep[0, :int(num[0])].cpu().numpy()
    want = oracle_lib.nms(b, sc, n, 0.7)   # O(n * kept) on the CPU: seconds
    assert np.array_equal(k, want)
    # capped run is a prefix of the uncapped one
    keep2, num2 = batch_nms(T(b[None], cuda), T(sc[None], cuda), 100, axis=1, iou_threshold=0.7)
    assert int(num2[0]) == 100 and np.array_equal(keep2[0].cpu().numpy(), want[:100])


def test_roi_align_config2_properties(cuda, oracle_lib):
    """16,000 ROIs over the batch-16 pyramid: oracle on a sampled subset + ROI-order permutation invariance."""
    N, R, C = 16, 1000, 256
    g = torch.Generator(device=cuda).manual_seed(0)
    feats = [torch.randn((N,) + syn.level_hw(s) + (C,), device=cuda, generator=g) for s in syn.FPN_STRIDES]
    boxes, idx = syn.rois(N, R, seed=1)
    scales = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    pooler = ROIPooler(7, scales, 0, "ROIAlignV2")
    out = pooler(feats, SparseBoxList(T(idx, cuda), BoxList(T(boxes, cuda)), (N, R)))
    perm = np.random.default_rng(0).permutation(N * R)
    out_p = pooler(feats, SparseBoxList(T(idx[perm], cuda), BoxList(T(boxes[perm], cuda)), (N, R)))
    assert torch.equal(out_p, out[torch.from_numpy(perm).to(cuda)])
    # oracle on the ROIs of image 3 only (its own feature maps)
    sel = np.flatnonzero(idx[:, 0] == 3)[:200]
    f3 = [f[3:4].cpu().numpy() for f in feats]
    want, _ = oracle_lib.roi_pooler(f3, scales, boxes[sel], np.zeros(len(sel), np.int64), (7, 7), 0)
    assert np.array_equal(out[torch.from_numpy(sel).to(cuda)].cpu().numpy(), want)
