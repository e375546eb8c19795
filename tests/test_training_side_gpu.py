"""GPU parity for SURVEY.md 8(f) #3 (training-side neighbours) and the 8(b) single-stage entry points:
pairwise_iou, Matcher, fused label assignment (+ get_deltas), ROIAlign / ROIPooler backward,
crop_and_resize entry, decode+clip+filter.

Bar: integer outputs (matches, labels, keep flags) and elementwise fp32 outputs bit-exact vs the oracle;
the backward pass sums fp32 contributions in a different order than the TF CPU kernel (atomics), so it
is compared within 1e-5 relative to the per-pixel sum of |contributions| (north_star tolerance 1e-5).
"""
import numpy as np
import pytest
import torch

from detectron2_tensorflow_b200.layers import ROIAlign, crop_and_resize
from detectron2_tensorflow_b200.modeling import Box2BoxTransform, Matcher, ROIPooler, label_boxes
from detectron2_tensorflow_b200.structures import BoxList, SparseBoxList, pairwise_iou
from detectron2_tensorflow_b200.utils import synthetic as syn
from detectron2_tensorflow_b200 import _native as nv

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def rand_boxes(rng, n, H=800, W=1333, smin=8, smax=400):
    cy, cx = rng.uniform(0, H, n), rng.uniform(0, W, n)
    h, w = rng.uniform(smin, smax, n), rng.uniform(smin, smax, n)
    return np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1).astype(np.float32)


def make_gt(rng, N, G, anchors=None):
    """GT boxes with flags; some copied from anchors (exact IoU 1 and ties), some degenerate."""
    gt = np.stack([rand_boxes(rng, G, smin=16, smax=500) for _ in range(N)])
    if anchors is not None:
        for n in range(N):
            pick = rng.integers(0, anchors.shape[0], 3)
            gt[n, :3] = anchors[pick]
            gt[n, 3] = gt[n, 2]  # duplicate GT: argmax tie -> first wins
    gt[:, -1] = 0.0  # zero-area GT (union == area of the other box)
    valid = rng.random((N, G)) < 0.8
    crowd = rng.random((N, G)) < 0.15
    difficult = rng.random((N, G)) < 0.1
    valid[0, :] = False  # image without any valid GT
    return gt.astype(np.float32), valid, crowd, difficult


# ------------------------------------------------------------------ pairwise_iou / get_deltas
def test_pairwise_iou(cuda, oracle_lib):
    rng = np.random.default_rng(0)
    b1, b2 = rand_boxes(rng, 37), rand_boxes(rng, 5001)
    b2[:37] = b1  # IoU exactly 1 on the diagonal
    b1[5] = 0.0
    b2[100] = b2[100][[2, 3, 0, 1]]  # inverted box: negative sides
    want = oracle_lib.pairwise_iou(b1, b2)
    got = pairwise_iou(BoxList(T(b1, cuda)), BoxList(T(b2, cuda))).cpu().numpy()
    assert np.array_equal(got, want)
    assert got[0, 0] == 1.0
    assert pairwise_iou(T(b1[:0], cuda), T(b2, cuda)).shape == (0, 5001)


def test_get_deltas_roundtrip(cuda, oracle_lib):
    rng = np.random.default_rng(1)
    src, tgt = rand_boxes(rng, 4000), rand_boxes(rng, 4000)
    for w in ((1., 1., 1., 1.), (10., 10., 5., 5.)):
        t = Box2BoxTransform(w)
        want = oracle_lib.get_deltas(src, tgt, w)
        got = t.get_deltas(T(src, cuda), T(tgt, cuda))
        assert np.array_equal(got.cpu().numpy(), want)
        # size-independent property: apply_deltas(get_deltas(src, tgt), src) == tgt up to fp32 rounding
        back = t.apply_deltas(got, T(src, cuda)).cpu().numpy()
        assert np.abs(back - tgt).max() < 1e-2


# ------------------------------------------------------------------ Matcher on matrices
@pytest.mark.parametrize("allow_lq", [False, True])
def test_matcher_matrix(cuda, oracle_lib, allow_lq):
    rng = np.random.default_rng(2)
    M, N = 23, 7001
    q = rng.random((M, N)).astype(np.float32)
    q[:, ::5] = np.round(q[:, ::5] * 8) / 8  # ties along both axes
    q[3] = 0.0  # a GT with no overlap at all: every prediction ties at its maximum 0
    crowd = (rng.random((4, N)) * (rng.random((4, N)) < 0.01)).astype(np.float32)
    diff = (rng.random((3, N)) * (rng.random((3, N)) < 0.02)).astype(np.float32)
    for th, lab in (([0.3, 0.7], [0, -1, 1]), ([0.5], [0, 1])):
        m = Matcher(th, lab, allow_low_quality_matches=allow_lq)
        for cm, dm in ((None, None), (crowd, None), (crowd, diff), (crowd[:0], diff[:0])):
            want = oracle_lib.matcher(q, th, lab, allow_lq, cm, dm)
            got = m(T(q, cuda), None if cm is None else T(cm, cuda), None if dm is None else T(dm, cuda))
            assert np.array_equal(got[0].cpu().numpy(), want[0])
            assert np.array_equal(got[1].cpu().numpy(), want[1])
    # M == 0: matches 0, labels 0 (matcher.py:117-122)
    got = Matcher([0.5], [0, 1], allow_lq)(T(q[:0], cuda))
    assert int(got[0].abs().sum()) == 0 and int(got[1].abs().sum()) == 0


# ------------------------------------------------------------------ fused label assignment
@pytest.mark.parametrize("allow_lq,boundary", [(True, -1), (True, 0), (False, -1)])
def test_label_boxes_rpn(cuda, oracle_lib, allow_lq, boundary):
    """RPNOutputs._get_ground_truth on the real 268 K-anchor pyramid (rpn_outputs.py:245-304)."""
    rng = np.random.default_rng(3)
    anchors = np.concatenate(syn.rpn_anchors(), 0)
    N, G = 3, 40
    gt, valid, crowd, _ = make_gt(rng, N, G, anchors)
    shapes = syn.image_shapes(N)
    th, lab, w = [0.3, 0.7], [0, -1, 1], (1.0, 1.0, 1.0, 1.0)
    wm, wl, wd = oracle_lib.label_boxes(anchors, gt, valid, th, lab, allow_lq, gt_crowd=crowd,
                                        boundary_threshold=float(boundary), image_shapes=shapes, weights=w)
    m = Matcher(th, lab, allow_low_quality_matches=allow_lq)
    gm, gl, gd = label_boxes(T(anchors, cuda), T(gt, cuda), T(valid, cuda), m, gt_crowd=T(crowd, cuda),
                             boundary_threshold=boundary, image_shapes=T(shapes, cuda),
                             box2box_transform=Box2BoxTransform(w))
    assert np.array_equal(gl.cpu().numpy(), wl)
    assert np.array_equal(gm.cpu().numpy(), wm)
    assert np.array_equal(gd.cpu().numpy(), wd, equal_nan=True)
    assert (wl == 1).sum() > 0 and (wl == -1).sum() > 0 and (wl == 0).sum() > 0


def test_label_boxes_roi_heads(cuda, oracle_lib):
    """ROIHeads.label_and_sample_proposals matching part: per-image proposals with a valid prefix,
    crowd + difficult lists (roi_heads.py:100-165)."""
    rng = np.random.default_rng(4)
    N, P, G = 4, 2000, 64
    pred = np.stack([rand_boxes(rng, P) for _ in range(N)])
    gt, valid, crowd, difficult = make_gt(rng, N, G)
    pred[:, :G] = gt  # proposal_append_gt (roi_heads.py:122-123): IoU 1 with their GT
    counts = np.array([2000, 1500, 0, 1], np.int32)
    th, lab = [0.5], [0, 1]
    wm, wl, _ = oracle_lib.label_boxes(pred, gt, valid, th, lab, False, gt_crowd=crowd, gt_difficult=difficult,
                                       pred_counts=counts)
    gm, gl, gd = label_boxes(T(pred, cuda), T(gt, cuda), T(valid, cuda), Matcher(th, lab), gt_crowd=T(crowd, cuda),
                             gt_difficult=T(difficult, cuda), pred_counts=T(counts, cuda))
    assert gd is None
    assert np.array_equal(gl.cpu().numpy(), wl)
    assert np.array_equal(gm.cpu().numpy(), wm)


def test_label_boxes_matches_unfused_chain(cuda):
    """The fused kernel equals pairwise_iou -> Matcher on the device (the reference's own composition)."""
    rng = np.random.default_rng(5)
    anchors = np.concatenate(syn.rpn_anchors(), 0)[::7]
    gt = rand_boxes(rng, 30, smin=16, smax=500)
    gt[:4] = anchors[[5, 50, 500, 5000]]
    m = Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
    q = pairwise_iou(T(gt, cuda), T(anchors, cuda))
    um, ul = m(q)
    fm, fl, _ = label_boxes(T(anchors, cuda), T(gt[None], cuda), torch.ones((1, 30), dtype=torch.bool, device=cuda), m)
    assert torch.equal(um, fm[0]) and torch.equal(ul, fl[0])


# ------------------------------------------------------------------ 8(b) single-stage entries
def test_decode_clip_filter(cuda, oracle_lib):
    rng = np.random.default_rng(6)
    N, n = 3, 5000
    anchors = rand_boxes(rng, n)
    deltas = (rng.standard_normal((N, n, 4)) * 0.5).astype(np.float32)
    shapes = np.array([[800, 1333], [600, 900], [333, 500]], np.int32)
    boxes = torch.empty((N, n, 4), dtype=torch.float32, device=cuda)
    keep = torch.empty((N, n), dtype=torch.uint8, device=cuda)
    p = nv.DecodeClipFilterParams()
    d_t, a_t, s_t = T(deltas, cuda), T(anchors, cuda), T(shapes, cuda)
    p.deltas, p.anchors, p.num_images, p.n, p.image_shapes = d_t.data_ptr(), a_t.data_ptr(), N, n, s_t.data_ptr()
    for i in range(4):
        p.weights[i] = 1.0
    p.scale_clamp, p.min_box_side_len = float(oracle_lib.SCALE_CLAMP), 12.0
    p.out_boxes, p.out_keep = boxes.data_ptr(), keep.data_ptr()
    nv.call("decode_clip_filter", p, cuda)
    want = oracle_lib.rpn_predict_proposals(deltas, anchors)
    for i in range(N):
        h, w = float(shapes[i, 0]), float(shapes[i, 1])
        wb = want[i].copy()
        wb[:, 0::2] = np.maximum(np.minimum(wb[:, 0::2], h), 0.0)
        wb[:, 1::2] = np.maximum(np.minimum(wb[:, 1::2], w), 0.0)
        assert np.array_equal(boxes[i].cpu().numpy(), wb)
        wk = ((wb[:, 3] - wb[:, 1]) >= 12.0) & ((wb[:, 2] - wb[:, 0]) >= 12.0)
        assert np.array_equal(keep[i].cpu().numpy().astype(bool), wk)


@pytest.mark.parametrize("aligned,pad", [(True, True), (False, True), (True, False)])
def test_crop_and_resize_entry(cuda, oracle_lib, aligned, pad):
    rng = np.random.default_rng(7)
    img = rng.standard_normal((2, 30, 40, 8)).astype(np.float32)
    boxes = rand_boxes(rng, 50, 30, 40, 2, 25)
    bi = rng.integers(0, 2, 50).astype(np.int32)
    bi[7] = 5  # out of range -> zero row
    want = oracle_lib.crop_and_resize(img, boxes, bi, (14, 14), aligned, pad)
    got = crop_and_resize(T(img, cuda), T(boxes, cuda), T(bi, cuda), (14, 14), aligned=aligned, pad_border=pad)
    assert np.array_equal(got.cpu().numpy(), want)


# ------------------------------------------------------------------ ROIAlign backward
def _tol_check(got, want, bound):
    err = np.abs(got - want)
    assert (err <= 1e-5 * bound + 1e-7).all(), float((err / (bound + 1e-12)).max())


@pytest.mark.parametrize("sr,aligned", [(0, True), (2, True), (0, False)])
def test_roi_align_backward_single(cuda, oracle_lib, sr, aligned):
    rng = np.random.default_rng(8)
    N, H, W, Cc, M = 2, 25, 42, 16, 300
    boxes = rand_boxes(rng, M, H * 16, W * 16, 8, 300)
    boxes[:5] += 500.0  # partly / fully outside
    bi = rng.integers(0, N, M).astype(np.int32)
    bi[9] = -1
    g = rng.standard_normal((M, 7, 7, Cc)).astype(np.float32)
    want = oracle_lib.roi_align_backward(g, (N, H, W, Cc), boxes, bi, 1 / 16., sr, aligned)
    bound = oracle_lib.roi_align_backward(np.abs(g), (N, H, W, Cc), boxes, bi, 1 / 16., sr, aligned)
    layer = ROIAlign((7, 7), 1 / 16., sr, aligned)
    got = layer.backward(T(g, cuda), (N, H, W, Cc), T(boxes, cuda), T(bi, cuda)).cpu().numpy()
    _tol_check(got, want, bound)
    # linearity / accumulation: a second call into the same buffer doubles the gradient
    buf = T(got, cuda).clone()
    layer.backward(T(g, cuda), (N, H, W, Cc), T(boxes, cuda), T(bi, cuda), grad_input=buf)
    _tol_check(buf.cpu().numpy(), 2 * want, 2 * bound)


def test_roi_pooler_backward_adjoint(cuda, oracle_lib):
    """Multi-level backward vs the oracle, and the adjoint identity <pool(x), g> == <x, pool^T(g)>
    against the (bit-exact) forward kernel."""
    rng = np.random.default_rng(9)
    N, Cc, R = 2, 32, 200
    feats = syn.fpn_features(N, Cc, seed=11)
    shapes = [f.shape for f in feats]
    boxes, idx = syn.rois(N, R, seed=12)
    g = rng.standard_normal((N * R, 7, 7, Cc)).astype(np.float32)
    scales = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    want = oracle_lib.roi_pooler_backward(g, shapes, scales, boxes, idx[:, 0], 0)
    bound = oracle_lib.roi_pooler_backward(np.abs(g), shapes, scales, boxes, idx[:, 0], 0)
    pooler = ROIPooler((7, 7), scales, 0, "ROIAlignV2")
    inst = SparseBoxList(T(idx, cuda), BoxList(T(boxes, cuda)), (N, R))
    got = pooler.backward(T(g, cuda), shapes, inst)
    for a, b, c in zip(got, want, bound):
        _tol_check(a.cpu().numpy(), b, c)
    out = pooler([T(f, cuda) for f in feats], inst)
    lhs = float((out.double() * T(g, cuda).double()).sum())
    rhs = float(sum((T(f, cuda).double() * gf.double()).sum() for f, gf in zip(feats, got)))
    assert abs(lhs - rhs) <= 1e-5 * float(sum((np.abs(f).astype(np.float64) * c).sum() for f, c in zip(feats, bound)))


@pytest.mark.parametrize("iou_type", ["iou", "giou", "diou", "ciou"])
def test_pairwise_iou_variants(cuda, oracle_lib, iou_type):
    """pairwise_iou(iou_type=...) (box_list_ops.py:295-371) vs the oracle (exact; CIoU 1e-5 through atan) and vs the
    reference-python golden."""
    import os
    from detectron2_tensorflow_b200.structures import pairwise_iou
    rng = np.random.default_rng(77)
    a = np.stack([rng.uniform(0, 200, 300), rng.uniform(0, 300, 300), rng.uniform(0, 200, 300), rng.uniform(0, 300, 300)], 1)
    a = np.concatenate([np.minimum(a[:, :2], a[:, 2:]), np.maximum(a[:, :2], a[:, 2:])], 1).astype(np.float32)
    b = a[rng.permutation(300)[:130]] + rng.normal(0, 3, (130, 4)).astype(np.float32)
    b[:10] = a[:10]
    b[10] = [5, 5, 5, 50]
    got = pairwise_iou(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda), iou_type).cpu().numpy()
    want = oracle_lib.pairwise_iou(a, b, iou_type)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_python.npz"))
    gz = pairwise_iou(torch.from_numpy(z["pi_a"]).to(cuda), torch.from_numpy(z["pi_b"]).to(cuda), iou_type).cpu().numpy()
    if iou_type == "ciou":
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.allclose(got, want, rtol=1e-5, atol=1e-6, equal_nan=True)
        assert np.allclose(gz, z["pi_ciou"], rtol=1e-5, atol=1e-6, equal_nan=True)
    else:
        assert np.array_equal(got, want)
        assert np.array_equal(gz, z[f"pi_{iou_type}"])
    with pytest.raises(ValueError):
        pairwise_iou(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda), "siou")


# ------------------------------------------------------------------ subsample_labels (sampling.py:6-45)
@pytest.mark.parametrize("P,num_samples,frac,bg,npos,nneg", [
    (268569, 256, 0.5, 0, 40, 200000),     # RPN: few positives, the sample is filled with negatives
    (2000, 512, 0.25, 80, 300, 1500),      # ROI heads
    (2000, 512, 0.25, 80, 20, 100),        # not enough of either class
    (1000, 10, 0.7, 0, 500, 400),          # int(10 * 0.7) == 7 in double arithmetic
    (50, 64, 0.5, 0, 0, 50), (7, 4, 0.5, 0, 7, 0), (0, 8, 0.5, 0, 0, 0)])
def test_subsample_labels(cuda, oracle_lib, P, num_samples, frac, bg, npos, nneg):
    from detectron2_tensorflow_b200.modeling import subsample_labels, subsample_labels_batched
    rng = np.random.default_rng(P + num_samples)
    N = 3
    labels = np.full((N, P), -1, np.int64)
    for n in range(N):
        perm = rng.permutation(P)
        labels[n, perm[:npos]] = rng.integers(1, 80, npos) if bg == 0 else rng.integers(0, 80, npos)
        labels[n, perm[npos:npos + nneg]] = bg
    pos, neg, cp, cn, lab = subsample_labels_batched(T(labels, cuda), num_samples, frac, bg, seed=1234, return_labels=True)
    pos, neg, cp, cn, lab = (t.cpu().numpy() for t in (pos, neg, cp, cn, lab))
    want_pos = min(npos, int(num_samples * frac))
    want_neg = min(nneg, num_samples - want_pos)
    for n in range(N):
        # the reference's contract (properties): exact counts, only eligible indices, no duplicates, -1 padding
        assert cp[n] == want_pos and cn[n] == want_neg
        ip, ineg = pos[n, :cp[n]], neg[n, :cn[n]]
        assert np.all(pos[n, cp[n]:] == -1) and np.all(neg[n, cn[n]:] == -1)
        assert len(set(ip.tolist())) == len(ip) and len(set(ineg.tolist())) == len(ineg)
        assert np.all((labels[n, ip] != -1) & (labels[n, ip] != bg)) and np.all(labels[n, ineg] == bg)
        # resampled labels: sampled elements keep their label, everything else is ignore (rpn_outputs.py:315-329)
        want_lab = np.full(P, -1, np.int64)
        want_lab[ip] = labels[n, ip]
        want_lab[ineg] = labels[n, ineg]
        assert np.array_equal(lab[n], want_lab)
        # the documented generator, restated in numpy: identical indices in identical order
        op, on = oracle_lib.subsample_labels(labels[n], num_samples, frac, bg, seed=1234, image=n)
        assert np.array_equal(ip, op) and np.array_equal(ineg, on)
    if P:
        # reference signature (1-D labels -> dynamic-length index tensors), determinism per seed, seeds differ
        a1, b1 = subsample_labels(T(labels[0], cuda), num_samples, frac, bg, seed=5)
        a2, b2 = subsample_labels(T(labels[0], cuda), num_samples, frac, bg, seed=5)
        a3, b3 = subsample_labels(T(labels[0], cuda), num_samples, frac, bg, seed=6)
        assert torch.equal(a1, a2) and torch.equal(b1, b2)
        assert a1.shape[0] == want_pos and b1.shape[0] == want_neg
        if nneg > 4 * num_samples:
            assert not torch.equal(b1, b3)


def test_subsample_labels_is_uniform(cuda):
    """Every negative is equally likely: 2000 seeds x 8-of-64 sampling, chi-square well inside its 99.9 % bound."""
    from detectron2_tensorflow_b200.modeling import subsample_labels_batched
    labels = torch.zeros((1, 64), dtype=torch.int64, device=cuda)
    hits = np.zeros(64, np.int64)
    for seed in range(2000):
        _, neg, _, cn = subsample_labels_batched(labels, 8, 0.5, 0, seed=seed)
        assert int(cn[0]) == 8
        hits[neg[0].cpu().numpy()] += 1
    expected = 2000 * 8 / 64.0
    chi2 = float(((hits - expected) ** 2 / expected).sum())
    assert chi2 < 110.0  # 63 degrees of freedom: P(chi2 > 103.4) = 0.001
