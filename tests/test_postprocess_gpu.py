"""GPU parity for top-k / decode / NMS / RPN / Fast R-CNN / RetinaNet / Matrix-NMS vs the CPU oracle.

Bar (BASELINE.json north_star): kept indices and counts bit-exact; box/score values are produced by
the same fp32 op order (no FMA) so they are compared for exact equality too.
"""
import numpy as np
import pytest
import torch

from detectron2_tensorflow_b200.layers import batch_nms, matrix_nms, segmented_top_k
from detectron2_tensorflow_b200.modeling import (Box2BoxTransform, RPNOutputs, RetinaNetInference, fast_rcnn_inference,
                                                 find_top_rpn_proposals)
from detectron2_tensorflow_b200.structures import BoxList, ImageList, SparseBoxList
from detectron2_tensorflow_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def rand_boxes(rng, n, H=800, W=1333, smin=8, smax=400):
    cy, cx = rng.uniform(0, H, n), rng.uniform(0, W, n)
    h, w = rng.uniform(smin, smax, n), rng.uniform(smin, smax, n)
    return np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1).astype(np.float32)


def clustered_boxes(rng, n, k=12):
    c = rand_boxes(rng, k)
    b = c[rng.integers(0, k, n)] + rng.normal(0, 6, (n, 4)).astype(np.float32)
    return b.astype(np.float32)


# ------------------------------------------------------------------ decode
def test_apply_deltas(cuda, oracle_lib):
    rng = np.random.default_rng(0)
    n, k = 3000, 5
    boxes = rand_boxes(rng, n)
    deltas = rng.standard_normal((n, k * 4)).astype(np.float32)
    deltas[::7, 2] = 50.0  # hits the scale clamp
    deltas[::11, 3] = -30.0
    for w in ((1., 1., 1., 1.), (10., 10., 5., 5.)):
        want = oracle_lib.apply_deltas(deltas, boxes, w)
        got = Box2BoxTransform(w).apply_deltas(T(deltas, cuda), T(boxes, cuda)).cpu().numpy()
        assert np.array_equal(got, want)


# ------------------------------------------------------------------ top-k
@pytest.mark.parametrize("variant", ["gaussian", "ties", "heavy_ties", "special"])
def test_segmented_topk(cuda, oracle_lib, variant):
    rng = np.random.default_rng(1)
    lens = [40000, 5000, 819, 33, 1]
    N, k = 3, 1000
    rows = []
    for ln in lens:
        x = (rng.standard_normal((N, ln)) * 2).astype(np.float32)
        if variant == "ties":
            x = (np.round(x * 64) / 64).astype(np.float32)
        elif variant == "heavy_ties":
            x = np.round(x).astype(np.float32)  # a dozen distinct values: boundary bucket is huge
        elif variant == "special":
            x[:, ::5] = 0.0
            x[:, 1::50] = -0.0
            x[:, 2::97] = np.nan
            x[:, 3::89] = -np.inf
            x[:, 4::101] = np.inf
        rows.append(x)
    vals, idx, cnt = segmented_top_k([T(x, cuda) for x in rows], k)
    vals, idx, cnt = vals.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
    for n in range(N):
        for l, x in enumerate(rows):
            wv, wi = oracle_lib.top_k(x[n], k)
            kr = len(wi)
            assert cnt[n, l] == kr == min(k, x.shape[1])
            assert np.array_equal(idx[n, l, :kr], wi), (variant, n, l)
            assert np.array_equal(vals[n, l, :kr], wv, equal_nan=True)
            assert np.all(idx[n, l, kr:] == -1) and np.all(vals[n, l, kr:] == 0)


def test_segmented_topk_sigmoid_and_limit(cuda, oracle_lib):
    rng = np.random.default_rng(2)
    K = 7
    hwa = [900, 60, 5]
    rows = [(rng.standard_normal((2, h * K)) * 6).astype(np.float32) for h in hwa]  # saturates -> ties
    k = 100
    vals, idx, cnt = segmented_top_k([T(x, cuda) for x in rows], k, sigmoid=True, k_limits=hwa)
    for n in range(2):
        for l, x in enumerate(rows):
            p = oracle_lib.sigmoidf(x[n])
            wv, wi = oracle_lib.top_k(p, min(k, hwa[l]))
            kr = len(wi)
            assert cnt[n, l].item() == kr
            assert np.array_equal(idx[n, l, :kr].cpu().numpy(), wi)
            assert np.array_equal(vals[n, l, :kr].cpu().numpy(), wv)


# ------------------------------------------------------------------ NMS
@pytest.mark.parametrize("n,max_out,thr,kind", [
    (1000, 1000, 0.7, "random"), (2000, 1000, 0.7, "clustered"), (777, 50, 0.5, "clustered"),
    (5000, 100, 0.5, "clustered"),   # capped lazy sweep (cap << n)
    (4500, 4500, 0.6, "random"),     # bitmask, several words per lane
    (300, 300, 0.3, "ties"), (64, 64, 0.5, "degenerate"), (1, 5, 0.5, "random"), (0, 5, 0.5, "random")])
def test_batched_nms(cuda, oracle_lib, n, max_out, thr, kind):
    rng = np.random.default_rng(n + max_out)
    B = 3
    boxes = np.zeros((B, n, 4), np.float32)
    scores = np.zeros((B, n), np.float32)
    for b in range(B):
        bx = clustered_boxes(rng, n) if kind in ("clustered", "ties") else rand_boxes(rng, n)
        sc = rng.standard_normal(n).astype(np.float32)
        if kind == "ties":
            sc = np.round(sc * 4).astype(np.float32) / 4
            bx[n // 2:] = bx[:n - n // 2]  # exact duplicates
        if kind == "degenerate" and n:
            bx[::4, 2] = bx[::4, 0]            # zero height
            bx[1::8] = bx[1::8][:, [2, 3, 0, 1]]  # inverted corners
            sc[::9] = -np.inf
            sc[5::13] = np.nan
        boxes[b], scores[b] = bx, sc
    keep, num = batch_nms(T(boxes, cuda), T(scores, cuda), max_out, axis=1, iou_threshold=thr)
    keep, num = keep.cpu().numpy(), num.cpu().numpy()
    for b in range(B):
        want = oracle_lib.nms(boxes[b], scores[b], max_out, thr)
        assert num[b] == len(want)
        assert np.array_equal(keep[b, :len(want)], want), (kind, b)
        assert np.all(keep[b, len(want):] == -1)


def test_batch_nms_axis0_and_idempotence(cuda, oracle_lib):
    rng = np.random.default_rng(7)
    n, B = 400, 2
    boxes = np.stack([clustered_boxes(rng, n) for _ in range(B)])
    scores = rng.standard_normal((B, n)).astype(np.float32)
    k1, n1 = batch_nms(T(boxes.transpose(1, 0, 2), cuda), T(scores.T, cuda), n, axis=0, iou_threshold=0.5)
    k2, n2 = batch_nms(T(boxes, cuda), T(scores, cuda), n, axis=1, iou_threshold=0.5)
    assert torch.equal(k1, k2) and torch.equal(n1, n2)
    # idempotence: NMS of the survivors keeps all of them
    for b in range(B):
        kept = k2[b, :n2[b]].long().cpu().numpy()
        k3, n3 = batch_nms(T(boxes[b][kept][None], cuda), T(scores[b][kept][None], cuda), n, axis=1, iou_threshold=0.5)
        assert n3[0].item() == len(kept)


# ------------------------------------------------------------------ RPN
def _small_rpn(seed, variant, N=2):
    padded = (160, 224)
    anchors = syn.rpn_anchors(padded_hw=padded)
    rng = np.random.default_rng(seed)
    logits, deltas = [], []
    for a in anchors:
        lg = (rng.standard_normal((N, a.shape[0])) * 2).astype(np.float32)
        if variant == "ties":
            lg = (np.round(lg * 8) / 8).astype(np.float32)
        d = (rng.standard_normal((N, a.shape[0], 4)) * np.array([0.5, 0.5, 0.25, 0.25])).astype(np.float32)
        if variant == "clustered":
            d *= 0.05
        logits.append(lg)
        deltas.append(d)
    shapes = np.array([[150, 200], [160, 224]][:N], np.int32)
    return anchors, logits, deltas, shapes


@pytest.mark.parametrize("variant,pre,post,min_len", [("gaussian", 300, 200, 0.0), ("ties", 200, 150, 0.0),
                                                      ("clustered", 500, 100, 0.0), ("gaussian", 300, 200, 12.0),
                                                      ("clustered", 2000, 1000, 4.0)])
def test_rpn_proposals(cuda, oracle_lib, variant, pre, post, min_len):
    anchors, logits, deltas, shapes = _small_rpn(11, variant)
    props = [oracle_lib.rpn_predict_proposals(d, a) for d, a in zip(deltas, anchors)]
    wb, wl, wv, wn = oracle_lib.find_top_rpn_proposals(props, logits, shapes, 0.7, pre, post, min_len)
    images = ImageList(None, T(shapes, cuda))
    # (1) reference signature: decoded proposals in
    res = find_top_rpn_proposals([T(p, cuda) for p in props], [T(x, cuda) for x in logits], images, 0.7, pre, post,
                                 min_len)
    # (2) fused: deltas + anchors, only the winners are decoded
    outs = RPNOutputs(Box2BoxTransform((1., 1., 1., 1.)), images,
                      [T(x, cuda) for x in logits], [T(d, cuda) for d in deltas], [T(a, cuda) for a in anchors])
    res2 = outs.find_top_proposals(0.7, pre, post, min_len)
    for r in (res, res2):
        assert np.array_equal(r.get_field("is_valid").cpu().numpy(), wv)
        assert np.array_equal(r.boxes.cpu().numpy(), wb)
        assert np.array_equal(r.get_field("objectness_logits").cpu().numpy(), wl)
    # the decode-all path of the reference is also available and bit-exact
    pp = outs.predict_proposals()
    for a, b in zip(pp, props):
        assert np.array_equal(a.cpu().numpy(), b)


@pytest.mark.gpu
@pytest.mark.parametrize("variant,pre,post,min_len", [("ties", 200, 150, 0.0), ("clustered", 2000, 1000, 4.0)])
def test_rpn_proposals_generic_chain(cuda, oracle_lib, monkeypatch, variant, pre, post, min_len):
    """The generic multi-launch chain (used when pre_nms_topk > 4096) on the same inputs as the cluster-fused path."""
    monkeypatch.setenv("D2B_RPN_GENERIC", "1")
    test_rpn_proposals(cuda, oracle_lib, variant, pre, post, min_len)


@pytest.mark.gpu
def test_rpn_proposals_large_k_uses_generic_chain(cuda, oracle_lib):
    """pre_nms_topk above the fused kernel's 4096-key limit on a row that is longer than that."""
    rng = np.random.default_rng(5)
    N, hwa = 2, [9000, 700]
    props = [np.stack([fuzz_like_boxes(rng, n) for _ in range(N)]) for n in hwa]
    logits = [(np.round(rng.standard_normal((N, n)) * 16) / 16).astype(np.float32) for n in hwa]
    shapes = np.array([[300, 400], [280, 390]], np.int32)
    wb, wl, wv, _ = oracle_lib.find_top_rpn_proposals(props, logits, shapes, 0.7, 6000, 1000, 0.0)
    res = find_top_rpn_proposals([T(p, cuda) for p in props], [T(x, cuda) for x in logits],
                                 ImageList(None, T(shapes, cuda)), 0.7, 6000, 1000, 0.0)
    assert np.array_equal(res.get_field("is_valid").cpu().numpy(), wv)
    assert np.array_equal(res.boxes.cpu().numpy(), wb)
    assert np.array_equal(res.get_field("objectness_logits").cpu().numpy(), wl)


def fuzz_like_boxes(rng, n):
    cy, cx = rng.uniform(0, 300, n), rng.uniform(0, 400, n)
    h, w = rng.uniform(4, 120, n), rng.uniform(4, 120, n)
    return np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1).astype(np.float32)


def test_rpn_full_size_one_image(cuda, oracle_lib):
    """Config 1 shapes: one 800x1333 image, 268,569 anchors, 1000 pre / 1000 post (test-time FPN settings)."""
    anchors = syn.rpn_anchors()
    logits, deltas = syn.rpn_inputs(1, seed=2, variant="clustered", anchors=anchors)
    shapes = syn.image_shapes(1)
    props = [oracle_lib.rpn_predict_proposals(d, a) for d, a in zip(deltas, anchors)]
    wb, wl, wv, wn = oracle_lib.find_top_rpn_proposals(props, logits, shapes, 0.7, 1000, 1000, 0.0)
    outs = RPNOutputs(Box2BoxTransform((1., 1., 1., 1.)), ImageList(None, T(shapes, cuda)),
                      [T(x, cuda) for x in logits], [T(d, cuda) for d in deltas], [T(a, cuda) for a in anchors])
    r = outs.find_top_proposals(0.7, 1000, 1000, 0.0)
    assert np.array_equal(r.get_field("is_valid").cpu().numpy(), wv)
    assert np.array_equal(r.boxes.cpu().numpy(), wb)
    assert np.array_equal(r.get_field("objectness_logits").cpu().numpy(), wl)


# ------------------------------------------------------------------ Fast R-CNN
@pytest.mark.parametrize("agnostic_reg,agnostic_nms,R,K,topk", [(False, False, 120, 20, 100), (True, False, 80, 10, 30),
                                                                (False, True, 60, 8, 100), (False, False, 1000, 80, 100)])
def test_fast_rcnn_inference(cuda, oracle_lib, agnostic_reg, agnostic_nms, R, K, topk):
    rng = np.random.default_rng(R + K)
    N = 2
    # ragged: image 1 has fewer valid proposals
    valid = np.ones((N, R), bool)
    valid[1, R // 2:] = False
    idx = np.argwhere(valid).astype(np.int64)
    M = idx.shape[0]
    prop = clustered_boxes(rng, M, k=15)
    Kb = 1 if agnostic_reg else K
    deltas = (rng.standard_normal((M, Kb * 4)) * 0.5).astype(np.float32)
    boxes = oracle_lib.apply_deltas(deltas, prop, (10., 10., 5., 5.))
    lg = rng.standard_normal((M, K + 1)) * 3
    e = np.exp(lg - lg.max(1, keepdims=True))
    scores = (e / e.sum(1, keepdims=True)).astype(np.float32)
    shapes = np.array([[800, 1333], [750, 1200]], np.int32)
    wb, ws, wc, wv, wr, wn = oracle_lib.fast_rcnn_inference(boxes, scores, idx, (N, R), shapes, 0.05, 0.5, topk,
                                                            agnostic_nms)
    inst = SparseBoxList(T(idx, cuda), BoxList(T(prop, cuda)), (N, R))
    inst.set_tracking('image_shape', T(shapes, cuda))
    res, kept = fast_rcnn_inference(T(boxes, cuda), T(scores, cuda), inst, 0.05, 0.5, topk, agnostic_nms)
    assert np.array_equal(res.get_field('is_valid').cpu().numpy(), wv)
    assert np.array_equal(res.get_field('pred_classes').cpu().numpy(), wc)
    assert res.get_field('pred_classes').dtype == torch.int64
    assert np.array_equal(res.get_field('scores').cpu().numpy(), ws)
    assert np.array_equal(res.boxes.cpu().numpy(), wb)
    assert np.array_equal(kept.cpu().numpy(), wr)
    # fused decode (FastRCNNOutputs.inference): deltas + proposals in, the boxes decoded inside the kernels
    res2, kept2 = fast_rcnn_inference(None, T(scores, cuda), inst, 0.05, 0.5, topk, agnostic_nms,
                                      pred_proposal_deltas=T(deltas, cuda),
                                      box2box_transform=Box2BoxTransform((10., 10., 5., 5.)))
    for f in ('is_valid', 'pred_classes', 'scores'):
        assert torch.equal(res2.get_field(f), res.get_field(f)), f
    assert torch.equal(res2.boxes, res.boxes) and torch.equal(kept2, kept)
    with pytest.raises(ValueError):
        fast_rcnn_inference(None, T(scores, cuda), inst, 0.05, 0.5, topk, agnostic_nms)


# ------------------------------------------------------------------ RetinaNet
def test_retinanet_inference(cuda, oracle_lib):
    rng = np.random.default_rng(21)
    N, K = 2, 12
    anchors = syn.retinanet_anchors(padded_hw=(256, 320))
    cls = [(rng.standard_normal((N, a.shape[0], K)) * 1.5 - 2.5).astype(np.float32) for a in anchors]
    dl = [(rng.standard_normal((N, a.shape[0], 4)) * 0.3).astype(np.float32) for a in anchors]
    wb, ws, wc, wv, wn = oracle_lib.retinanet_inference(cls, dl, anchors, K, 300, 0.05, 0.5, 100)
    head = RetinaNetInference(num_classes=K, topk_candidates=300)
    res = head.inference([T(x, cuda) for x in cls], [T(x, cuda) for x in dl], [T(a, cuda) for a in anchors])
    assert np.array_equal(res.get_field('is_valid').cpu().numpy(), wv)
    assert np.array_equal(res.get_field('pred_classes').cpu().numpy(), wc)
    assert res.get_field('pred_classes').dtype == torch.int32
    assert np.array_equal(res.get_field('scores').cpu().numpy(), ws)
    assert np.array_equal(res.boxes.cpu().numpy(), wb)
    assert wn.sum() > 0


# ------------------------------------------------------------------ Matrix NMS
@pytest.mark.parametrize("kernel", ["gaussian", "linear"])
def test_matrix_nms(cuda, oracle_lib, kernel):
    masks, classes, scores = syn.solo_masks(120, hw=(50, 84), num_classes=6, seed=7)
    want = oracle_lib.matrix_nms(masks, classes, scores, None, kernel, 2.0)
    got = matrix_nms(T(masks, cuda), T(classes, cuda), T(scores, cuda), kernel=kernel, sigma=2.0).cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    sm = masks.reshape(120, -1).sum(1)
    got2 = matrix_nms(T(masks, cuda), T(classes, cuda), T(scores, cuda), sum_masks=T(sm, cuda), kernel=kernel).cpu().numpy()
    assert np.array_equal(got2, want, equal_nan=True)
    # batched launch == per-image launches
    m2, c2, s2 = syn.solo_masks(120, hw=(50, 84), num_classes=6, seed=8)
    gb = matrix_nms(T(np.stack([masks, m2]), cuda), T(np.stack([classes, c2]), cuda), T(np.stack([scores, s2]), cuda),
                    kernel=kernel).cpu().numpy()
    assert np.array_equal(gb[0], want, equal_nan=True)
    assert np.array_equal(gb[1], oracle_lib.matrix_nms(m2, c2, s2, None, kernel, 2.0), equal_nan=True)
    with pytest.raises(NotImplementedError):
        matrix_nms(T(masks, cuda), T(classes, cuda), T(scores, cuda), kernel="cosine")


@pytest.mark.parametrize("kernel", ["gaussian", "linear"])
def test_matrix_nms_empty_masks_and_big_class(cuda, oracle_lib, kernel):
    """Masks that include EMPTY ones (0 / 0 unions -> the NaN rules of the column maximum, which the kernels derive
    without reading the matrix) and one class that holds most masks."""
    masks, classes, scores = syn.solo_masks(150, hw=(40, 70), num_classes=5, seed=11)
    masks[[0, 7, 8, 60]] = 0.0
    classes[20:90] = 3
    classes[7] = classes[8]
    want = oracle_lib.matrix_nms(masks, classes, scores, None, kernel, 2.0)
    got = matrix_nms(T(masks, cuda), T(classes, cuda), T(scores, cuda), kernel=kernel, sigma=2.0).cpu().numpy()
    if kernel == "gaussian":  # (linear: 0/0 next to finite values in a column is order-dependent in TF)
        assert np.array_equal(got, want, equal_nan=True)
    else:
        ok = ~np.isnan(want)
        assert np.array_equal(got[ok], want[ok])


def test_matrix_nms_many_masks(cuda, oracle_lib):
    """n > 4096: the decay kernel's per-row table no longer fits shared memory (its global-load variant runs)."""
    rng = np.random.default_rng(5)
    n, H, W = 4200, 6, 10
    masks = (rng.random((n, H, W)) < 0.3).astype(np.float32)
    masks[rng.integers(0, n, 40)] = 0.0
    classes = rng.integers(0, 30, n).astype(np.int64)
    scores = np.sort(rng.uniform(0.1, 1.0, n).astype(np.float32))[::-1].copy()
    want = oracle_lib.matrix_nms(masks, classes, scores, None, "gaussian", 2.0)
    got = matrix_nms(T(masks, cuda), T(classes, cuda), T(scores, cuda), kernel="gaussian", sigma=2.0).cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)


# ------------------------------------------------------------------ sigmoid top-k: cutoff / candidate-list paths
@pytest.mark.parametrize("case", ["typical", "plateau_overflow", "tied_boundary", "negative_tail", "saturated",
                                  "sampled_long", "sample_misleads"])
def test_sigmoid_topk_paths(cuda, oracle_lib, case):
    """The RetinaNet top-k has a logit pre-histogram cutoff and a candidate list; exercise: the fast path,
    a candidate list that overflows its capacity (falls back to re-scanning the row), index ties exactly at
    the k-th value, cutoffs in the negative range, and the saturated regime where the shortcut is disabled.
    Long rows take their cutoff from a 1/16 SAMPLE of the row: `sampled_long` is that fast path, and
    `sample_misleads` plants the large logits exactly where the sampler looks so the cutoff keeps fewer than k
    elements and pass 0 must detect it by exact count and repeat without a cutoff."""
    rng = np.random.default_rng({"typical": 1, "plateau_overflow": 2, "tied_boundary": 3, "negative_tail": 4, "saturated": 5,
                                 "sampled_long": 6, "sample_misleads": 7}[case])
    n, k = 300000, 1000
    if case in ("sampled_long", "sample_misleads"):
        n = 1300000
        x = (rng.standard_normal(n) * 1.5 - 4.6).astype(np.float32)
        if case == "sample_misleads":
            for c0 in range(0, n, 65536):
                x[c0 + rng.integers(0, 4096, 12)] = 3.0 + rng.random(12).astype(np.float32)
    elif case == "typical":
        x = (rng.standard_normal(n) * 1.5 - 4.6).astype(np.float32)
    elif case == "plateau_overflow":
        x = np.full(n, 1.0, np.float32)          # > 65536 elements above the cutoff, all tied
        x[rng.integers(0, n, 500)] = 3.0
    elif case == "tied_boundary":
        x = (rng.standard_normal(n) * 1.5 - 4.6).astype(np.float32)
        kth = np.sort(x)[-k]
        x[rng.integers(0, n, 400)] = kth         # many exact copies of the k-th value
    elif case == "negative_tail":
        x = (rng.standard_normal(n) * 0.5 - 30.0).astype(np.float32)
    else:
        x = (rng.standard_normal(n) * 40.0).astype(np.float32)
    vals, idx, cnt = segmented_top_k([T(x[None], cuda)], k, sigmoid=True)
    p = oracle_lib.sigmoid_array(x)
    wv, wi = oracle_lib.top_k(p, k)
    assert cnt[0, 0].item() == k
    assert np.array_equal(idx[0, 0].cpu().numpy(), wi), case
    assert np.array_equal(vals[0, 0].cpu().numpy(), wv)


# ------------------------------------------------------------------ in-kernel anchors (SURVEY.md 8f "next" #2)
def test_anchor_generator_and_in_kernel_anchors(cuda, oracle_lib):
    from detectron2_tensorflow_b200.modeling import DefaultAnchorGenerator
    padded = (160, 224)
    strides = syn.RPN_STRIDES
    gen = DefaultAnchorGenerator([[s] for s in syn.RPN_SIZES], [list(syn.ASPECT_RATIOS)], strides, device=cuda)
    grids = [syn.level_hw(s, padded) for s in strides]
    tables = gen.grid_anchors(grids)
    want_tables = syn.rpn_anchors(padded_hw=padded)          # numpy restatement of anchor_generator.py:92-144
    for t, w in zip(tables, want_tables):
        assert np.array_equal(t.cpu().numpy(), w)
    anchors, logits, deltas, shapes = _small_rpn(13, "gaussian")
    props = [oracle_lib.rpn_predict_proposals(d, a) for d, a in zip(deltas, anchors)]
    wb, wl, wv, _ = oracle_lib.find_top_rpn_proposals(props, logits, shapes, 0.7, 300, 200, 0.0)
    outs = RPNOutputs(Box2BoxTransform((1., 1., 1., 1.)), ImageList(None, T(shapes, cuda)),
                      [T(x, cuda) for x in logits], [T(d, cuda) for d in deltas], gen.grid_descriptors(grids))
    r = outs.find_top_proposals(0.7, 300, 200, 0.0)
    assert np.array_equal(r.boxes.cpu().numpy(), wb) and np.array_equal(r.get_field("is_valid").cpu().numpy(), wv)
    assert np.array_equal(r.get_field("objectness_logits").cpu().numpy(), wl)
    # RetinaNet with synthesised anchors
    rng = np.random.default_rng(3)
    K = 5
    rstr = syn.RETINA_STRIDES
    rgen = DefaultAnchorGenerator([[s * 4 * 2 ** (i / 3.0) for i in range(3)] for s in rstr], [list(syn.ASPECT_RATIOS)],
                                  rstr, device=cuda)
    rgrids = [syn.level_hw(s, (256, 320)) for s in rstr]
    ran = [a.cpu().numpy() for a in rgen.grid_anchors(rgrids)]
    cls = [(rng.standard_normal((2, a.shape[0], K)) * 1.5 - 2.5).astype(np.float32) for a in ran]
    dl = [(rng.standard_normal((2, a.shape[0], 4)) * 0.3).astype(np.float32) for a in ran]
    wb, ws, wc, wv, wn = oracle_lib.retinanet_inference(cls, dl, ran, K, 200, 0.05, 0.5, 50)
    head = RetinaNetInference(num_classes=K, topk_candidates=200, max_detections_per_image=50)
    res = head.inference([T(x, cuda) for x in cls], [T(x, cuda) for x in dl], rgen.grid_descriptors(rgrids))
    assert np.array_equal(res.boxes.cpu().numpy(), wb) and np.array_equal(res.get_field('scores').cpu().numpy(), ws)
    assert np.array_equal(res.get_field('pred_classes').cpu().numpy(), wc)


# ------------------------------------------------------------------ YOLOv4 post-processing / point_nms (8f #4)
@pytest.mark.parametrize("n,K,topk", [(22743, 80, 100), (3000, 20, 300), (50, 3, 10), (0, 5, 10)])
def test_yolo_postprocess(cuda, oracle_lib, n, K, topk):
    from detectron2_tensorflow_b200.modeling import YOLOv4Inference
    rng = np.random.default_rng(n + K)
    N = 3
    boxes = np.stack([clustered_boxes(rng, n, k=40) if n else np.zeros((0, 4), np.float32) for _ in range(N)])
    probs = (rng.random((N, n, K)) ** 6).astype(np.float32)
    if n:
        probs[:, ::5] = np.round(probs[:, ::5] * 8) / 8  # class ties (first argmax) and score ties (index order)
        probs[1] *= 0.01  # an image with no candidate at all
    want = oracle_lib.yolo_inference(boxes, probs, 0.25, 0.45, topk)
    res = YOLOv4Inference(0.25, 0.45, topk).inference(T(boxes, cuda), T(probs, cuda))
    assert np.array_equal(res.get_field("is_valid").cpu().numpy(), want[3])
    assert np.array_equal(res.get_field("pred_classes").cpu().numpy(), want[2])
    assert np.array_equal(res.get_field("scores").cpu().numpy(), want[1])
    assert np.array_equal(res.boxes.cpu().numpy(), want[0])


@pytest.mark.parametrize("shape", [(2, 40, 40, 80), (1, 12, 12, 80), (2, 7, 5, 3), (1, 1, 1, 4)])
def test_point_nms(cuda, oracle_lib, shape):
    from detectron2_tensorflow_b200.modeling import point_nms
    rng = np.random.default_rng(sum(shape))
    x = (np.round(rng.standard_normal(shape) * 3) / 3).astype(np.float32)
    assert np.array_equal(point_nms(T(x, cuda)).cpu().numpy(), oracle_lib.point_nms(x))


@pytest.mark.parametrize("hw,thr", [((40, 64), 0.5), ((25, 37), 0.5), ((200, 336), 0.5), ((16, 20), 0.9995), ((16, 20), 0.0)])
def test_solo_mask_encode_and_packed_matrix_nms(cuda, oracle_lib, hw, thr):
    """Mask stage of SOLOv2Head.inference (solo_v2.py:513-533) + Matrix-NMS fed with the packed words."""
    from detectron2_tensorflow_b200.modeling import solo_mask_encode
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    n = 60 if hw[0] < 100 else 24
    H, W = hw
    # smooth blobs so masks are object-like; values dense around logit(thr) to stress the guard band
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    logits = np.empty((n, H, W), np.float32)
    for i in range(n):
        cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(3, max(H, W) / 2)
        logits[i] = (r - np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)) * rng.uniform(0.01, 2.0)
    if 0.0 < thr < 1.0:
        x0 = np.float32(np.log(thr / (1 - thr)))
        near = x0 + np.arange(-40, 41, dtype=np.float32) * np.spacing(x0) * 0.5  # +-20 ulp around the crossing
        logits[0].ravel()[:near.size] = near
        logits[1] = x0
    logits[2, 0, :4] = [np.nan, np.inf, -np.inf, 0.0]
    logits[n // 2:] = logits[:n - n // 2] + rng.normal(0, 0.05, (n - n // 2, H, W)).astype(np.float32)  # duplicates
    masks, sm, ss = oracle_lib.solo_mask_stage(logits, thr)
    packed, gsm, gss = solo_mask_encode(T(logits, cuda), thr)
    bits = np.unpackbits(packed.cpu().numpy().view(np.uint8), axis=-1, bitorder="little")[:, :H * W]
    assert np.array_equal(bits.reshape(n, H, W), masks.astype(np.uint8))
    assert np.array_equal(gsm.cpu().numpy(), sm)
    assert np.allclose(gss.cpu().numpy(), ss, rtol=1e-5, atol=1e-6)  # fp32 sum order (tf.reduce_sum is unspecified too)
    # Matrix-NMS on the packed words == Matrix-NMS on the fp32 masks (both CUDA) == oracle
    classes = rng.integers(0, 3, n).astype(np.int64)
    scores = np.sort(rng.uniform(0.1, 1, n).astype(np.float32))[::-1].copy()
    want = oracle_lib.matrix_nms(masks, classes, scores, sm, "gaussian", 2.0)
    got_packed = matrix_nms(None, T(classes, cuda), T(scores, cuda), sum_masks=gsm, packed_masks=packed, mask_hw=H * W)
    got_dense = matrix_nms(T(masks, cuda), T(classes, cuda), T(scores, cuda), sum_masks=T(sm, cuda))
    assert np.array_equal(got_packed.cpu().numpy(), want, equal_nan=True)
    assert np.array_equal(got_dense.cpu().numpy(), want, equal_nan=True)
    got_nosum = matrix_nms(None, T(classes, cuda), T(scores, cuda), packed_masks=packed, mask_hw=H * W)
    assert np.array_equal(got_nosum.cpu().numpy(), want, equal_nan=True)


@pytest.mark.parametrize("n,hw,pre,D,kern", [(300, (40, 64), 120, 40, "gaussian"), (77, (25, 37), 500, 100, "linear"),
                                             (1, (8, 8), 5, 3, "gaussian"), (600, (50, 84), 500, 100, "gaussian")])
def test_solo_postprocess(cuda, oracle_lib, n, hw, pre, D, kern):
    """SOLOv2Head.inference tail after the dynamic conv (solo_v2.py:507-558), batched with ragged counts."""
    from detectron2_tensorflow_b200.modeling import SOLOv2Inference
    rng = np.random.default_rng(n + D)
    B, (H, W) = 3, hw
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    logits = np.empty((B, n, H, W), np.float32)
    for b in range(B):
        for i in range(n):
            cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(1, max(H, W) / 2)
            logits[b, i] = (r - np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)) * rng.uniform(0.05, 2.0)
        logits[b, n // 2:] = logits[b, :n - n // 2] + rng.normal(0, 0.02, (n - n // 2, H, W)).astype(np.float32)
    scores = rng.uniform(0.1, 1.0, (B, n)).astype(np.float32)
    classes = rng.integers(0, 3, (B, n)).astype(np.int64)
    strides = rng.choice([8.0, 16.0, 32.0], (B, n)).astype(np.float32)
    counts = np.array([n, max(n - 7, 0), 0], np.int32)
    head = SOLOv2Inference(0.5, pre, kern, 2.0, 0.05, D)
    got = head.postprocess(T(logits, cuda), T(scores, cuda), T(classes, cuda), T(strides, cuda), T(counts, cuda))
    for b in range(B):
        c = int(counts[b])
        wm, wc, ws, wv, wn = oracle_lib.solo_postprocess(logits[b, :c], scores[b, :c], classes[b, :c], strides[b, :c], 0.5,
                                                        pre, kern, 2.0, 0.05, D)
        assert int(got["num"][b]) == wn
        assert np.array_equal(got["is_valid"][b].cpu().numpy(), wv)
        assert np.array_equal(got["pred_classes"][b].cpu().numpy(), wc)
        # scores carry the fp32 mask-scoring sums (order of summation differs): 1e-5 relative
        assert np.allclose(got["scores"][b].cpu().numpy(), ws, rtol=1e-5, atol=1e-7, equal_nan=True)
        assert np.array_equal(got["pred_masks"][b].cpu().numpy(), wm)
        bits = np.unpackbits(got["packed_masks"][b].cpu().numpy().view(np.uint8), axis=-1, bitorder="little")[:, :H * W]
        assert np.array_equal(bits.reshape(D, H, W), wm.astype(np.uint8))
    assert int(got["num"].sum()) > 0
