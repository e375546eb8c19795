"""GPU: randomized differential testing of the CUDA path against the oracle -- many small, irregular
configurations per operator (sizes that are not multiples of anything, empty inputs, heavy ties, duplicates,
degenerate boxes, out-of-range indices).  Seeds are fixed: every case is reproducible from its parameter id.
Bar as everywhere: integer outputs and elementwise fp32 values bit-exact."""
import numpy as np
import pytest
import torch

from detectron2_tensorflow_b200.layers import batch_nms, matrix_nms, segmented_top_k
from detectron2_tensorflow_b200.modeling import (Matcher, ROIPooler, YOLOv4Inference, fast_rcnn_inference,
                                                 find_top_rpn_proposals, label_boxes)
from detectron2_tensorflow_b200.structures import BoxList, ImageList, SparseBoxList

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def fuzz_boxes(rng, n, H, W):
    """Mixture: random, clustered near-duplicates, exact duplicates, zero-area, inverted, partly outside."""
    if n == 0:
        return np.zeros((0, 4), np.float32)
    cy, cx = rng.uniform(-20, H + 20, n), rng.uniform(-20, W + 20, n)
    h, w = np.exp(rng.uniform(np.log(2), np.log(H), n)), np.exp(rng.uniform(np.log(2), np.log(W), n))
    b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1).astype(np.float32)
    k = max(n // 3, 1)
    src = rng.integers(0, n, k)
    dst = rng.integers(0, n, k)
    b[dst] = b[src] + rng.normal(0, rng.choice([0.0, 0.5, 4.0]), (k, 4)).astype(np.float32)
    if rng.random() < 0.5:
        b = np.round(b)  # integer coordinates: many exactly equal IoUs
    z = rng.integers(0, n, max(n // 20, 1))
    b[z, 2] = b[z, 0]
    inv = rng.integers(0, n, max(n // 25, 1))
    b[inv] = b[inv][:, [2, 3, 0, 1]]
    return b.astype(np.float32)


def fuzz_scores(rng, shape):
    s = rng.standard_normal(shape).astype(np.float32) * rng.choice([0.1, 1.0, 10.0])
    mode = rng.integers(0, 4)
    if mode == 1:
        s = np.round(s * 2) / 2
    elif mode == 2:
        s = np.round(s)
    elif mode == 3 and s.size:
        flat = s.reshape(-1)
        flat[rng.integers(0, flat.size, max(flat.size // 15, 1))] = -np.inf
        flat[rng.integers(0, flat.size, max(flat.size // 40, 1))] = np.inf
    return s.astype(np.float32)


@pytest.mark.parametrize("seed", range(60))
def test_fuzz_batched_nms(cuda, oracle_lib, seed):
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 6))
    n = int(rng.choice([0, 1, 2, 63, 64, 65, 129, 500, 1111, 2049, 3000]))
    max_out = int(rng.choice([1, 7, 100, max(n, 1), 2 * max(n, 1)]))
    thr = float(rng.choice([0.0, 0.3, 0.5, 0.7, 0.95, 1.0]))
    boxes = np.stack([fuzz_boxes(rng, n, 300, 400) for _ in range(B)])
    scores = fuzz_scores(rng, (B, n))
    keep, num = batch_nms(T(boxes, cuda), T(scores, cuda), max_out, axis=1, iou_threshold=thr)
    keep, num = keep.cpu().numpy(), num.cpu().numpy()
    for b in range(B):
        want = oracle_lib.nms(boxes[b], scores[b], max_out, thr)
        assert num[b] == len(want), (seed, b)
        assert np.array_equal(keep[b, :len(want)], want), (seed, b)
        assert np.all(keep[b, len(want):] == -1)


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_topk(cuda, oracle_lib, seed):
    rng = np.random.default_rng(2000 + seed)
    N = int(rng.integers(1, 4))
    L = int(rng.integers(1, 5))
    lens = [int(rng.choice([1, 2, 31, 257, 1000, 8191, 8193, 20000, 70001])) for _ in range(L)]
    k = int(rng.choice([1, 5, 100, 1000, 2000]))
    sig = bool(rng.random() < 0.4)
    rows = [fuzz_scores(rng, (N, ln)) for ln in lens]
    if sig:
        rows = [np.where(np.isinf(r), np.float32(3.0), r) * np.float32(0.5) - np.float32(2.0) for r in rows]
    vals, idx, cnt = segmented_top_k([T(r, cuda) for r in rows], k, sigmoid=sig)
    vals, idx, cnt = vals.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
    for n in range(N):
        for l, r in enumerate(rows):
            x = oracle_lib.sigmoid_array(r[n]) if sig else r[n]
            wv, wi = oracle_lib.top_k(x, k)
            kr = len(wi)
            assert cnt[n, l] == kr
            assert np.array_equal(idx[n, l, :kr], wi), (seed, n, l)
            assert np.array_equal(vals[n, l, :kr], wv, equal_nan=True)


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_roi_pooler(cuda, oracle_lib, seed):
    rng = np.random.default_rng(3000 + seed)
    N = int(rng.integers(1, 4))
    L = int(rng.choice([1, 2, 4]))
    Cc = int(rng.choice([4, 12, 128, 256]))
    first = int(rng.choice([2, 3]))
    scales = [1.0 / (1 << (first + l)) for l in range(L)]
    H0, W0 = int(rng.integers(20, 70)), int(rng.integers(20, 90))
    feats = [rng.standard_normal((N, max(H0 >> l, 1), max(W0 >> l, 1), Cc)).astype(np.float32) for l in range(L)]
    M = int(rng.choice([0, 1, 17, 200]))
    img_h, img_w = H0 << first, W0 << first
    boxes = fuzz_boxes(rng, M, img_h, img_w)
    bidx = rng.integers(0, N, M).astype(np.int64)
    if M > 5:
        bidx[3] = N + 2  # out of range -> zero row
    osz = (int(rng.choice([1, 2, 7, 14])), int(rng.choice([1, 3, 7, 14])))
    sr = int(rng.choice([0, 0, 1, 2, 3]))
    aligned = bool(rng.random() < 0.7)
    want, wc = oracle_lib.roi_pooler(feats, scales, boxes, bidx, osz, sr, aligned, canonical_box_size=56,
                                     canonical_level=min(first + L - 1, max(first, 4)))
    pooler = ROIPooler(osz, scales, sr, "ROIAlignV2" if aligned else "ROIAlign", canonical_box_size=56,
                       canonical_level=min(first + L - 1, max(first, 4)))
    idx = np.stack([bidx, np.arange(M, dtype=np.int64)], 1)
    inst = SparseBoxList(T(idx, cuda), BoxList(T(boxes, cuda)), (N, max(M, 1)))
    got = pooler([T(f, cuda) for f in feats], inst).cpu().numpy()
    if sr <= 1:
        assert np.array_equal(got, want), seed
    else:  # avg-pool summation order unspecified in TF; here it is the same order -> still exact
        assert np.allclose(got, want, rtol=1e-5, atol=1e-6), seed
        assert np.array_equal(got, want), seed


@pytest.mark.parametrize("seed", range(30))
def test_fuzz_rpn_and_fast_rcnn(cuda, oracle_lib, seed):
    rng = np.random.default_rng(4000 + seed)
    N = int(rng.integers(1, 4))
    L = int(rng.integers(1, 5))
    hwa = [int(rng.choice([3, 50, 700, 2500, 9000])) for _ in range(L)]
    shapes = np.stack([rng.integers(100, 400, N), rng.integers(100, 500, N)], 1).astype(np.int32)
    props = [np.stack([fuzz_boxes(rng, n, 400, 500) for _ in range(N)]) for n in hwa]
    logits = [fuzz_scores(rng, (N, n)) for n in hwa]
    logits = [np.where(np.isinf(x), np.float32(0.0), x) for x in logits]
    pre, post = int(rng.choice([10, 300, 2000])), int(rng.choice([5, 100, 1000]))
    msl = float(rng.choice([0.0, 0.0, 8.0]))
    thr = float(rng.choice([0.5, 0.7]))
    wb, wl, wv, _ = oracle_lib.find_top_rpn_proposals(props, logits, shapes, thr, pre, post, msl)
    res = find_top_rpn_proposals([T(p, cuda) for p in props], [T(x, cuda) for x in logits],
                                 ImageList(None, T(shapes, cuda)), thr, pre, post, msl)
    assert np.array_equal(res.get_field("is_valid").cpu().numpy(), wv), seed
    assert np.array_equal(res.boxes.cpu().numpy(), wb), seed
    assert np.array_equal(res.get_field("objectness_logits").cpu().numpy(), wl), seed
    # Fast R-CNN post-processing on ragged ROIs
    R, K = int(rng.choice([1, 33, 200])), int(rng.choice([1, 3, 20]))
    keep = rng.random(N * R) < 0.85
    idx = np.stack([np.repeat(np.arange(N), R), np.tile(np.arange(R), N)], 1).astype(np.int64)[keep]
    M = idx.shape[0]
    agn = bool(rng.random() < 0.3)
    Kb = 1 if agn else K
    boxes = fuzz_boxes(rng, M * Kb, 400, 500).reshape(M, Kb * 4)
    sc = rng.random((M, K + 1)).astype(np.float32) ** 3
    sc[:, :K] = np.round(sc[:, :K] * 16) / 16 if rng.random() < 0.5 else sc[:, :K]
    topk = int(rng.choice([1, 10, 100]))
    want = oracle_lib.fast_rcnn_inference(boxes, sc, idx, (N, R), shapes, 0.05, 0.5, topk, agn)
    proposals = SparseBoxList(T(idx, cuda), BoxList(T(np.zeros((M, 4), np.float32), cuda)), (N, R))
    proposals.set_tracking("image_shape", T(shapes, cuda))
    got, _ = fast_rcnn_inference(T(boxes, cuda), T(sc, cuda), proposals, 0.05, 0.5, topk, agn)
    assert np.array_equal(got.get_field("is_valid").cpu().numpy(), want[3]), seed
    assert np.array_equal(got.get_field("pred_classes").cpu().numpy(), want[2]), seed
    assert np.array_equal(got.get_field("scores").cpu().numpy(), want[1]), seed
    assert np.array_equal(got.boxes.cpu().numpy(), want[0]), seed


@pytest.mark.parametrize("seed", range(30))
def test_fuzz_label_boxes_yolo_matrix_nms(cuda, oracle_lib, seed):
    rng = np.random.default_rng(5000 + seed)
    # label assignment
    N, P, G = int(rng.integers(1, 4)), int(rng.choice([1, 255, 257, 3000])), int(rng.choice([0, 1, 7, 40]))
    shared = bool(rng.random() < 0.5)
    pred = fuzz_boxes(rng, P, 300, 400) if shared else np.stack([fuzz_boxes(rng, P, 300, 400) for _ in range(N)])
    gt = np.stack([fuzz_boxes(rng, G, 300, 400) for _ in range(N)]) if G else np.zeros((N, 0, 4), np.float32)
    if G and P > 4:
        gt[:, 0] = (pred if shared else pred[0])[:1]  # exact match somewhere
    valid, crowd, diff = rng.random((N, G)) < 0.7, rng.random((N, G)) < 0.2, rng.random((N, G)) < 0.1
    counts = None if shared else rng.integers(0, P + 1, N).astype(np.int32)
    th, lab = ([0.3, 0.7], [0, -1, 1]) if rng.random() < 0.5 else ([0.5], [0, 1])
    lq = bool(rng.random() < 0.5)
    shapes = np.stack([rng.integers(100, 300, N), rng.integers(100, 400, N)], 1).astype(np.int32)
    bthr = float(rng.choice([-1.0, 0.0, 5.0]))
    w = (10., 10., 5., 5.)
    wm, wl, wd = oracle_lib.label_boxes(pred, gt, valid, th, lab, lq, gt_crowd=crowd, gt_difficult=diff, pred_counts=counts,
                                        boundary_threshold=bthr, image_shapes=shapes, weights=w)
    from detectron2_tensorflow_b200.modeling import Box2BoxTransform
    gm, gl, gd = label_boxes(T(pred, cuda), T(gt, cuda), T(valid, cuda), Matcher(th, lab, lq), gt_crowd=T(crowd, cuda),
                             gt_difficult=T(diff, cuda), pred_counts=None if counts is None else T(counts, cuda),
                             boundary_threshold=bthr, image_shapes=T(shapes, cuda), box2box_transform=Box2BoxTransform(w))
    assert np.array_equal(gl.cpu().numpy(), wl), seed
    assert np.array_equal(gm.cpu().numpy(), wm), seed
    assert np.array_equal(gd.cpu().numpy(), wd, equal_nan=True), seed
    # YOLO post-processing
    n, K, topk = int(rng.choice([0, 1, 100, 5000])), int(rng.choice([1, 3, 33, 80])), int(rng.choice([1, 20, 200]))
    yb = np.stack([fuzz_boxes(rng, n, 300, 400) for _ in range(N)]) if n else np.zeros((N, 0, 4), np.float32)
    yp = (rng.random((N, n, K)) ** 3).astype(np.float32)
    if rng.random() < 0.5:
        yp = np.round(yp * 8) / 8
    want = oracle_lib.yolo_inference(yb, yp.astype(np.float32), 0.2, 0.5, topk)
    res = YOLOv4Inference(0.2, 0.5, topk).inference(T(yb, cuda), T(yp.astype(np.float32), cuda))
    assert np.array_equal(res.get_field("is_valid").cpu().numpy(), want[3]), seed
    assert np.array_equal(res.get_field("pred_classes").cpu().numpy(), want[2]), seed
    assert np.array_equal(res.get_field("scores").cpu().numpy(), want[1]), seed
    assert np.array_equal(res.boxes.cpu().numpy(), want[0]), seed
    # Matrix-NMS on random blobs with duplicates
    nm, H, W = int(rng.choice([1, 2, 33, 100])), int(rng.integers(5, 40)), int(rng.integers(5, 50))
    m = (rng.random((nm, H, W)) < rng.uniform(0.05, 0.6)).astype(np.float32)
    m[nm // 2:] = m[:nm - nm // 2]
    m[0] = 0.0  # empty mask: 0/0 unions
    cls = rng.integers(0, 3, nm).astype(np.int64)
    sc = np.sort(rng.random(nm).astype(np.float32))[::-1].copy()
    for kern in ("gaussian", "linear"):
        want = oracle_lib.matrix_nms(m, cls, sc, None, kern, 2.0)
        got = matrix_nms(T(m, cuda), T(cls, cuda), T(sc, cuda), kernel=kern, sigma=2.0).cpu().numpy()
        assert np.array_equal(got, want, equal_nan=True), (seed, kern)


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_solo_upsample_and_select(cuda, oracle_lib, seed):
    """SOLOv2 neighbours: image-size masks + boxes (both resize conventions, up- and down-scaling, odd sizes, empty /
    full / single-pixel masks) and the ordered candidate selection (random thresholds, caps that overflow)."""
    from detectron2_tensorflow_b200.modeling import SOLOv2Inference, solo_upsample_masks
    rng = np.random.default_rng(7000 + seed)
    B, D = int(rng.integers(1, 3)), int(rng.integers(1, 6))
    h, w = int(rng.integers(1, 40)), int(rng.integers(1, 50))
    H, W = int(rng.integers(1, 150)), int(rng.integers(1, 200))
    if max(h / H, 1.0) * w * 4 * 40 > 150_000:  # keep a CTA's staged source rows inside shared memory
        H = max(H, h // 4)
    ac = bool(rng.integers(0, 2))
    thr = float(rng.choice([0.5, 0.5, 0.3, 0.7]))
    m = (rng.random((B, D, h, w)) < rng.choice([0.02, 0.3, 0.7])).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for b in range(B):
        for d in range(D):
            kind = rng.integers(0, 6)
            if kind == 0:
                m[b, d] = 0
            elif kind == 1:
                m[b, d] = 1
            elif kind == 2:
                m[b, d] = 0
                m[b, d, rng.integers(0, h), rng.integers(0, w)] = 1
            elif kind == 3:
                cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(0.5, max(h, w))
                m[b, d] = ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.float32)
    flat = m.reshape(B, D, -1).astype(np.uint8)
    flat = np.concatenate([flat, np.zeros((B, D, (-flat.shape[-1]) % 64), np.uint8)], -1)
    packed = np.packbits(flat, axis=-1, bitorder="little").view(np.int64)
    got = solo_upsample_masks(T(packed, cuda), (h, w), (H, W), thr, ac, return_masks=True, return_packed=True)
    for b in range(B):
        wm, wb = oracle_lib.solo_upsample_boxes(m[b], (H, W), ac, thr)
        assert np.array_equal(got["pred_masks"][b].cpu().numpy(), wm)
        assert np.array_equal(got["boxes"][b].cpu().numpy(), wb)
        bits = np.unpackbits(got["packed_masks"][b].cpu().numpy().view(np.uint8), axis=-1, bitorder="little")
        assert np.array_equal(bits[:, :H * W].reshape(D, H, W), wm) and not bits[:, H * W:].any()
    # candidate selection
    grids = tuple(int(g) for g in rng.integers(1, 9, rng.integers(1, 4)))
    strides = tuple(int(s_) for s_ in rng.choice([4, 8, 16, 32], len(grids)))
    K, E = int(rng.integers(1, 9)), 4 * int(rng.integers(1, 5))
    G = sum(g * g for g in grids)
    sc = rng.random((B, G, K)).astype(np.float32)
    sc[rng.random(sc.shape) < 0.2] = 0.5  # ties with the threshold (strict >)
    kn = rng.standard_normal((B, G, E)).astype(np.float32)
    sthr = float(rng.choice([0.5, 0.1, 0.9, 0.0, 1.0]))
    cap = int(rng.choice([1, 7, 64, 2048]))
    head = SOLOv2Inference(score_threshold=sthr, num_grids=grids, strides=strides, max_candidates=cap)
    sel = head.select_candidates(T(sc, cuda), T(kn, cuda))
    for b in range(B):
        ws, wc, wk, wst = oracle_lib.solo_select(sc[b], kn[b], grids, strides, sthr)
        n = int(sel["counts"][b])
        assert int(sel["total"][b]) == len(ws) and n == min(len(ws), cap)
        assert np.array_equal(sel["scores"][b, :n].cpu().numpy(), ws[:n])
        assert np.array_equal(sel["classes"][b, :n].cpu().numpy(), wc[:n])
        assert np.array_equal(sel["strides"][b, :n].cpu().numpy(), wst[:n])
        assert np.array_equal(sel["kernels"][b, :n].cpu().numpy(), wk[:n])


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_solo_dynamic_masks(cuda, oracle_lib, seed):
    """The tensor-core dynamic conv + mask stage on irregular shapes (row blocks 1..3, channels not a multiple of the K
    block, maps that end inside a pixel tile, ragged counts): logits within 1e-5 of sum|terms| of the oracle's fp32 sum,
    everything after the logits exact."""
    from detectron2_tensorflow_b200.modeling import solo_dynamic_masks
    rng = np.random.default_rng(9000 + seed)
    B = int(rng.integers(1, 4))
    n = int(rng.choice([1, 3, 127, 128, 129, 200, 256, 257, 300]))
    H, W = int(rng.integers(1, 40)), int(rng.integers(1, 60))
    E = 4 * int(rng.integers(1, 24))
    thr = float(rng.choice([0.5, 0.5, 0.4, 0.6]))
    feat = (rng.standard_normal((B, H, W, E)) * rng.choice([0.1, 1.0, 30.0])).astype(np.float32)
    kern = (rng.standard_normal((B, n, E)) * rng.choice([0.05, 1.0])).astype(np.float32)
    counts = rng.integers(0, n + 1, B).astype(np.int32)
    counts[0] = n
    packed, sm, ss, logits = solo_dynamic_masks(T(feat, cuda), T(kern, cuda), thr, T(counts, cuda), return_logits=True)
    bits = np.unpackbits(packed.cpu().numpy().view(np.uint8), axis=-1, bitorder="little")[..., :H * W]
    glog = logits.cpu().numpy().reshape(B, n, H * W)
    for b in range(B):
        c = int(counts[b])
        want, absum = oracle_lib.solo_dynamic_conv(feat[b], kern[b, :c])
        assert (np.abs(glog[b, :c].astype(np.float64) - want) <= 1e-5 * absum + 1e-30).all()
        m, wsm, wss = oracle_lib.solo_mask_stage(glog[b, :c].reshape(c, H, W), thr)
        assert np.array_equal(bits[b, :c].reshape(c, H, W), m.astype(np.uint8))
        assert np.array_equal(sm[b, :c].cpu().numpy(), wsm)
        assert np.allclose(ss[b, :c].cpu().numpy(), wss, rtol=1e-5, atol=1e-5)
        assert not bits[b, c:].any() and not sm[b, c:].any() and not ss[b, c:].any()
