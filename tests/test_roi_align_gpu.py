"""GPU parity: d2b_roi_align_multilevel (through the host layer / C-ABI) vs the CPU oracle.

Tolerance (BASELINE.json north_star): ROIAlign features within 1e-5 relative error.  The kernel
follows the oracle's fp32 op order with -fmad=false, so we additionally expect bit equality.
"""
import numpy as np
import pytest
import torch

from detectron2_tensorflow_b200.layers import ROIAlign, crop_and_resize
from detectron2_tensorflow_b200.modeling import ROIPooler, assign_boxes_to_levels
from detectron2_tensorflow_b200.structures import BoxList, SparseBoxList
from detectron2_tensorflow_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _instances(boxes, indices, dev):
    return SparseBoxList(torch.from_numpy(indices).to(dev), BoxList(torch.from_numpy(boxes).to(dev)),
                         (int(indices[:, 0].max()) + 1, int(indices[:, 1].max()) + 1))


@pytest.mark.parametrize("out,sr,ptype", [(7, 0, "ROIAlignV2"), (14, 0, "ROIAlignV2"), (7, 2, "ROIAlignV2"),
                                          (7, 0, "ROIAlign"), (5, 3, "ROIAlign")])
def test_pooler_multilevel_vs_oracle(cuda, oracle_lib, out, sr, ptype):
    N, R, C = 2, 150, 32
    feats = syn.fpn_features(N, C, seed=0, padded_hw=(256, 320))
    boxes, idx = syn.rois(N, R, seed=1, image_hw=(250, 310))
    scales = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    want, want_counts = oracle_lib.roi_pooler(feats, scales, boxes, idx[:, 0], (out, out), sr,
                                              aligned=(ptype == "ROIAlignV2"))
    pooler = ROIPooler(out, scales, sr, ptype)
    got = pooler([torch.from_numpy(f).to(cuda) for f in feats], _instances(boxes, idx, cuda))
    got = got.cpu().numpy()
    assert got.shape == want.shape
    assert _rel_err(got, want) <= RTOL
    assert np.array_equal(got, want), "expected bit-exact (same op order, no FMA)"
    assert np.array_equal(pooler.last_level_counts.cpu().numpy(), want_counts)


def test_level_assignment_bit_exact(cuda, oracle_lib):
    rng = np.random.default_rng(3)
    s = np.exp(rng.uniform(np.log(1), np.log(2000), 20000))
    a = np.exp(rng.uniform(np.log(0.25), np.log(4), 20000))
    b = np.stack([np.zeros_like(s), np.zeros_like(s), s * np.sqrt(a), s / np.sqrt(a)], 1).astype(np.float32)
    # exact bin edges and degenerate boxes
    edges = np.array([[0, 0, 112, 112], [0, 0, 224, 224], [0, 0, 448, 448], [0, 0, 0, 0], [5, 5, 3, 9],
                      [0, 0, 111.99999, 112], [0, 0, 224.00002, 224]], np.float32)
    b = np.concatenate([b, edges])
    want = oracle_lib.assign_boxes_to_levels(b, 2, 5, 224, 4)
    got = assign_boxes_to_levels(BoxList(torch.from_numpy(b).to(cuda)), 2, 5, 224, 4).cpu().numpy()
    assert got.dtype == np.int64
    assert np.array_equal(got, want)


def test_single_level_roi_align_and_crop(cuda, oracle_lib):
    rng = np.random.default_rng(5)
    img = rng.standard_normal((3, 40, 56, 24)).astype(np.float32)
    boxes, idx = syn.rois(3, 64, seed=9, image_hw=(640, 896))
    bi = idx[:, 0].astype(np.int32)
    want = oracle_lib.roi_align(img, boxes, bi, (7, 7), 1 / 16., 2, True)
    got = ROIAlign((7, 7), 1 / 16., 2, True)(torch.from_numpy(img).to(cuda), torch.from_numpy(boxes).to(cuda),
                                             torch.from_numpy(bi).to(cuda)).cpu().numpy()
    assert np.array_equal(got, want)
    # crop_and_resize, both pad_border modes, feature-map coordinates
    fb = boxes / 16.0
    for pad in (True, False):
        for aligned in (True, False):
            want = oracle_lib.crop_and_resize(img, fb, bi, (9, 5), aligned, pad)
            got = crop_and_resize(torch.from_numpy(img).to(cuda), torch.from_numpy(fb).to(cuda),
                                  torch.from_numpy(bi).to(cuda), (9, 5), aligned, pad_border=pad).cpu().numpy()
            assert np.array_equal(got, want), (pad, aligned)


def test_large_output_more_samples_than_cta_threads(cuda, oracle_lib):
    """64x64 output with sampling_ratio 2: 128 + 128 crop taps per ROI, more than the CTA has threads (the tap
    prologue loops), and 4,096 bins split over many CTAs."""
    rng = np.random.default_rng(15)
    img = rng.standard_normal((2, 30, 44, 8)).astype(np.float32)
    boxes, idx = syn.rois(2, 6, seed=4, image_hw=(480, 704))
    bi = idx[:, 0].astype(np.int32)
    want = oracle_lib.roi_align(img, boxes, bi, (64, 64), 1 / 16., 2, True)
    got = ROIAlign((64, 64), 1 / 16., 2, True)(torch.from_numpy(img).to(cuda), torch.from_numpy(boxes).to(cuda),
                                               torch.from_numpy(bi).to(cuda)).cpu().numpy()
    assert np.array_equal(got, want)


def test_edge_cases(cuda, oracle_lib):
    rng = np.random.default_rng(6)
    img = rng.standard_normal((2, 20, 30, 8)).astype(np.float32)
    t = torch.from_numpy(img).to(cuda)
    # empty ROI set
    out = ROIAlign((7, 7), 0.25, 0, True)(t, torch.zeros((0, 4), device=cuda), torch.zeros(0, dtype=torch.int32, device=cuda))
    assert tuple(out.shape) == (0, 7, 7, 8)
    # boxes far outside, zero-area, inverted, 1x1 output, batch index out of range
    boxes = np.array([[-500, -500, -400, -400], [10, 10, 10, 10], [50, 60, 20, 30], [0, 0, 80, 120],
                      [79, 119, 200, 300], [4, 4, 12, 12]], np.float32)
    bi = np.array([0, 1, 0, 1, 0, 7], np.int32)
    for osz in ((1, 1), (7, 7), (2, 3)):
        want = oracle_lib.roi_align(img, boxes, bi, osz, 0.25, 0, True)
        got = ROIAlign(osz, 0.25, 0, True)(t, torch.from_numpy(boxes).to(cuda), torch.from_numpy(bi).to(cuda)).cpu().numpy()
        assert np.array_equal(got, want), osz
    # constant map stays constant inside [-1, H] and is zero outside
    const = torch.full((1, 16, 16, 4), 3.5, device=cuda)
    got = ROIAlign((4, 4), 1.0, 2, True)(const, torch.tensor([[2., 2., 9., 11.]], device=cuda),
                                         torch.zeros(1, dtype=torch.int32, device=cuda))
    assert torch.all(got == 3.5)
    with pytest.raises(ValueError):
        ROIPooler(7, [0.25], 0, "ROIPool")


def test_host_buffers_roundtrip(cuda, oracle_lib):
    """Host tensors in -> host tensor out (H2D/D2H inside the operator), same values."""
    N, R, C = 1, 40, 16
    feats = syn.fpn_features(N, C, seed=2, padded_hw=(128, 160))
    boxes, idx = syn.rois(N, R, seed=4, image_hw=(120, 150))
    scales = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    want, _ = oracle_lib.roi_pooler(feats, scales, boxes, idx[:, 0], (7, 7), 0)
    inst = SparseBoxList(torch.from_numpy(idx), BoxList(torch.from_numpy(boxes)), (N, R))
    got = ROIPooler(7, scales, 0, "ROIAlignV2")([torch.from_numpy(f) for f in feats], inst)
    assert not got.is_cuda
    assert np.array_equal(got.numpy(), want)


def test_bf16_features(cuda, oracle_lib):
    N, R, C = 2, 60, 64
    feats = syn.fpn_features(N, C, seed=0, padded_hw=(128, 192))
    boxes, idx = syn.rois(N, R, seed=1, image_hw=(120, 190))
    scales = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    tb = [torch.from_numpy(f).to(cuda).to(torch.bfloat16) for f in feats]
    ref_in = [t.float().cpu().numpy() for t in tb]  # oracle on the bf16-rounded values
    want, _ = oracle_lib.roi_pooler(ref_in, scales, boxes, idx[:, 0], (7, 7), 2)
    got = ROIPooler(7, scales, 2, "ROIAlignV2")(tb, _instances(boxes, idx, cuda))
    assert got.dtype == torch.bfloat16
    want_bf = torch.from_numpy(want).to(torch.bfloat16).float().numpy()
    assert np.array_equal(got.float().cpu().numpy(), want_bf)
