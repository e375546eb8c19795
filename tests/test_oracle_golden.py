"""CPU: pin the oracle against the committed golden vectors (tests/golden/, see make_golden.py)."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_nms_matches_reference_numpy_nms(oracle_lib):
    """Outputs of the reference's own numpy NMS (lib/structures/np_box_list_ops.py:146-216)."""
    z = np.load(os.path.join(G, "nms_reference_numpy.npz"))
    for c in range(int(z["num_cases"])):
        keep = oracle_lib.nms(z[f"c{c}_boxes"], z[f"c{c}_scores"], int(z[f"c{c}_max_out"]), float(z[f"c{c}_thr"]))
        assert np.array_equal(keep, z[f"c{c}_keep"]), c


def test_nms_matches_torchvision(oracle_lib):
    z = np.load(os.path.join(G, "nms_torchvision.npz"))
    keep = oracle_lib.nms(z["boxes"], z["scores"], 500, float(z["thr"]))
    assert np.array_equal(keep, z["keep"])


def test_topk_matches_torch(oracle_lib):
    z = np.load(os.path.join(G, "topk_torch.npz"))
    v, i = oracle_lib.top_k(z["x"], 1000)
    assert np.array_equal(i, z["indices"]) and np.array_equal(v, z["values"])


@pytest.mark.parametrize("o,sr", [(7, 0), (7, 2), (14, 0)])
def test_roi_align_close_to_torchvision(oracle_lib, o, sr):
    """Independent implementation (textbook aligned ROIAlign): agreement to fp32 rounding of the
    reference's normalise/denormalise coordinate round trip, i.e. ~4e-5 abs."""
    z = np.load(os.path.join(G, "roi_align_torchvision.npz"))
    got = oracle_lib.roi_align(z["img"], z["boxes"], z["box_ind"], (o, o), float(z["scale"]), sr, True)
    assert np.abs(got - z[f"out_{o}_{sr}"]).max() < 2e-4


def test_oracle_regression_pins(oracle_lib):
    z = np.load(os.path.join(G, "oracle_pins.npz"))
    assert np.array_equal(oracle_lib.expf(z["exp_x"]), z["exp_y"])
    assert np.array_equal(oracle_lib.logf(z["log_x"]), z["log_y"], equal_nan=True)
    assert np.array_equal(oracle_lib.assign_boxes_to_levels(z["lvl_boxes"], 2, 5, 224, 4), z["lvl"])
    assert np.array_equal(oracle_lib.nms(z["tie_boxes"], z["tie_scores"], 20, 0.5), z["tie_keep"])
    v, i = oracle_lib.top_k(z["tk_x"], 64)
    assert np.array_equal(i, z["tk_i"]) and np.array_equal(v, z["tk_v"], equal_nan=True)
    fb, fs, fc, fv, fr, fn = oracle_lib.fast_rcnn_inference(z["fr_boxes"], z["fr_scores"], z["fr_idx"], (2, 40),
                                                            z["fr_shapes"], 0.05, 0.5, 20, False)
    assert np.array_equal(fb, z["fr_ob"]) and np.array_equal(fs, z["fr_os"]) and np.array_equal(fc, z["fr_oc"])
    assert np.array_equal(fv, z["fr_ov"]) and np.array_equal(fr, z["fr_or"]) and np.array_equal(fn, z["fr_on"])
    shp = tuple(z["mn_shape"])
    m = np.unpackbits(z["mn_masks"])[:int(np.prod(shp))].reshape(shp).astype(np.float32)
    assert np.array_equal(oracle_lib.matrix_nms(m, z["mn_classes"], z["mn_scores"], None, "gaussian", 2.0), z["mn_gauss"])
    assert np.array_equal(oracle_lib.matrix_nms(m, z["mn_classes"], z["mn_scores"], None, "linear", 2.0), z["mn_linear"],
                          equal_nan=True)


def test_math_accuracy(oracle_lib):
    """The shared Cephes exp/log are accurate to ~1 ulp against float64 libm."""
    rng = np.random.default_rng(0)
    x = rng.uniform(-20, 20, 4000).astype(np.float32)
    assert np.max(np.abs(oracle_lib.expf(x) - np.exp(x.astype(np.float64))) / np.exp(x.astype(np.float64))) < 2e-7
    y = np.exp(rng.uniform(-20, 20, 4000)).astype(np.float32)
    assert np.max(np.abs(oracle_lib.logf(y) - np.log(y.astype(np.float64)))) < 2e-6


def test_oracle_properties(oracle_lib):
    rng = np.random.default_rng(4)
    # level assignment is monotone in area
    s = np.sort(np.exp(rng.uniform(0, 8, 2000))).astype(np.float32)
    lv = oracle_lib.assign_boxes_to_levels(np.stack([0 * s, 0 * s, s, s], 1), 2, 5)
    assert np.all(np.diff(lv) >= 0) and lv.min() == 0 and lv.max() == 3
    # ROIAlign of a constant map is that constant inside [-1, H], zero far outside
    img = np.full((1, 12, 12, 4), 2.5, np.float32)
    out = oracle_lib.roi_align(img, np.array([[1, 1, 9, 10], [-300, -300, -200, -200]], np.float32),
                               np.zeros(2, np.int32), (3, 3), 1.0, 2, True)
    assert np.all(out[0] == 2.5) and np.all(out[1] == 0)
    # NMS idempotence and the cap
    b = rng.uniform(0, 200, (300, 2)).astype(np.float32)
    boxes = np.concatenate([b, b + rng.uniform(20, 80, (300, 2)).astype(np.float32)], 1)
    sc = rng.permutation(300).astype(np.float32)
    k = oracle_lib.nms(boxes, sc, 300, 0.5)
    k2 = oracle_lib.nms(boxes[k], sc[k], 300, 0.5)
    assert len(k2) == len(k)
    assert np.array_equal(oracle_lib.nms(boxes, sc, 7, 0.5), k[:7])
    # empty inputs
    assert len(oracle_lib.nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 5, 0.5)) == 0
    v, i = oracle_lib.top_k(np.zeros(0, np.float32), 3)
    assert len(v) == 0
