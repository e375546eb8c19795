"""GPU parity of the full SOLOv2 inference path (`SOLOv2Inference.inference`, the drop-in for
`MaskKernelBranch.inference`, solo_v2.py:476-627): candidate selection (`d2b_solo_select`), dynamic conv + mask stage,
tail, image-size masks + boxes -- against the oracle and the reference-python golden, from the RAW head outputs."""
import os

import numpy as np
import pytest
import torch

from detectron2_tensorflow_b200.modeling import SOLOv2Inference

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.mark.parametrize("grids,K,E,thr,cap", [((6, 4), 3, 8, 0.3, 64), ((40, 36, 24, 16, 12), 80, 32, 0.97, 2048),
                                                ((5,), 2, 4, 0.0, 64), ((5,), 2, 4, 2.0, 8), ((12, 8), 7, 16, 0.5, 100)])
def test_select_candidates(cuda, oracle_lib, grids, K, E, thr, cap):
    rng = np.random.default_rng(len(grids) * 10 + K)
    B = 3
    G = sum(g * g for g in grids)
    strides = tuple(8 * (i + 1) for i in range(len(grids)))
    sc = rng.random((B, G, K)).astype(np.float32)
    sc[2] = 0  # an image without candidates
    kn = rng.standard_normal((B, G, E)).astype(np.float32)
    head = SOLOv2Inference(score_threshold=thr, num_grids=grids, strides=strides, max_candidates=cap)
    got = head.select_candidates(T(sc, cuda), T(kn, cuda))
    for b in range(B):
        ws, wc, wk, wst = oracle_lib.solo_select(sc[b], kn[b], grids, strides, thr)
        n = int(got["counts"][b])
        assert int(got["total"][b]) == len(ws) and n == min(len(ws), cap)
        assert np.array_equal(got["scores"][b, :n].cpu().numpy(), ws[:n])
        assert np.array_equal(got["classes"][b, :n].cpu().numpy(), wc[:n])
        assert np.array_equal(got["strides"][b, :n].cpu().numpy(), wst[:n])
        assert np.array_equal(got["kernels"][b, :n].cpu().numpy(), wk[:n])
        assert not got["scores"][b, n:].any() and not got["kernels"][b, n:].any()


def test_inference_matches_reference_python_golden(cuda):
    """The reference's `MaskKernelBranch.inference` (executed unmodified on the numpy TF shim, golden case 14/15) from
    its own arguments: classes / validity / image masks / boxes exact, scores 1e-5 (sigmoid + summation order)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_python.npz"))
    head = SOLOv2Inference(0.5, 30, "gaussian", 2.0, 0.05, 12, score_threshold=0.3, num_grids=(6, 4), strides=(8, 16),
                           max_candidates=64)
    probs = [T(z[f"so_raw_probs_{i}"], cuda) for i in range(2)]
    kerns = [T(z[f"so_raw_kernels_{i}"], cuda) for i in range(2)]
    got = head.inference(probs, kerns, T(z["so_in_mask_features"], cuda), tuple(int(v) for v in z["so2_image_shape"]))
    assert np.array_equal(got["num_candidates"].cpu().numpy(), z["so_in_counts"])
    assert np.array_equal(got["is_valid"].cpu().numpy(), z["so_valid"])
    assert np.array_equal(got["pred_classes"].cpu().numpy(), z["so_classes"])
    assert np.allclose(got["scores"].cpu().numpy(), z["so_scores"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(got["pred_masks"].cpu().numpy(), z["so2_masks"])
    assert np.array_equal(got["boxes"].cpu().numpy(), z["so2_boxes"])


def test_inference_vs_oracle_composition(cuda, oracle_lib):
    """Random head outputs at a mid size: the CUDA path equals the oracle pieces chained on the GPU's own logits
    (the conv is the one tolerance-based step; everything after it is exact)."""
    from detectron2_tensorflow_b200.modeling import solo_dynamic_masks
    rng = np.random.default_rng(21)
    grids, strides, K, E, (H, W), (IH, IW) = (8, 6), (8, 16), 4, 32, (28, 40), (111, 158)
    B = 2
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    feat = np.stack([np.stack([np.sin(yy * rng.uniform(0.05, 0.4) + xx * rng.uniform(0.05, 0.4) + rng.uniform(0, 6))
                               for _ in range(E)], -1) for _ in range(B)]).astype(np.float32)
    probs = [rng.random((B, g, g, K)).astype(np.float32) ** 3 for g in grids]
    kerns = [(rng.standard_normal((B, g, g, E)) * 0.6).astype(np.float32) for g in grids]
    head = SOLOv2Inference(0.5, 40, "gaussian", 2.0, 0.05, 15, score_threshold=0.2, num_grids=grids, strides=strides,
                           max_candidates=256)
    got = head.inference([T(p, cuda) for p in probs], [T(k, cuda) for k in kerns], T(feat, cuda), (IH, IW))
    sc = np.concatenate([p.reshape(B, -1, K) for p in probs], 1)
    kn = np.concatenate([k.reshape(B, -1, E) for k in kerns], 1)
    for b in range(B):
        ws, wc, wk, wst = oracle_lib.solo_select(sc[b], kn[b], grids, strides, 0.2)
        assert int(got["num_candidates"][b]) == len(ws) > 0
        logits = solo_dynamic_masks(T(feat[b:b + 1], cuda), T(wk[None], cuda), return_logits=True)[3][0].cpu().numpy()
        m, oc, os_, ov, n = oracle_lib.solo_postprocess(logits, ws, wc, wst, 0.5, 40, "gaussian", 2.0, 0.05, 15)
        assert int(got["num"][b]) == n
        assert np.array_equal(got["is_valid"][b].cpu().numpy(), ov)
        assert np.array_equal(got["pred_classes"][b].cpu().numpy(), oc)
        assert np.allclose(got["scores"][b].cpu().numpy(), os_, rtol=1e-5, atol=1e-7)
        wm, wb = oracle_lib.solo_upsample_boxes(m, (IH, IW), False, 0.5)
        assert np.array_equal(got["pred_masks"][b].cpu().numpy(), wm)
        assert np.array_equal(got["boxes"][b].cpu().numpy(), wb)
    assert int(got["num"].sum()) > 0
