"""GPU: memory-safety evidence without compute-sanitizer (closed on this pool).  Every device buffer the host layer
hands to libd2b200 as an OUTPUT or as WORKSPACE is placed between two 4 KB canary bands filled with 0xA5 -- outputs
by intercepting `torch.empty` / `torch.zeros` for CUDA tensors, workspaces by replacing `_native._workspace` with an
allocator that returns EXACTLY the bytes `d2b_<op>_workspace_bytes` asked for -- and the differential tests of the
other GPU files are re-run on irregular shapes (n not a multiple of 64, edge words, empty inputs, candidate-list
overflow, both proposal-stage paths).  Afterwards every band must be untouched: a kernel that writes one byte before
or after any output or scratch buffer fails here even when the oracle comparison still passes."""
import numpy as np
import pytest
import torch

from detectron2_tensorflow_b200 import _native as nv

import test_fuzz_gpu as fz
import test_paste_masks_gpu as pm
import test_postprocess_gpu as pp
import test_solo_inference_gpu as si
import test_training_side_gpu as ts

pytestmark = pytest.mark.gpu

BAND = 4096
PATTERN = 0xA5


class Guard(object):
    def __init__(self, monkeypatch):
        self.records = []
        self._empty, self._zeros = torch.empty, torch.zeros
        monkeypatch.setattr(torch, "empty", self.empty)
        monkeypatch.setattr(torch, "zeros", self.zeros)
        monkeypatch.setattr(nv, "_workspace", self.workspace)

    @staticmethod
    def _is_cuda(device):
        if device is None:
            return False
        return torch.device(device).type == "cuda"

    def _guarded(self, nbytes, device):
        buf = self._empty(2 * BAND + nbytes, dtype=torch.uint8, device=device)
        buf.fill_(PATTERN)
        self.records.append((buf, nbytes))
        return buf[BAND:BAND + nbytes]

    def empty(self, *size, **kw):
        if not self._is_cuda(kw.get("device")) or kw.get("pin_memory"):
            return self._empty(*size, **kw)
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        dtype = kw.get("dtype") or torch.get_default_dtype()
        n = int(np.prod(shape)) if len(shape) else 1
        item = self._empty(0, dtype=dtype).element_size()
        return self._guarded(n * item, kw["device"]).view(dtype).reshape(shape)

    def zeros(self, *size, **kw):
        if not self._is_cuda(kw.get("device")):
            return self._zeros(*size, **kw)
        return self.empty(*size, **kw).zero_()

    def workspace(self, nbytes, device):
        return self._guarded(max(int(nbytes), 256) if nbytes == 0 else int(nbytes), device)

    def check(self):
        torch.cuda.synchronize()
        assert self.records, "nothing was allocated through the guard"
        for buf, nbytes in self.records:
            lo, hi = buf[:BAND], buf[BAND + nbytes:]
            assert bool((lo == PATTERN).all()) and bool((hi == PATTERN).all()), \
                f"canary band damaged around a {nbytes}-byte buffer"
        n = len(self.records)
        self.records = []
        return n


@pytest.fixture
def guard(monkeypatch):
    return Guard(monkeypatch)


@pytest.mark.parametrize("seed", range(16))
def test_canary_batched_nms_topk(cuda, oracle_lib, guard, seed):
    fz.test_fuzz_batched_nms(cuda, oracle_lib, seed)
    fz.test_fuzz_topk(cuda, oracle_lib, seed)
    assert guard.check() >= 4


@pytest.mark.parametrize("seed", range(12))
def test_canary_roi_pooler_rpn_fast_rcnn(cuda, oracle_lib, guard, seed):
    fz.test_fuzz_roi_pooler(cuda, oracle_lib, seed)
    fz.test_fuzz_rpn_and_fast_rcnn(cuda, oracle_lib, seed)
    guard.check()


@pytest.mark.parametrize("seed", range(8))
def test_canary_labels_yolo_matrix_nms_solo(cuda, oracle_lib, guard, seed):
    fz.test_fuzz_label_boxes_yolo_matrix_nms(cuda, oracle_lib, seed)
    fz.test_fuzz_solo_upsample_and_select(cuda, oracle_lib, seed)
    fz.test_fuzz_solo_dynamic_masks(cuda, oracle_lib, seed)
    guard.check()


def test_canary_proposal_stage_both_paths(cuda, oracle_lib, guard, monkeypatch):
    for variant, pre, post, min_len in (("gaussian", 300, 200, 0.0), ("ties", 200, 150, 0.0), ("clustered", 2000, 1000, 4.0)):
        pp.test_rpn_proposals(cuda, oracle_lib, variant, pre, post, min_len)
    pp.test_rpn_proposals_large_k_uses_generic_chain(cuda, oracle_lib)
    pp.test_rpn_full_size_one_image(cuda, oracle_lib)
    monkeypatch.setenv("D2B_RPN_GENERIC", "1")
    pp.test_rpn_proposals(cuda, oracle_lib, "clustered", 2000, 1000, 4.0)
    guard.check()


def test_canary_detection_heads(cuda, oracle_lib, guard):
    pp.test_fast_rcnn_inference(cuda, oracle_lib, False, False, 120, 20, 100)
    pp.test_retinanet_inference(cuda, oracle_lib)
    pp.test_yolo_postprocess(cuda, oracle_lib, 3000, 20, 300)
    pp.test_matrix_nms(cuda, oracle_lib, "gaussian")
    pp.test_solo_mask_encode_and_packed_matrix_nms(cuda, oracle_lib, (25, 37), 0.5)
    pp.test_solo_postprocess(cuda, oracle_lib, 77, (25, 37), 500, 100, "linear")
    for case in ("typical", "plateau_overflow", "tied_boundary"):
        pp.test_sigmoid_topk_paths(cuda, oracle_lib, case)
    guard.check()


def test_canary_masks_and_training_side(cuda, oracle_lib, guard):
    pm.test_paste_masks_small(cuda, oracle_lib, 50, 75, 7, 9)
    pm.test_mask_rcnn_inference(cuda, oracle_lib)
    ts.test_label_boxes_rpn(cuda, oracle_lib, True, 0)
    ts.test_decode_clip_filter(cuda, oracle_lib)
    ts.test_crop_and_resize_entry(cuda, oracle_lib, True, True)
    ts.test_roi_align_backward_single(cuda, oracle_lib, 2, True)
    si.test_select_candidates(cuda, oracle_lib, (6, 4), 3, 8, 0.3, 64)
    guard.check()


def test_guard_detects_an_overrun(cuda, guard):
    """The harness itself: one byte written past an output must be reported."""
    t = torch.empty(10, dtype=torch.float32, device=cuda)
    buf, nbytes = guard.records[-1]
    buf[BAND + nbytes] = 0
    with pytest.raises(AssertionError):
        guard.check()
