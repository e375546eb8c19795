"""CPU: the TensorFlow custom-op shim (csrc/tf_ops/d2b_tf_ops.cc) type-checks against include/d2b200.h and a minimal
restatement of the TF C++ op API (tests/tf_stub/): `g++ -std=c++14 -fsyntax-only`.  TensorFlow itself cannot be
installed here, so this is the strongest check available: names, argument types, attr getters, the params structs and
every C-ABI call of the shim are verified by a compiler; TensorFlow's runtime behaviour is not."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "detectron2_tensorflow_b200", "csrc", "tf_ops", "d2b_tf_ops.cc")


def _gxx():
    for c in ("/usr/bin/g++", shutil.which("g++")):
        if c and os.path.exists(c):
            return c
    return None


def test_tf_shim_type_checks_against_stub_headers():
    gxx = _gxx()
    if gxx is None:
        pytest.skip("no g++")
    cmd = [gxx, "-std=c++14", "-fsyntax-only", "-Wall", "-Wextra", "-Werror=return-type",
           "-I", os.path.join(ROOT, "tests", "tf_stub"), "-I", os.path.join(ROOT, "include")]
    if os.path.isdir("/usr/local/cuda/include"):
        cmd += ["-isystem", "/usr/local/cuda/include"]
    r = subprocess.run(cmd + [SHIM], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-4000:]


def test_tf_shim_registers_every_north_star_op():
    """One REGISTER_OP + REGISTER_KERNEL_BUILDER per op of SURVEY.md 8(b), each calling its C-ABI entry point."""
    src = open(SHIM).read()
    ops = set(re.findall(r'REGISTER_OP\("(\w+)"\)', src))
    kernels = set(re.findall(r'REGISTER_KERNEL_BUILDER\(Name\("(\w+)"\)', src))
    want = {"D2RoiAlignMultilevel": "d2b_roi_align_multilevel", "D2BatchedNms": "d2b_batched_nms",
            "D2RpnProposals": "d2b_rpn_proposals", "D2FastRcnnPostprocess": "d2b_fast_rcnn_postprocess",
            "D2RetinanetPostprocess": "d2b_retinanet_postprocess", "D2CropAndResizeAligned": "d2b_crop_and_resize_aligned",
            "D2MatrixNms": "d2b_matrix_nms", "D2RoiAlignMultilevelGrad": "d2b_roi_align_backward"}
    for op, entry in want.items():
        assert op in ops and op in kernels, op
        assert re.search(r"RunOp\(ctx, p, %s_workspace_bytes, %s\)" % (entry, entry), src), entry
    assert ops == kernels
    # int64 classes for Fast R-CNN, int32 for RetinaNet (fast_rcnn.py:178, retinanet.py:381)
    assert re.search(r'REGISTER_OP\("D2FastRcnnPostprocess"\)[^;]*?Output\("pred_classes: int64"\)', src, re.S)
    assert re.search(r'REGISTER_OP\("D2RetinanetPostprocess"\)[^;]*?Output\("pred_classes: int32"\)', src, re.S)
