"""CPU: host-side containers and synthetic generators (no GPU, no oracle)."""
import numpy as np
import pytest
import torch

from detectron2_tensorflow_b200.structures import BoxList, SparseBoxList
from detectron2_tensorflow_b200.utils import synthetic as syn


def test_boxlist_contract():
    b = BoxList(torch.zeros(3, 4))
    b.add_field("scores", torch.ones(3))
    assert b.has_field("scores") and b.num_boxes() == 3 and set(b.get_all_fields()) == {"boxes", "scores"}
    with pytest.raises(ValueError):
        BoxList(torch.zeros(3, 5))
    with pytest.raises(ValueError):
        BoxList(torch.zeros(3, 4, dtype=torch.float64))
    with pytest.raises(ValueError):
        b.get_field("nope")


def test_sparse_dense_roundtrip_row_major():
    """from_dense keeps valid rows in row-major (tf.where) order; to_dense zero-pads (box_list.py:204-264)."""
    boxes = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)
    valid = torch.tensor([[True, False, True], [False, True, True]])
    d = BoxList(boxes)
    d.add_field("is_valid", valid)
    d.add_field("scores", torch.arange(6, dtype=torch.float32).reshape(2, 3))
    d.set_tracking("image_shape", torch.tensor([[8, 9], [8, 9]]))
    s = SparseBoxList.from_dense(d)
    assert s.indices.tolist() == [[0, 0], [0, 2], [1, 1], [1, 2]]
    assert s.data.get_field("scores").tolist() == [0., 2., 4., 5.]
    back = s.to_dense()
    assert torch.equal(back.get_field("is_valid"), valid)
    assert torch.equal(back.boxes[valid], boxes[valid]) and torch.all(back.boxes[~valid] == 0)
    assert back.has_tracking("image_shape")


def test_synthetic_shapes_match_survey_appendix_b():
    assert [syn.level_hw(s) for s in syn.RPN_STRIDES] == [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    anchors = syn.rpn_anchors()
    assert [a.shape[0] for a in anchors] == [201600, 50400, 12600, 3150, 819]
    assert sum(a.shape[0] for a in anchors) == 268569
    assert sum(a.shape[0] for a in syn.retinanet_anchors()) == 201600
    # cell anchors: ratio = h / w, area = size^2 (anchor_generator.py:131-144)
    c = syn.cell_anchors([32], (0.5, 1.0, 2.0))
    h, w = c[:, 2] - c[:, 0], c[:, 3] - c[:, 1]
    assert np.allclose(h * w, 32 * 32, rtol=1e-5) and np.allclose(h / w, [0.5, 1.0, 2.0], rtol=1e-5)
    # flattened order is (y, x, a)
    a = anchors[4].reshape(13, 21, 3, 4)
    assert np.allclose(a[2, 5, 1, :2] + a[2, 5, 1, 2:], [2 * 2 * 64, 2 * 5 * 64])
    b, idx = syn.rois(2, 10)
    assert b.shape == (20, 4) and idx[:, 0].tolist() == [0] * 10 + [1] * 10
    assert np.all(b[:, 2] >= b[:, 0]) and np.all(b[:, 3] >= b[:, 1])


def test_peer_gather_segment_tables():
    """sharding.pack_segments / unpack_segments (the copy tables of PeerGatherPlan's two kernels), executed here
    with memmove on host arrays: every rank's blocks -> its receive slot -> the full-batch tensors == concatenation
    in image order, for even and uneven partitions and 1..3 image blocks per rank."""
    import ctypes
    import numpy as np
    import torch
    from detectron2_tensorflow_b200 import sharding
    spec = {"boxes": ((5, 4), torch.float32), "valid": ((5,), torch.bool), "classes": ((3,), torch.int64),
            "odd": ((7,), torch.uint8)}
    np_dtype = {torch.float32: np.float32, torch.bool: np.bool_, torch.int64: np.int64, torch.uint8: np.uint8}
    rng = np.random.default_rng(0)
    for n_images, world, chunks in ((16, 8, 1), (16, 4, 2), (7, 3, 2), (5, 4, 3), (3, 4, 1)):
        layout = sharding.block_layout(n_images, world, lambda k: chunks)
        offsets, nbytes = sharding.gather_offsets(spec, layout)
        full = {k: rng.integers(0, 255, size=(n_images,) + shape).astype(np_dtype[dt]) for k, (shape, dt) in spec.items()}
        slots = [np.zeros(nbytes, np.uint8) for _ in range(world)]
        out = {k: np.zeros_like(v) for k, v in full.items()}
        keep = []
        for r in range(world):
            blocks = [{k: np.ascontiguousarray(v[b:e]) for k, v in full.items()} for (b, e) in layout[r]]
            keep.append(blocks)
            segs = sharding.pack_segments([{k: a.ctypes.data for k, a in blk.items()} for blk in blocks], layout, r,
                                          offsets, slots[r].ctypes.data)
            lo, hi = slots[r].ctypes.data, slots[r].ctypes.data + nbytes
            for src, dst, nb in segs:
                assert lo <= dst and dst + nb <= hi
                ctypes.memmove(dst, src, nb)
        segs = sharding.unpack_segments([s.ctypes.data for s in slots], layout, offsets,
                                        {k: a.ctypes.data for k, a in out.items()})
        assert len(segs) <= 112
        for src, dst, nb in segs:
            ctypes.memmove(dst, src, nb)
        for k in full:
            assert np.array_equal(out[k], full[k]), (n_images, world, chunks, k)
