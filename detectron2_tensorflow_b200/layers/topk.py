"""Segmented top-k (tf.nn.top_k per (image, level) row; rpn_outputs.py:70, retinanet.py:326)."""
import torch

from .. import _native as nv


def segmented_top_k(rows_per_level, k, sigmoid=False, k_limits=None):
    """rows_per_level: list of [N, len_l] fp32 tensors.  Returns (values, indices, counts) with
    values/indices [N, L, k] sorted (value desc, index asc), padded with 0 / -1, counts [N, L]."""
    dev = nv.device_of(*rows_per_level)
    host = not rows_per_level[0].is_cuda
    xs = [nv.to_device(x, dev, torch.float32) for x in rows_per_level]
    N, L = xs[0].shape[0], len(xs)
    vals = torch.empty((N, L, k), dtype=torch.float32, device=dev)
    idx = torch.empty((N, L, k), dtype=torch.int32, device=dev)
    cnt = torch.empty((N, L), dtype=torch.int32, device=dev)
    p = nv.SegmentedTopkParams()
    for l, x in enumerate(xs):
        p.scores[l] = x.data_ptr()
        p.row_len[l] = x.shape[1]
        p.k_limit[l] = 0 if k_limits is None else int(k_limits[l])
    p.num_groups, p.rows_per_group, p.k = L, N, int(k)
    p.transform = nv.TOPK_SIGMOID if sigmoid else nv.TOPK_IDENTITY
    p.out_values, p.out_indices, p.out_counts = vals.data_ptr(), idx.data_ptr(), cnt.data_ptr()
    nv.call("segmented_topk", p, dev)
    if host:
        return nv.to_host(vals), nv.to_host(idx), nv.to_host(cnt)
    return vals, idx, cnt
