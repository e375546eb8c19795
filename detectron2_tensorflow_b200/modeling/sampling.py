"""Drop-in for lib/modeling/sampling.py: `subsample_labels` (:6-45)."""
import torch

from .. import _native as nv


def subsample_labels_batched(labels, num_samples, positive_fraction, bg_label, seed=0, return_labels=False):
    """`subsample_labels` for a batch of label vectors in one call, without a device->host sync.

    labels [N, P] int64 (-1 ignore, `bg_label` negative, anything else positive).  Returns
    (pos_idx [N, num_samples] int64 padded with -1, neg_idx likewise, num_pos [N] int32, num_neg [N] int32) and, with
    `return_labels`, the resampled label matrix [N, P] (labels of the sampled elements, -1 elsewhere: what
    RPNOutputs.losses' `resample` builds with tf.dynamic_stitch, rpn_outputs.py:315-329)."""
    host = not labels.is_cuda
    dev = nv.device_of(labels)
    lab = nv.to_device(labels, dev, torch.int64)
    assert lab.dim() == 2
    N, P = lab.shape
    k = int(num_samples)
    pos = torch.empty((N, k), dtype=torch.int64, device=dev)
    neg = torch.empty((N, k), dtype=torch.int64, device=dev)
    npos = torch.empty(N, dtype=torch.int32, device=dev)
    nneg = torch.empty(N, dtype=torch.int32, device=dev)
    out_labels = torch.empty((N, P), dtype=torch.int64, device=dev) if return_labels else None
    p = nv.SubsampleLabelsParams()
    p.labels = lab.data_ptr()
    p.num_images, p.num_labels, p.num_samples = N, P, k
    p.max_positives = int(k * positive_fraction)  # sampling.py:37, in Python's double arithmetic
    p.bg_label = int(bg_label)
    p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    p.out_pos_idx, p.out_neg_idx = pos.data_ptr(), neg.data_ptr()
    p.out_num_pos, p.out_num_neg = npos.data_ptr(), nneg.data_ptr()
    p.out_labels = nv.ptr(out_labels)
    nv.call("subsample_labels", p, dev)
    res = (pos, neg, npos, nneg) + ((out_labels,) if return_labels else ())
    return tuple(nv.to_host(t) for t in res) if host else res


def subsample_labels(labels, num_samples, positive_fraction, bg_label, seed=0):
    """
    Return `num_samples` random samples from `labels`, with a fraction of positives no larger than
    `positive_fraction` (reference signature, sampling.py:6-45, plus `seed`: the sample is a deterministic function
    of it -- TF's random_shuffle has no defined bit pattern, so parity is distributional).

    Args:
        labels (Tensor): (N, ) label vector: -1 ignore, `bg_label` negative, otherwise positive.
    Returns:
        pos_idx, neg_idx (Tensor): 1D int64 indices (dynamic lengths, hence one device->host read of the counts).
    """
    assert labels.dim() == 1
    pos, neg, npos, nneg = subsample_labels_batched(labels[None], num_samples, positive_fraction, bg_label, seed)
    return pos[0, :int(npos[0])], neg[0, :int(nneg[0])]
