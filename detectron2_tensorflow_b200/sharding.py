"""Multi-GPU sharding of the hot path: images are independent (the reference runs every stage under
tf.map_fn over the batch: rpn_outputs.py:123, fast_rcnn.py:171, retinanet.py:375, solo_v2.py:587), so the
batch is split into contiguous image blocks, one per rank, with NO collective inside the path.  The only
communication is an optional final gather of the fixed-size padded outputs to rank 0 (NCCL over
NVLink/NVSwitch on GPUs; gloo in the CPU tests).  ROI features are consumed on the GPU that produced them
and are never gathered.
"""
import torch
import torch.distributed as dist


def image_block(num_images, world_size, rank):
    """Contiguous block [begin, end) of images owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_images, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(tensors, world_size, rank):
    """Slice every [N, ...] tensor (or list of them) to this rank's image block."""
    def cut(t):
        b, e = image_block(t.shape[0], world_size, rank)
        return t[b:e]
    return {k: ([cut(t) for t in v] if isinstance(v, (list, tuple)) else cut(v)) for k, v in tensors.items()}


def shard_instances(indices, boxes, num_images, world_size, rank):
    """Rows of a SparseBoxList (indices [M,2] image-major) that belong to this rank, image index re-based."""
    b, e = image_block(num_images, world_size, rank)
    sel = (indices[:, 0] >= b) & (indices[:, 0] < e)
    idx = indices[sel].clone()
    idx[:, 0] -= b
    return idx, boxes[sel]


def gather_to_rank0(local, num_images, group=None):
    """Gather fixed-size per-image outputs ([n_local, ...] tensors in a dict) to rank 0 in image order.
    Returns the full-batch dict on rank 0 and None elsewhere.  Blocks may differ by one image, so shorter
    blocks are padded to the longest before the (equal-size) gather and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [image_block(num_images, world, r) for r in range(world)]
    longest = max(e - b for b, e in sizes)
    out = {}
    for k, t in local.items():
        pad = longest - t.shape[0]
        if pad:
            t = torch.cat([t, t.new_zeros((pad,) + tuple(t.shape[1:]))])
        t = t.contiguous()
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, bufs, dst=0, group=group)
        if rank == 0:
            out[k] = torch.cat([bufs[r][:sizes[r][1] - sizes[r][0]] for r in range(world)])
    return out if rank == 0 else None
