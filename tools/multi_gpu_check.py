"""SURVEY.md 8(e) on real GPUs: images sharded over the ranks (one process per GPU, torchrun), no collective on the
data path, final gather of the fixed-size padded outputs to rank 0 over NCCL (NVLink/NVSwitch), and the gathered
bytes compared with a single-GPU run of the whole batch.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from detectron2_tensorflow_b200 import sharding
from detectron2_tensorflow_b200.engine import MaskRCNNPostBackbone
from detectron2_tensorflow_b200.utils import synthetic as syn

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

N = 2 * world + 1  # uneven blocks on purpose
R, D, K, C = 300, 50, 80, 64
anchors = syn.rpn_anchors()
logits, deltas = syn.rpn_inputs(N, seed=2, variant="gaussian", anchors=anchors)
feats = syn.fpn_features(N, C, seed=0)
scores, cls_deltas = syn.fast_rcnn_inputs(N, R, K, seed=5)
shapes = syn.image_shapes(N)


def to_dev(b, e):
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return dict(anchors=[T(a) for a in anchors], logits=[T(x[b:e]) for x in logits], deltas=[T(x[b:e]) for x in deltas],
                feats=[T(f[b:e]) for f in feats], shapes=T(shapes[b:e]), scores=T(scores[b * R:e * R]),
                cls_deltas=T(cls_deltas[b * R:e * R]))


eng = MaskRCNNPostBackbone(rois_per_image=R, dets_per_image=D, pre_nms_topk=1000)
b, e = sharding.image_block(N, world, rank)
out = MaskRCNNPostBackbone.flatten_outputs(eng(to_dev(b, e)))
small = {k: (v.to(torch.uint8) if v.dtype == torch.bool else v) for k, v in out.items() if not k.endswith("_feats")}
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
full = sharding.gather_to_rank0(small, N)
e1.record()
torch.cuda.synchronize()
gather_ms = e0.elapsed_time(e1)
# the same gather over NVLink peer memory (sharding.PeerGatherPlan: pack kernel with remote stores + flags, unpack kernel
# on rank 0), several steps so that the acknowledgement hand-shake is exercised
from detectron2_tensorflow_b200.engine import GATHER_KEYS
layout = sharding.block_layout(N, world)
plan = sharding.PeerGatherPlan(eng.gather_spec(), layout, dev)
blk = {k: out[k].contiguous() for k in GATHER_KEYS}
peer_ms = []
for step in range(5):
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    plan.pack([blk])
    plan.unpack()
    e1.record()
    torch.cuda.synchronize()
    peer_ms.append(e0.elapsed_time(e1))
plan.check()
peer_full = {k: v.clone() for k, v in plan.out.items()} if rank == 0 else None
plan.close()
res = None
if rank == 0:
    ref = MaskRCNNPostBackbone.flatten_outputs(eng(to_dev(0, N)))
    same = {k: bool(torch.equal(full[k], ref[k].to(torch.uint8) if ref[k].dtype == torch.bool else ref[k])) for k in full}
    same_peer = {k: bool(torch.equal(peer_full[k], ref[k])) for k in GATHER_KEYS}
    nbytes = sum(v.numel() * v.element_size() for v in full.values())
    res = {"check": "sharded run + NCCL gather == single-GPU run (byte-identical)", "world": world, "images": N,
           "blocks": [sharding.image_block(N, world, r) for r in range(world)], "identical": same,
           "all_identical": all(same.values()) and all(same_peer.values()), "gathered_bytes": nbytes,
           "gather_ms": gather_ms, "peer_memory_gather_identical": same_peer, "peer_memory_gather_ms_per_step": peer_ms}
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
if rank == 0 and not res["all_identical"]:
    sys.exit(1)
