"""CPU, world_size 2 over gloo: the multi-GPU plumbing (image blocks, instance re-basing, gather to rank 0).
The per-rank computation is stood in for by the CPU oracle -- this tests the host logic of the N>1 path; the
CUDA kernels themselves are covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_images, result_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from detectron2_tensorflow_b200 import sharding
    from detectron2_tensorflow_b200.utils import synthetic as syn
    boxes, idx = syn.rois(n_images, 50, seed=1, image_hw=(250, 310))
    rng = np.random.default_rng(0)
    scores = rng.standard_normal((n_images, 50)).astype(np.float32)
    my_idx, my_boxes = sharding.shard_instances(torch.from_numpy(idx), torch.from_numpy(boxes), n_images, world, rank)
    b, e = sharding.image_block(n_images, world, rank)
    assert my_idx[:, 0].min().item() == 0 and my_idx[:, 0].max().item() == e - b - 1
    sh = sharding.shard_batch({"scores": torch.from_numpy(scores)}, world, rank)
    keep, num = oracle.batch_nms(my_boxes.numpy().reshape(e - b, 50, 4), sh["scores"].numpy(), 20, 0.5)
    full = sharding.gather_to_rank0({"keep": torch.from_numpy(keep), "num": torch.from_numpy(num)}, n_images)
    if rank == 0:
        np.savez(result_path, keep=full["keep"].numpy(), num=full["num"].numpy())
    else:
        assert full is None
    dist.destroy_process_group()


def _worker_p2p(rank, world, port, n_images, result_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from detectron2_tensorflow_b200 import sharding
    g = torch.Generator().manual_seed(7)
    full = {"boxes": torch.randn((n_images, 6, 4), generator=g), "valid": torch.rand((n_images, 6), generator=g) > 0.5,
            "classes": torch.randint(0, 80, (n_images, 6), generator=g, dtype=torch.int64)}
    layout = sharding.block_layout(n_images, world, chunks_of=lambda n: 2)
    blocks = [{k: v[b:e].clone() for k, v in full.items()} for (b, e) in layout[rank]]
    got = sharding.gather_blocks_to_rank0(blocks, layout)
    if rank == 0:
        assert all(torch.equal(got[k], full[k]) and got[k].dtype == full[k].dtype for k in full)
        # reuse of the preallocated outputs
        got2 = sharding.gather_blocks_to_rank0(blocks, layout, out=got)
        assert got2 is got and all(torch.equal(got[k], full[k]) for k in full)
        open(result_path, "w").write("ok")
    else:
        assert got is None
        assert sharding.gather_blocks_to_rank0(blocks, layout) is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5, 1])
def test_two_rank_grouped_p2p_gather(tmp_path, n_images):
    """The grouped send/recv gather (one batch of P2P ops, no packing): uneven blocks, sub-blocks, bool / int64."""
    path = str(tmp_path / "ok.txt")
    port = 31500 + os.getpid() % 2000 + n_images
    mp.spawn(_worker_p2p, args=(2, port, n_images, path), nprocs=2, join=True)
    assert open(path).read() == "ok"


def _worker_plan(rank, world, port, n_images, result_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from detectron2_tensorflow_b200 import sharding
    g = torch.Generator().manual_seed(11)
    full = {"boxes": torch.randn((n_images, 6, 4), generator=g), "valid": torch.rand((n_images, 6), generator=g) > 0.5,
            "classes": torch.randint(0, 80, (n_images, 6), generator=g, dtype=torch.int64)}
    spec = {k: (tuple(v.shape[1:]), v.dtype) for k, v in full.items()}
    layout = sharding.block_layout(n_images, world, chunks_of=lambda n: 2)
    plan = sharding.GatherPlan(spec, layout, torch.device("cpu"))
    for rep in range(2):  # the plan is persistent: a second step reuses every buffer
        blocks = [{k: v[b:e] + (rep if v.dtype == torch.float32 else 0) for k, v in full.items()} for (b, e) in layout[rank]]
        plan.pack(blocks)
        plan.gather()
        out = plan.unpack()
        if rank == 0:
            for k, v in full.items():
                assert torch.equal(out[k], v + (rep if v.dtype == torch.float32 else 0)), k
        else:
            assert out is None
    if rank == 0:
        open(result_path, "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5])
def test_two_rank_gather_plan(tmp_path, n_images):
    """GatherPlan (persistent send / receive buffers, one dist.gather, unpack into full-batch tensors)."""
    path = str(tmp_path / "ok.txt")
    port = 33500 + os.getpid() % 2000 + n_images
    mp.spawn(_worker_plan, args=(2, port, n_images, path), nprocs=2, join=True)
    assert open(path).read() == "ok"


def test_block_layout_covers_batch():
    from detectron2_tensorflow_b200.sharding import block_layout
    for n in (1, 5, 16):
        for w in (1, 2, 4, 8):
            lay = block_layout(n, w, chunks_of=lambda k: 2)
            flat = [be for r in lay for be in r if be[1] > be[0]]
            assert flat[0][0] == 0 and flat[-1][1] == n
            assert all(flat[i][1] == flat[i + 1][0] for i in range(len(flat) - 1))


@pytest.mark.parametrize("n_images", [4, 5])
def test_two_rank_gather_matches_single_rank(oracle_lib, tmp_path, n_images):
    from detectron2_tensorflow_b200.utils import synthetic as syn
    path = str(tmp_path / "out.npz")
    port = 29500 + os.getpid() % 2000 + n_images
    mp.spawn(_worker, args=(2, port, n_images, path), nprocs=2, join=True)
    z = np.load(path)
    boxes, idx = syn.rois(n_images, 50, seed=1, image_hw=(250, 310))
    scores = np.random.default_rng(0).standard_normal((n_images, 50)).astype(np.float32)
    keep, num = oracle_lib.batch_nms(boxes.reshape(n_images, 50, 4), scores, 20, 0.5)
    assert z["keep"].tobytes() == keep.tobytes() and z["num"].tobytes() == num.tobytes()


def test_image_blocks_cover_batch():
    from detectron2_tensorflow_b200.sharding import image_block
    for n in (0, 1, 7, 16, 32):
        for w in (1, 2, 4, 8):
            blocks = [image_block(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
