"""GPU: the peer-memory gather (csrc/peer.cu, sharding.PeerGatherPlan).

One GPU is enough: (1) the kernel's copy / wait / signal / time-out semantics with local pointers, (2) the whole
plan with TWO processes that both use cuda:0 (CUDA IPC maps an arena of another process of the same device too;
the kernels of the two processes time-slice, so the flag hand-shake is exercised for real).  On a multi-GPU box
tools/multi_gpu_check.py runs the same plan over NVLink."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _u64(dev, n=1):
    return torch.zeros(n * 16, dtype=torch.int64, device=dev)  # 128 bytes per flag


def test_peer_copy_segments_any_alignment(cuda):
    from detectron2_tensorflow_b200 import _native as nv
    g = torch.Generator().manual_seed(3)
    src = torch.randint(0, 256, (1 << 20,), generator=g, dtype=torch.uint8).to(cuda)
    dst = torch.zeros(1 << 20, dtype=torch.uint8, device=cuda)
    ctr, ticket, err = _u64(cuda), _u64(cuda), _u64(cuda)
    # (src offset, dst offset, bytes): 16-byte, 4-byte and byte paths, tails, an empty and a large segment
    table = [(0, 0, 4096), (4096 + 4, 8192 + 8, 1000), (20001, 30003, 777), (40000, 50000, 0), (65536, 131072, 300001),
             (500000, 600016, 15), (700001, 800001, 64)]
    segs = [(src.data_ptr() + s, dst.data_ptr() + d, n) for s, d, n in table]
    nv.peer_copy(segs, cuda, ctr.data_ptr(), ticket.data_ptr(), err.data_ptr())
    want = torch.zeros_like(dst)
    for s, d, n in table:
        want[d:d + n] = src[s:s + n]
    assert torch.equal(dst, want)
    assert int(ctr[0]) == 1 and int(ticket[0]) == 0 and int(err[0]) == 0
    nv.peer_copy(segs, cuda, ctr.data_ptr(), ticket.data_ptr(), err.data_ptr())
    assert int(ctr[0]) == 2


def test_peer_copy_wait_signal_and_timeout(cuda):
    from detectron2_tensorflow_b200 import _native as nv
    a = torch.arange(1000, dtype=torch.float32, device=cuda)
    b = torch.zeros_like(a)
    c = torch.zeros_like(a)
    flags = _u64(cuda, 2)
    f0, f1 = flags.data_ptr(), flags.data_ptr() + 128
    send = [_u64(cuda) for _ in range(3)]
    recv = [_u64(cuda) for _ in range(3)]
    s1, s2 = torch.cuda.Stream(cuda), torch.cuda.Stream(cuda)
    torch.cuda.synchronize()
    for step in range(1, 4):
        a.fill_(float(step))
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):  # the receiver is enqueued FIRST and has to wait for the sender's flag
            nv.peer_copy([(b.data_ptr(), c.data_ptr(), 4000)], cuda, *[t.data_ptr() for t in recv], wait_flags=[f0],
                         wait_lag=0, signal_flags=[f1])
        with torch.cuda.stream(s2):  # sender: waits for the acknowledgement of the previous step
            nv.peer_copy([(a.data_ptr(), b.data_ptr(), 4000)], cuda, *[t.data_ptr() for t in send], wait_flags=[f1],
                         wait_lag=1, signal_flags=[f0])
        torch.cuda.synchronize()
        assert torch.equal(c, a) and int(flags[0]) == step and int(flags[16]) == step
        assert int(recv[2][0]) == 0 and int(send[2][0]) == 0
    # a flag that never comes: the kernel gives up after the time-out, reports it and still finishes
    nv.peer_copy([(a.data_ptr(), c.data_ptr(), 4000)], cuda, *[t.data_ptr() for t in recv], wait_flags=[f0], wait_lag=0,
                 timeout_ms=50)
    torch.cuda.synchronize()
    assert int(recv[2][0]) == 1 and int(recv[0][0]) == 4
    # with the error word set, later calls do not wait again (one time-out per plan, not one per step)
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        nv.peer_copy([(a.data_ptr(), c.data_ptr(), 4000)], cuda, *[t.data_ptr() for t in recv], wait_flags=[f0], wait_lag=0,
                     timeout_ms=1000)
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 1.0 and int(recv[0][0]) == 24


def test_peer_copy_validation(cuda):
    from detectron2_tensorflow_b200 import _native as nv
    t = _u64(cuda, 3)
    p = [t.data_ptr(), t.data_ptr() + 128, t.data_ptr() + 256]
    with pytest.raises(ValueError):
        nv.peer_copy([(p[0], p[1], 8)] * 113, cuda, *p)
    with pytest.raises(ValueError):
        nv.peer_copy([], cuda, *p, wait_flags=[p[0]] * 17)
    with pytest.raises(ValueError):
        nv.peer_copy([], cuda, *p, wait_flags=[p[0]], wait_lag=2)
    with pytest.raises(ValueError):
        nv.peer_copy([(0, p[1], 8)], cuda, *p)


def _worker(rank, world, port, n_images, chunks, steps, result_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from detectron2_tensorflow_b200 import sharding
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator().manual_seed(11)
    full = {"boxes": torch.randn((n_images, 1000, 4), generator=g), "valid": torch.rand((n_images, 1000), generator=g) > 0.5,
            "classes": torch.randint(0, 80, (n_images, 100), generator=g, dtype=torch.int64),
            "odd": torch.randint(0, 255, (n_images, 7), generator=g, dtype=torch.uint8)}
    spec = {k: (tuple(v.shape[1:]), v.dtype) for k, v in full.items()}
    layout = sharding.block_layout(n_images, world, chunks_of=lambda n: chunks)
    plan = sharding.PeerGatherPlan(spec, layout, dev, timeout_ms=20000)
    blocks = [{k: v[b:e].to(dev).contiguous() for k, v in full.items()} for (b, e) in layout[rank]]
    ok = True
    graph = torch.cuda.CUDAGraph()  # steps 2.. replay the captured kernels (the epoch lives in device memory)
    side = torch.cuda.Stream(dev)
    for step in range(steps):
        for blk, (b, e) in zip(blocks, layout[rank]):
            blk["boxes"].copy_(full["boxes"][b:e] + step)
        if step == 0:
            plan.pack(blocks)
            plan.unpack()
        elif step == 1:
            with torch.cuda.graph(graph, stream=side):
                plan.pack(blocks)
                plan.unpack()
            graph.replay()
        else:
            graph.replay()
        torch.cuda.synchronize()
        plan.check()
        if rank == 0:
            for k, v in full.items():
                ok = ok and torch.equal(plan.out[k].cpu(), v + step if k == "boxes" else v)
        dist.barrier()
    plan.close()
    if rank == 0:
        open(result_path, "w").write("ok" if ok else "mismatch")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_images,chunks", [(2, 5, 2), (3, 16, 1), (4, 3, 1)])
def test_peer_gather_plan_processes_sharing_one_gpu(cuda, tmp_path, world, n_images, chunks):
    path = str(tmp_path / "ok.txt")
    port = 35500 + os.getpid() % 2000 + world
    mp.spawn(_worker, args=(world, port, n_images, chunks, 4, path), nprocs=world, join=True)
    assert open(path).read() == "ok"
