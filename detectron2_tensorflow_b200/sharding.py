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


def _field_layout(local):
    """Per-image byte layout of the packed row: key -> (offset, nbytes, trailing shape, dtype); 16-byte aligned."""
    layout, off = {}, 0
    for k, t in local.items():
        per = t.element_size()
        for d in t.shape[1:]:
            per *= int(d)
        layout[k] = (off, per, tuple(t.shape[1:]), t.dtype)
        off += (per + 15) // 16 * 16
    return layout, off


def pack_rows(local, rows):
    """One [rows, bytes_per_image] uint8 buffer holding every tensor of `local` ([n_local, ...] each, n_local <=
    rows; missing rows are zero): a single concatenation kernel."""
    layout, width = _field_layout(local)
    n = next(iter(local.values())).shape[0]
    parts = []
    for k, t in local.items():
        off, per, _, _ = layout[k]
        b = t.contiguous().reshape(n, -1).view(torch.uint8)
        pad = (per + 15) // 16 * 16 - per
        parts.append(b if not pad else torch.nn.functional.pad(b, (0, pad)))
    buf = torch.cat(parts, dim=1)
    if n < rows:
        buf = torch.cat([buf, buf.new_zeros((rows - n, width))])
    return buf, layout


def unpack_rows(buf, layout):
    """Inverse of pack_rows on a [rows, bytes_per_image] buffer."""
    out = {}
    for k, (off, per, shape, dtype) in layout.items():
        out[k] = buf[:, off:off + per].contiguous().view(dtype).reshape((buf.shape[0],) + shape)
    return out


def gather_to_rank0(local, num_images, group=None):
    """Gather fixed-size per-image outputs ([n_local, ...] tensors in a dict) to rank 0 in image order with ONE
    collective: every rank packs its tensors into one [longest_block, bytes_per_image] byte buffer (one kernel),
    `dist.gather` moves it (NCCL: one grouped ncclSend/ncclRecv over NVLink; gloo in the CPU tests), rank 0 drops
    the padding rows of shorter blocks and unpacks.  Returns the full-batch dict on rank 0 and None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [image_block(num_images, world, r) for r in range(world)]
    longest = max(e - b for b, e in sizes)
    buf, layout = pack_rows(local, longest)
    recv = None
    if rank == 0:
        recv = buf.new_empty((world,) + tuple(buf.shape))
    dist.gather(buf, list(recv.unbind(0)) if rank == 0 else None, dst=0, group=group)
    if rank != 0:
        return None
    if all(e - b == longest for b, e in sizes):
        rows = recv.reshape(world * longest, -1)
    else:
        rows = torch.cat([recv[r, :sizes[r][1] - sizes[r][0]] for r in range(world)])
    return unpack_rows(rows, layout)


def block_layout(num_images, world_size, chunks_of=None):
    """layout[r] = the (begin, end) global image ranges rank r holds: its image block, cut into `chunks_of(n_r)`
    sub-blocks (the image blocks a graphed step runs on concurrent streams).  Identical on every rank."""
    layout = []
    for r in range(world_size):
        b, e = image_block(num_images, world_size, r)
        c = max(1, min(int(chunks_of(e - b)) if chunks_of else 1, max(e - b, 1)))
        layout.append([(b + image_block(e - b, c, i)[0], b + image_block(e - b, c, i)[1]) for i in range(c)])
    return layout


def gather_blocks_to_rank0(blocks, layout, out=None, group=None):
    """Gather the fixed-size per-image outputs of every rank to rank 0 with ONE grouped batch of point-to-point
    transfers and no packing: `blocks[i]` is this rank's dict of [e - b, ...] tensors for `layout[rank][i]`; rank 0
    receives every (rank, block, tensor) straight into its slot `out[key][b:e]` of the full-batch tensors (NCCL: the
    whole batch is one ncclGroupStart/End, i.e. one fused send/recv kernel per peer over NVLink; gloo on CPU).
    `out` (rank 0, optional) = preallocated {key: [N, ...]} tensors to reuse.  Returns `out` on rank 0, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ops = []
    if rank == 0:
        n_total = layout[-1][-1][1]
        if out is None:
            out = {k: t.new_empty((n_total,) + tuple(t.shape[1:])) for k, t in blocks[0].items()}
        for (b, e), blk in zip(layout[0], blocks):
            for k, t in blk.items():
                out[k][b:e].copy_(t)
        for r in range(1, world):
            for (b, e) in layout[r]:
                if e > b:
                    for k in out:
                        ops.append(dist.P2POp(dist.irecv, out[k][b:e], r, group))
    else:
        for (b, e), blk in zip(layout[rank], blocks):
            if e > b:
                for k, t in blk.items():
                    ops.append(dist.P2POp(dist.isend, t.contiguous(), 0, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out if rank == 0 else None


class GatherPlan(object):
    """Steady-state gather of the fixed-size per-image outputs to rank 0, set up once per graphed step:

      pack()    copies this rank's block tensors into ONE persistent send buffer (call it inside the CUDA-graph
                capture of the step: the copies then cost graph nodes, not eager launches);
      gather()  ONE `dist.gather` of that buffer (NCCL: one grouped send/recv over NVLink; gloo on CPU);
      unpack()  rank 0: copies every rank's blocks from the receive buffer into the full-batch tensors `out[key]`
                ([N, ...], image order; capture it once as a CUDA graph and replay it).

    Buffer layout: one block per key sized for the LONGEST image block, so the key offsets are the same on every
    rank and uneven blocks only leave a tail unused.  spec = {key: (trailing shape, dtype)}."""

    def __init__(self, spec, layout, device, group=None, collective="auto"):
        """collective: "gather" (dist.gather: grouped send/recv to rank 0), "all_gather" (one
        all_gather_into_tensor: every rank receives the 48 KB blocks -- NCCL has no native gather and this is its
        cheapest fixed-size collective), or "auto" (all_gather on NCCL, gather elsewhere)."""
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if collective == "auto":
            collective = "all_gather" if dist.get_backend(group) == "nccl" else "gather"
        self.collective = collective
        self.layout = layout
        self.spec = dict(spec)
        self.begin = [lay[0][0] for lay in layout]
        self.count = [lay[-1][1] - lay[0][0] for lay in layout]
        longest = max(self.count)
        self.offsets, off = {}, 0
        for k, (shape, dtype) in self.spec.items():
            per = torch.empty(0, dtype=dtype).element_size()
            for d in shape:
                per *= int(d)
            self.offsets[k] = (off, per)
            off += (longest * per + 255) // 256 * 256
        self.nbytes = max(off, 256)
        self.sendbuf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        need_recv = self.rank == 0 or self.collective == "all_gather"
        self.recv = torch.zeros((self.world, self.nbytes), dtype=torch.uint8, device=device) if need_recv else None
        n_total = layout[-1][-1][1]
        self.out = {k: torch.zeros((n_total,) + tuple(shape), dtype=dtype, device=device)
                    for k, (shape, dtype) in self.spec.items()} if self.rank == 0 else None

    def _view(self, buf, key, n):
        off, per = self.offsets[key]
        shape, dtype = self.spec[key]
        return buf[off:off + n * per].view(dtype).reshape((n,) + tuple(shape))

    def pack(self, blocks):
        """blocks[i] = dict of [e - b, ...] tensors for layout[rank][i]."""
        n = self.count[self.rank]
        for (b, e), blk in zip(self.layout[self.rank], blocks):
            lb = b - self.begin[self.rank]
            for k in self.spec:
                self._view(self.sendbuf, k, n)[lb:lb + (e - b)].copy_(blk[k])

    def gather(self):
        if self.collective == "all_gather":
            dist.all_gather_into_tensor(self.recv.view(-1), self.sendbuf, group=self.group)
        else:
            dist.gather(self.sendbuf, list(self.recv.unbind(0)) if self.rank == 0 else None, dst=0, group=self.group)

    def unpack(self):
        if self.rank != 0:
            return None
        for r in range(self.world):
            n = self.count[r]
            if n:
                for k in self.spec:
                    self.out[k][self.begin[r]:self.begin[r] + n].copy_(self._view(self.recv[r], k, n))
        return self.out


def gather_offsets(spec, layout):
    """Byte layout of one rank's receive slot: {key: (offset, bytes per image)}, slot size.  One region per key sized
    for the LONGEST image block (256-byte aligned), so the offsets are the same for every rank."""
    longest = max(lay[-1][1] - lay[0][0] for lay in layout)
    offsets, off = {}, 0
    for k, (shape, dtype) in spec.items():
        per = torch.empty(0, dtype=dtype).element_size()
        for d in shape:
            per *= int(d)
        offsets[k] = (off, per)
        off += (longest * per + 255) // 256 * 256
    return offsets, max(off, 256)


def pack_segments(block_addrs, layout, rank, offsets, slot_addr):
    """Copy table of a sender: block_addrs[i][key] = address of this rank's [e - b, ...] tensor of image block
    layout[rank][i]; every tensor goes to its rows of the receive slot at `slot_addr` (rank 0's memory).
    Returns [(src, dst, nbytes)]."""
    begin = layout[rank][0][0]
    segs = []
    for (b, e), blk in zip(layout[rank], block_addrs):
        if e > b:
            for k, (off, per) in offsets.items():
                segs.append((blk[k], slot_addr + off + (b - begin) * per, (e - b) * per))
    return segs


def unpack_segments(slot_addrs, layout, offsets, out_addrs, ranks=None):
    """Copy table of rank 0: the rows of every rank's receive slot go to that rank's image rows of the full-batch
    tensors (out_addrs[key] = address of the [N, ...] tensor).  Returns [(src, dst, nbytes)]."""
    segs = []
    for r in (range(len(layout)) if ranks is None else ranks):
        begin, n = layout[r][0][0], layout[r][-1][1] - layout[r][0][0]
        if n > 0:
            for k, (off, per) in offsets.items():
                segs.append((slot_addrs[r] + off, out_addrs[k] + begin * per, n * per))
    return segs


class PeerGatherPlan(object):
    """The same steady-state gather as GatherPlan, over NVLink PEER MEMORY instead of a collective (CUDA only):

      pack()    ONE kernel on every sender: waits for rank 0's acknowledgement of the previous step, stores this
                rank's block tensors straight into its receive slot in rank 0's memory, raises its flag there.  On
                rank 0 the same kernel copies the local blocks into the full-batch tensors.  Capture it inside the
                step's CUDA graph.
      gather()  nothing to do (the transfer is the pack kernel's stores).
      unpack()  rank 0, ONE kernel: waits for every sender's flag of this step, scatters the slots into `out[key]`
                ([N, ...], image order), acknowledges into the senders' arenas.  Capturable too.

    Every rank must run pack() / unpack() the same number of times (the epoch lives in device memory and is advanced
    by the kernels).  A peer that never shows up makes the waiting kernel give up after `timeout_ms` and set the
    error flag, which `check()` turns into an exception: the GPU never hangs.  Arena layout (one cudaMalloc per
    rank, mapped by CUDA IPC): flags[world] | ack | counters | receive slots[world]; 128 bytes per flag."""

    FLAG = 128

    def __init__(self, spec, layout, device, group=None, timeout_ms=2000):
        from . import _native as nv
        self._nv = nv
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.layout = layout
        self.spec = dict(spec)
        self.timeout_ms = int(timeout_ms)
        self.offsets, self.nbytes = gather_offsets(self.spec, layout)
        F = self.FLAG
        self._flags_off, self._ack_off, self._ctr_off = 0, self.world * F, (self.world + 1) * F
        self._slot_off = (self.world + 1) * F + 8 * F
        arena_bytes = self._slot_off + (self.world * self.nbytes if self.rank == 0 else 0)
        # set-up is collective: every step is followed by an agreement on its outcome, so that a rank whose CUDA IPC
        # call fails (container without IPC, peer access unavailable) makes EVERY rank raise instead of leaving the
        # others in a barrier
        self.arena, self._mapped, err = None, {}, None
        try:
            self.arena = nv.peer_alloc(arena_bytes, device)
            handle = nv.peer_export(self.arena, device)
        except Exception as e:  # noqa: BLE001
            handle, err = None, repr(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, (handle, err), group=group)
        if all(h is not None for h, _ in handles):
            try:
                if self.rank == 0:
                    for r in range(1, self.world):
                        self._mapped[r] = nv.peer_open(handles[r][0], device)
                else:
                    self._mapped[0] = nv.peer_open(handles[0][0], device)
            except Exception as e:  # noqa: BLE001
                err = repr(e)
        status = [None] * self.world
        dist.all_gather_object(status, err, group=group)
        failed = [f"rank {r}: {handles[r][1] or status[r]}" for r in range(self.world) if handles[r][1] or status[r]]
        if failed:
            self._release()
            raise nv.D2BError("PeerGatherPlan: peer-memory set-up failed (" + "; ".join(failed)[:400] + ")")
        n_total = layout[-1][-1][1]
        self.out = {k: torch.zeros((n_total,) + tuple(shape), dtype=dtype, device=device)
                    for k, (shape, dtype) in self.spec.items()} if self.rank == 0 else None
        self._keep = []
        nv.peer_copy([], device, self._ctr(5), self._ctr(6), self._ctr(7))  # loads the kernel outside any capture
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    # counters of this rank's arena: 0 = pack epoch, 1 = pack ticket, 2 = unpack epoch, 3 = unpack ticket, 4 = error
    def _ctr(self, i):
        return self.arena + self._ctr_off + i * self.FLAG

    def pack(self, blocks):
        """blocks[i] = dict of [e - b, ...] contiguous tensors for layout[rank][i]."""
        nv = self._nv
        addrs = []
        self._keep = []  # the tensors whose addresses the (possibly captured) kernel was given
        for blk in blocks:
            for k in self.spec:
                assert blk[k].is_contiguous() and blk[k].dtype == self.spec[k][1]
            addrs.append({k: blk[k].data_ptr() for k in self.spec})
            self._keep.append([blk[k] for k in self.spec])
        if self.rank == 0:
            segs = []  # the local blocks go straight to their rows of the full-batch tensors
            for (b, e), blk in zip(self.layout[0], addrs):
                for k, (off, per) in self.offsets.items():
                    segs.append((blk[k], self.out[k].data_ptr() + b * per, (e - b) * per))
            nv.peer_copy(segs, self.device, self._ctr(0), self._ctr(1), self._ctr(4), timeout_ms=self.timeout_ms)
        else:
            base0 = self._mapped[0]
            slot = base0 + self._slot_off + self.rank * self.nbytes
            segs = pack_segments(addrs, self.layout, self.rank, self.offsets, slot)
            nv.peer_copy(segs, self.device, self._ctr(0), self._ctr(1), self._ctr(4),
                         wait_flags=[self.arena + self._ack_off], wait_lag=1,
                         signal_flags=[base0 + self._flags_off + self.rank * self.FLAG], timeout_ms=self.timeout_ms)

    def gather(self):
        return None

    def unpack(self):
        if self.rank != 0:
            return None
        nv = self._nv
        slots = [self.arena + self._slot_off + r * self.nbytes for r in range(self.world)]
        segs = unpack_segments(slots, self.layout, self.offsets, {k: t.data_ptr() for k, t in self.out.items()},
                               ranks=range(1, self.world))
        nv.peer_copy(segs, self.device, self._ctr(2), self._ctr(3), self._ctr(4),
                     wait_flags=[self.arena + self._flags_off + r * self.FLAG for r in range(1, self.world)], wait_lag=0,
                     signal_flags=[self._mapped[r] + self._ack_off for r in range(1, self.world)],
                     timeout_ms=self.timeout_ms)
        return self.out

    def check(self):
        """Raise if a wait of this rank ever timed out (host sync)."""
        err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._nv.peer_copy([(self._ctr(4), err.data_ptr(), 4)], self.device, self._ctr(5), self._ctr(6), self._ctr(7))
        if int(err.item()) != 0:
            raise self._nv.D2BError("PeerGatherPlan: a wait on a peer's flag timed out (ranks out of step or a peer died)")

    def _release(self):
        for a in self._mapped.values():
            try:
                self._nv.peer_close(a, self.device)
            except Exception:  # noqa: BLE001
                pass
        self._mapped = {}
        if self.arena is not None:
            try:
                self._nv.peer_free(self.arena, self.device)
            except Exception:  # noqa: BLE001
                pass
        self.arena = None

    def close(self):
        """Unmap the peers' arenas and free this rank's (collective: every rank calls it)."""
        if self.arena is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for a in self._mapped.values():
            self._nv.peer_close(a, self.device)
        self._mapped = {}
        dist.barrier(group=self.group)
        self._nv.peer_free(self.arena, self.device)
        self.arena = None
