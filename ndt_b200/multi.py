"""Multi-GPU plumbing: a dynamic tile queue and the gather of finished tiles.

The path shards trivially (pixels and frames are independent, ndt.c:750-757):
there is no data-path collective.  The only exchange is "finished tiles ->
rank 0", the analogue of the reference's mpi_collect_image (ndt.c:1277-1309,
a software tree of image_add over zero-padded full frames).  Here every rank
sends exactly the rows it rendered with NCCL send/recv, straight into the
destination rows of rank 0's frame buffer.

The work distribution replaces the reference's static cyclic rows
(ndt.c:812-820) with one shared counter: ranks pull the next (frame, band) item
when they are free, so expensive bands (the hypercube's silhouette) do not
serialise behind a static assignment.  The counter lives in the process
group's key-value store (rank 0's TCPStore); one add() per item.

Backend-agnostic on purpose: the CPU test tier runs it with gloo and
world_size 2 (tests/test_multi.py), the bench with nccl.
"""
import torch


class TileQueue:
    def __init__(self, dist, rank, world, name="ndtq"):
        self.dist, self.rank, self.world, self.name = dist, rank, world, name
        self.store = None
        if dist is not None and world > 1:
            from torch.distributed import distributed_c10d
            self.store = distributed_c10d._get_default_store()

    def pull(self, step_id, n_items):
        """Yields item indices until the step's queue is empty."""
        if self.store is None:
            yield from range(n_items)
            return
        key = f"{self.name}/{step_id}"
        while True:
            i = self.store.add(key, 1) - 1
            if i >= n_items:
                return
            yield i


def gather_tiles(dist, rank, world, items, mine, stage, frames, band):
    """Move the tiles this rank rendered to rank 0.

    items : list of (frame, y0, rows) for the whole step
    mine  : indices into `items` this rank rendered, in render order
    stage : [len(items), band, W, 4] uint8 on ranks != 0: tile k of `mine` is stage[k, :rows]
    frames: [n_frames, H, W, 4] uint8 on rank 0 (rank 0 renders straight into it)
    """
    n = len(items)
    dev = frames.device if rank == 0 else stage.device
    ids = torch.full((n + 1,), -1, dtype=torch.int32, device=dev)
    ids[0] = len(mine)
    if mine:
        ids[1:1 + len(mine)] = torch.tensor(mine, dtype=torch.int32, device=dev)
    table = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(table, ids)
    ops = []
    if rank == 0:
        for r in range(1, world):
            t = table[r].tolist()
            for k in range(t[0]):
                f, y0, rows = items[t[1 + k]]
                ops.append(dist.P2POp(dist.irecv, frames[f, y0:y0 + rows], r))
    else:
        for k, idx in enumerate(mine):
            f, y0, rows = items[idx]
            ops.append(dist.P2POp(dist.isend, stage[k, :rows], 0))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    owners = {}
    for r in range(world):
        t = table[r].tolist()
        for k in range(t[0]):
            owners[t[1 + k]] = r
    return owners


class FrameGather:
    """Finished frames -> rank 0, each sent as soon as it is rendered, while the next one renders.

    Every rank renders `frames_per_rank` frames per step (global frame f = rank * frames_per_rank + j).  Rank 0
    posts the receives of a whole step up front, one grouped NCCL operation per round j (world - 1 receives
    straight into the destination frames); rank k issues send j when its frame j is done -- frames finish out of
    order with two contexts per GPU, the sends go out in order j = 0, 1, ... because NCCL matches the operations
    of a pair by their order.  Nothing here waits for the GPU except end().  This is the overlapped form of the
    reference's mpi_collect_image (ndt.c:1277-1309); backend-agnostic (gloo in the CPU tests)."""

    def __init__(self, dist, rank, world, frames_per_rank):
        self.dist, self.rank, self.world, self.F = dist, rank, world, frames_per_rank
        self.reqs = []
        self.ready = {}
        self.next_j = 0

    def begin(self, frames):
        """frames: [world * F, ...] on rank 0 (None elsewhere)"""
        d = self.dist
        self.reqs, self.ready, self.next_j = [], {}, 0
        if self.rank == 0:
            for j in range(self.F):
                ops = [d.P2POp(d.irecv, frames[k * self.F + j], k) for k in range(1, self.world)]
                if ops:
                    self.reqs += d.batch_isend_irecv(ops)

    def send(self, tensor, j):
        """rank != 0: frame j of this step is complete in `tensor`"""
        d = self.dist
        self.ready[j] = tensor
        while self.next_j in self.ready:
            t = self.ready.pop(self.next_j)
            self.reqs += d.batch_isend_irecv([d.P2POp(d.isend, t, 0)])
            self.next_j += 1

    def end(self):
        if not self.reqs and not self.ready and self.next_j == 0:
            return                  # nothing begun
        assert self.rank == 0 or self.next_j == self.F, "a frame of this step was never handed to send()"
        cuda = False
        for r in self.reqs:
            r.wait()
            cuda = True
        self.reqs, self.next_j = [], 0
        if cuda and torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
