"""N>1 host logic on the CPU: world_size-2 gloo processes share the dynamic tile
queue and gather tiles to rank 0; the assembled frames must equal the
single-process result.  The oracle stands in for the GPU renderer here (tests
may use it); the queue / gather code is the same the NCCL bench runs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, load_flat, oracle_render


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import ctypes as C
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ndt_b200 import multi
    from conftest import load_flat, oracle_render
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = C.CDLL(os.path.join(ROOT, "oracle", "libndt_oracle.so"))
    L.ndo_render.argtypes = [C.c_char_p] + [C.c_int] * 5 + [C.c_void_p] * 6
    flat = load_flat("config4_balls5d")
    W, H = flat.header.width, flat.header.height
    n_frames, tiles = 2, 5
    band = (H + tiles - 1) // tiles
    items = [(f, t * band, min(band, H - t * band)) for f in range(n_frames) for t in range(tiles)]
    q = multi.TileQueue(dist, rank, world)
    frames = torch.zeros((n_frames, H, W, 4), dtype=torch.uint8) if rank == 0 else None
    stage = torch.zeros((len(items), band, W, 4), dtype=torch.uint8) if rank != 0 else None
    counts = []
    for step in range(2):
        mine = []
        for idx in q.pull(step, len(items)):
            f, y0, th = items[idx]
            t = oracle_render(L, flat, 0, y0, W, th, threads=1)
            if rank == 0:
                frames[f, y0:y0 + th] = torch.from_numpy(t.u8)
            else:
                stage[len(mine), :th] = torch.from_numpy(t.u8)
            mine.append(idx)
        owners = multi.gather_tiles(dist, rank, world, items, mine, stage, frames, band)
        assert sorted(owners) == list(range(len(items)))        # every item rendered exactly once
        counts.append(len(mine))
    if rank == 0:
        np.save(out_path, frames.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_queue_and_gather_equal_single_process(tmp_path, oracle_lib):
    out = str(tmp_path / "frames.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    frames = np.load(out)
    flat = load_flat("config4_balls5d")
    want = oracle_render(oracle_lib, flat).u8
    assert frames.shape[0] == 2
    for f in range(2):
        assert np.array_equal(frames[f], want)


def test_single_rank_queue_is_a_plain_range():
    from ndt_b200 import multi
    q = multi.TileQueue(None, 0, 1)
    assert list(q.pull(0, 7)) == list(range(7))


def _gather_worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from ndt_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    F, shape = 3, (5, 7, 4)
    g = multi.FrameGather(dist, rank, world, F)
    frames = torch.zeros((world * F,) + shape, dtype=torch.uint8) if rank == 0 else None
    for step in range(3):
        g.begin(frames)
        mine = [torch.full(shape, 10 * step + rank * F + j, dtype=torch.uint8) for j in range(F)]
        for j in (1, 0, 2):                 # frames finish out of order (two contexts per GPU); sends go out in order
            if rank == 0:
                frames[j] = mine[j]
            else:
                g.send(mine[j], j)
        g.end()
        if rank == 0:
            for f in range(world * F):
                assert int(frames[f].min()) == int(frames[f].max()) == 10 * step + f, (step, f)
    if rank == 0:
        np.save(out_path, frames.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_frames_are_sent_as_they_finish_and_land_in_order(tmp_path):
    """FrameGather (bench.py's N>1 gather): rank 0 posts a step's receives up front, the other rank sends frame j
    when it is done -- whatever order the frames finish in, frame f ends up in slot f."""
    out = str(tmp_path / "g.npy")
    mp.spawn(_gather_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    frames = np.load(out)
    assert [int(frames[f, 0, 0, 0]) for f in range(6)] == [20 + f for f in range(6)]
