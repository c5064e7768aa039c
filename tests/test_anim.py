"""BASELINE config 4 as stated: scenes/balls.c -d 5 at 4K as an ANIMATION.  The scene carries its physics from
frame to frame (balls.c:27,181), so the flat scenes of frames 0..31 were produced by the reference running
scene_setup in order (tests/golden/make_anim.py).  Three sampled frames at the full 3840x2160 against the oracle
on sample tiles, and the whole sequence streamed through the multi-GPU queue against one-at-a-time renders."""
import glob
import os

import numpy as np
import pytest

import ndt_b200
from conftest import GOLDEN, bits_equal, oracle_render
from test_gpu_parity import LSB_OK_FRACTION, colour_report, pick_tiles

FRAMES = sorted(glob.glob(os.path.join(GOLDEN, "anim_balls5d", "frame_*.ndsf.gz")))


def test_the_fixture_is_the_sequence():
    assert len(FRAMES) == 32
    a = ndt_b200.FlatScene.load(FRAMES[2]).retarget(96, 54)
    b = ndt_b200.FlatScene.load(os.path.join(GOLDEN, "config4_balls5d.ndsf.gz"))
    assert a.blob == b.blob                 # frame 2 is the single-frame fixture of config 4
    sizes = {ndt_b200.FlatScene.load(p).header.n_items for p in FRAMES}
    assert sizes == {132}                   # 100 balls + 16 corner spheres + 15 edge cylinders + the ground
    blobs = {ndt_b200.FlatScene.load(p).blob for p in FRAMES}
    assert len(blobs) == 32                 # the balls move: no two frames are the same scene


@pytest.mark.gpu
@pytest.mark.parametrize("f", [0, 15, 31])
def test_sampled_frames_at_4k_against_oracle_tiles(f, oracle_lib):
    flat = ndt_b200.FlatScene.load(FRAMES[f])
    w, h = flat.header.width, flat.header.height
    assert (w, h) == (3840, 2160)
    with ndt_b200.Context(0) as ctx:
        ctx.upload(flat)
        got = ctx.render_tile(0, 0, w, h, want=("u8", "hit", "id", "depth"))
    tiles = pick_tiles(got.hit, got.obj_id)[:6]
    for (x0, y0) in tiles:
        want = oracle_render(oracle_lib, flat, x0=x0, y0=y0, tw=64, th=64)
        sl = (slice(y0, y0 + 64), slice(x0, x0 + 64))
        assert np.array_equal(got.hit[sl], want.hit) and np.array_equal(got.obj_id[sl], want.id), (f, x0, y0)
        assert bits_equal(got.inv_depth[sl], want.depth), (f, x0, y0)
        ok, dmax, nbad, exact = colour_report(got.rgba_u8[sl], want.u8)
        assert ok >= LSB_OK_FRACTION, (f, x0, y0, ok)


@pytest.mark.gpu
def test_the_animation_streams_over_all_gpus():
    """ndt_b200_mgpu_submit: the host hands the frames over in order, whichever context is free renders them; every
    frame equals its one-at-a-time render (at a reduced size: 32 frames of 4K would be 1 GB of host buffers)."""
    flats = [ndt_b200.FlatScene.load(p).retarget(480, 270) for p in FRAMES]
    with ndt_b200.Context(0) as ctx:
        want = []
        for fl in flats:
            ctx.upload(fl)
            want.append(ctx.render_tile(0, 0, 480, 270, want=("u8",)).rgba_u8.copy())
    with ndt_b200.MultiGpu(0) as m:
        outs = [np.zeros((270, 480, 4), np.uint8) for _ in flats]
        for fl, o in zip(flats, outs):
            m.submit(fl, o)
        m.wait()
    for k, (o, w_) in enumerate(zip(outs, want)):
        assert bits_equal(o, w_), k
    assert not bits_equal(outs[0], outs[31])
