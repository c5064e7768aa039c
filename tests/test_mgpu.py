"""All GPUs of one box behind the C ABI (ndt_b200_mgpu_*, mgpu.cu): a frame as row bands pulled from one
counter, an animation frame by frame, gathered in host buffers -- and byte-identical to the one-GPU render
(the reference's counterpart is mpi_collect_image, ndt.c:1277-1309, over disjoint rows)."""
import numpy as np
import pytest

import ndt_b200
from conftest import bits_equal, load_flat

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def single(flat):
    with ndt_b200.Context(0) as c:
        c.upload(flat)
        return c.render_tile(0, 0, flat.header.width, flat.header.height)


def same_frame(a, b):
    return (bits_equal(a.rgba_f64, b.rgba_f64) and bits_equal(a.rgba_u8, b.rgba_u8) and np.array_equal(a.hit, b.hit)
            and np.array_equal(a.obj_id, b.obj_id) and bits_equal(a.inv_depth, b.inv_depth))


@pytest.mark.parametrize("devices", [1, 2, 0], ids=["1gpu", "2gpu", "all"])
@pytest.mark.parametrize("key,band", [("config1_default4d", 8), ("config2_hypercube8d", 0), ("default5d_odd", 7)])
def test_a_frame_split_over_contexts_equals_the_single_render(devices, key, band):
    if devices == 2 and n_gpus() < 2:
        pytest.skip("needs two GPUs")
    flat = load_flat(key)
    want = single(flat)
    with ndt_b200.MultiGpu(devices) as m:
        assert m.devices == (devices or n_gpus())
        got = m.render_frame(flat, band_rows=band)
        assert same_frame(got, want)
        assert got.stats.rays_unique == want.stats.rays_unique
        again = m.render_frame(flat, band_rows=band)        # the contexts are reusable
        assert same_frame(again, want)


def test_an_animation_streams_through_the_queue():
    keys = ["config4_balls5d", "config1_default4d", "mixed7d", "config4_balls5d", "default3d", "config5_mixed10d"] * 3
    flats = {k: load_flat(k) for k in set(keys)}
    want = {k: single(f) for k, f in flats.items()}
    with ndt_b200.MultiGpu(0) as m:
        outs = []
        for k in keys:
            h = flats[k].header
            u8 = np.zeros((h.height, h.width, 4), np.uint8)
            f64 = np.zeros((h.height, h.width, 4), np.float64)
            m.submit(flats[k], u8, f64)
            outs.append((k, u8, f64))
        st = m.wait()
        for k, u8, f64 in outs:
            assert bits_equal(u8, want[k].rgba_u8) and bits_equal(f64, want[k].rgba_f64), k
        assert st.rays_unique == sum(want[k].stats.rays_unique for k in keys)
        # a frame split is refused while frames stream, and works again afterwards
        got = m.render_frame(flats["default3d"])
        assert same_frame(got, want["default3d"])


def test_errors_surface_from_the_workers():
    with ndt_b200.MultiGpu(1) as m:
        flat = load_flat("default3d")
        bad = bytearray(flat.blob)
        bad[0] ^= 0xFF                                      # magic
        with pytest.raises(ndt_b200.NdtB200Error):
            ndt_b200._check(ndt_b200.lib().ndt_b200_mgpu_submit(m._h, bytes(bad), None, None))
        m.wait()
    with pytest.raises(ndt_b200.NdtB200Error):
        ndt_b200.MultiGpu(99)
