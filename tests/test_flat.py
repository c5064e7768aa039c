"""Flat-scene container: validation rejects malformed blobs, retarget keeps the
camera contract, save/load round-trips."""
import ctypes as C

import pytest

import ndt_b200
from conftest import load_flat


def test_roundtrip_and_header(tmp_path, golden):
    f = load_flat("config2_hypercube8d")
    g = golden["config2_hypercube8d"]
    assert f.header.n == 8 and f.header.npad == 8
    assert f.header.n_items == g["n_items"] == 6561          # SURVEY.md section 8a / BASELINE.md
    assert f.header.n_nodes == 513 and f.header.n_leaf_refs == 58224 and f.header.max_leaf == 227
    p = tmp_path / "x.ndsf.gz"
    f.save(p)
    assert ndt_b200.FlatScene.load(p).blob == f.blob


def test_odd_dimension_is_padded():
    f = load_flat("default5d_odd")
    assert f.header.n == 5 and f.header.npad == 6


@pytest.mark.parametrize("mutate", ["truncate", "magic", "node", "leaf", "object", "light_type", "light_vec", "geom_extent",
                                    "node_cycle", "tree_depth", "max_leaf"])
def test_validate_rejects_corruption(mutate):
    f = load_flat("config4_balls5d")
    b = bytearray(f.blob)
    h = f.header
    if mutate == "truncate":
        b = b[:-16]
    elif mutate == "magic":
        b[0] ^= 0xFF
    elif mutate == "node":
        C.c_int32.from_buffer(b, h.off_nodes + 4).value = 10 ** 6      # left child out of range
    elif mutate == "leaf":
        C.c_int32.from_buffer(b, h.off_leaf_refs).value = h.n_items    # id out of range
    elif mutate == "object":
        C.c_int32.from_buffer(b, h.off_objects).value = 99             # unknown type
    elif mutate == "light_type":
        C.c_int32.from_buffer(b, h.off_lights).value = 7               # ndt_flat_light.type outside scene.h:17-23
    elif mutate == "light_vec":                                        # ndt_flat_light.vec_off (offset 48): 4 npad doubles must fit geom[]
        C.c_uint32.from_buffer(b, h.off_lights + 48).value = h.n_geom - 2
    elif mutate == "geom_extent":                                      # ndt_flat_object.geom_off (offset 24): the block must fit geom[]
        C.c_uint32.from_buffer(b, h.off_objects + 24).value = (h.n_geom - 2) & ~1
    elif mutate == "node_cycle":                                       # a child that is not behind its parent (pre-order)
        C.c_int32.from_buffer(b, h.off_nodes + 4).value = 0
    elif mutate == "tree_depth":                                       # header says shallower than the nodes are
        hdr = ndt_b200.FlatHeader.from_buffer(b)
        hdr.tree_depth = 1
    elif mutate == "max_leaf":
        hdr = ndt_b200.FlatHeader.from_buffer(b)
        hdr.max_leaf = 1
    with pytest.raises(ndt_b200.NdtB200Error):
        ndt_b200.FlatScene(bytes(b))


def test_retarget_needs_same_aspect():
    f = load_flat("config1_default4d")
    big = f.retarget(1920, 1080)
    assert big.header.width == 1920 and big.blob[C.sizeof(ndt_b200.FlatHeader):] == f.blob[C.sizeof(ndt_b200.FlatHeader):]
    with pytest.raises(ValueError):
        f.retarget(1000, 1000)


def test_flatten_refuses_what_the_device_cannot_do(ref):
    """Unknown camera types / area lights / a VR stereo view without the host's vectNd_rotate2 are hard
    errors, not CPU fallbacks."""
    import numpy as np
    ref.open_scene(None)
    ref.begin_frame(4, 0, 300, None)
    try:
        # camera.type is the first int of scene.cam (offset 16); camera.h:16-20 knows 0..2
        cam_type = C.c_int.from_address(ref.scene_ptr + 16)
        cam_type.value = 7
        with pytest.raises(ndt_b200.NdtB200Error) as e:
            ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, 64, 36, 128, 1, ref.get_bounds_ptr)
        assert e.value.code == -2
        cam_type.value = 1          # CAMERA_VR side by side: the eye is rotated per column by the host's vectNd_rotate2
        with pytest.raises(ndt_b200.NdtB200Error) as e:
            ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, 64, 36, 128, 1, ref.get_bounds_ptr,
                             stereo_mode=ndt_b200.SIDE_SIDE_3D)
        assert e.value.code == -1
        with pytest.raises(ndt_b200.NdtB200Error) as e:
            ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, 64, 36, 128, 1, ref.get_bounds_ptr, stereo_mode=9)
        assert e.value.code == -2
        cam_type.value = 0
        # first light -> LIGHT_DISK (4): lights[0]->type at offset 248
        lights = C.c_void_p.from_address(ref.scene_ptr + 704).value
        l0 = C.c_void_p.from_address(lights).value
        lt = C.c_int.from_address(l0 + 248)
        old = lt.value
        lt.value = 4
        with pytest.raises(ndt_b200.NdtB200Error) as e:
            ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, 64, 36, 128, 1, ref.get_bounds_ptr)
        assert e.value.code == -2
        lt.value = old
        ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, 64, 36, 128, 1, ref.get_bounds_ptr)
    finally:
        ref.end_frame()
