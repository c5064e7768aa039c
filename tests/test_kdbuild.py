"""ndt_b200_kd_tree_build is a drop-in for kd_tree_build (kd-tree.c:421-477): built
over the reference's own kd_item_list_t it must give the SAME kd_tree_t --
checked by flattening the scene once with the reference-built tree and once
with ours and comparing the blobs byte for byte (nodes, split planes, leaf
order, infinite list, root AABB all live in the blob)."""
import time

import pytest

import ndt_b200
from scenes import BY_KEY

pytestmark = pytest.mark.gpu

KEYS = ["config1_default4d", "config4_balls5d", "hypercube_points6d", "hypercube6d_walls",
        "config5_mixed10d", "config3_random6d", "config2_hypercube8d"]


@pytest.mark.parametrize("key", KEYS)
def test_gpu_kd_build_equals_reference_tree(key, ref):
    c = BY_KEY[key]
    ref.open_scene(c.scene)
    frames = ref.scene_frames(c.dims, c.cfg) if c.scene else 300
    ref.begin_frame(c.dims, c.frame, frames if frames > 0 else 300, c.cfg)
    try:
        # Build FIRST, at the point where main() calls kd_tree_build (ndt.c:1908): the bounds of
        # cluster children are still lazy (radius == 0) there, and the builder's finite/infinite
        # classification reads them (kd-tree.c:433,385) -- flattening forces them.
        ndt_b200.kd_tree_build(ref.alt_tree_begin(), ref.items_ptr)      # warm-up: CUDA context, scratch
        alt = ref.alt_tree_begin()
        t0 = time.perf_counter()
        rc = ndt_b200.kd_tree_build(alt, ref.items_ptr)
        ours = time.perf_counter() - t0
        assert rc in (0, 1)
        want = ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, c.w, c.h, 128, 1, ref.get_bounds_ptr)
        got = ndt_b200.flatten(ref.scene_ptr, alt, c.w, c.h, 128, 1, ref.get_bounds_ptr)
        print(f"\n{key}: {want.header.n_items} items, {want.header.n_nodes} nodes, {want.header.n_leaf_refs} leaf refs: "
              f"reference kd_tree_build {ref.kd_seconds*1e3:.1f} ms, ndt_b200_kd_tree_build {ours*1e3:.1f} ms")
        assert got.header.n_nodes == want.header.n_nodes
        assert got.header.n_leaf_refs == want.header.n_leaf_refs
        assert got.blob == want.blob
    finally:
        ref.end_frame()
