"""BASELINE config 3 (scenes/random.c, 6-D, thousands of overlapping spheres / orthotopes) at a size the
reference's kd builder cannot finish (SURVEY note 8, section 8d): the tree comes from
ndt_b200_kd_tree_build_bounded -- a valid kd_tree_t in host memory -- and the device must answer every
query exactly like the reference's own trace_kd (object.c:683) walking THAT tree.  How often that
answer differs from the reference's tree-less trace() (object.c:692) over all objects is reported too:
kd_tree_intersect prunes with *t_ptr and per-leaf EPSILON hysteresis, so for rays that start inside
the cloud the two reference code paths do not agree with each other either (not a device property)."""
import time

import numpy as np
import pytest

import ndt_b200

pytestmark = pytest.mark.gpu
W, H = 160, 90


def primary_rays(flat, step):
    """(o, v) of every step-th pixel, camera.c:557-575 on the flat scene's camera block."""
    h = flat.header
    n, npad = h.n, h.npad
    cam = np.frombuffer(flat.blob, np.float64, 4 * npad, h.off_camera).reshape(4, npad)[:, :n]
    pos, orig, dx, dy = cam
    rays_o, rays_v, pix = [], [], []
    for j in range(0, h.height, step):
        for i in range(0, h.width, step):
            x = i / h.width - 0.5
            y = -(j / h.height - 0.5)
            p = orig + dx * x + dy * y
            if h.use_focal:
                p = pos + (p - pos) * h.focal_scale
            v = p - pos
            rays_o.append(pos.copy()); rays_v.append(v / np.sqrt((v * v).sum())); pix.append((j, i))
    return np.array(rays_o), np.array(rays_v), pix


# (objects, max_growth, pixel step of the sub-sampled grid, extra rays aimed at objects).  The last row is BASELINE
# config 3 as stated -- ~10 k objects: a tree the reference cannot build (SURVEY note 8); its own trace_kd walking
# the bounded tree is the expected answer, ray for ray
CONFIG3_CASES = [(1500, 1.5, 3, 1500), (1500, 1.95, 3, 1500), (10000, 1.95, 9, 700)]


@pytest.mark.parametrize("N_OBJECTS,growth,step,n_extra", CONFIG3_CASES, ids=["n1500_g1.5", "n1500_g1.95", "n10000_g1.95"])
def test_random_scene_with_a_bounded_tree_against_brute_force(ref, N_OBJECTS, growth, step, n_extra):
    ref.open_scene("random")
    t0 = time.perf_counter()
    ref.begin_frame_nokd(6, 0, 300, str(N_OBJECTS))
    t_scene = time.perf_counter() - t0
    try:
        t0 = time.perf_counter()
        rc = ndt_b200.kd_tree_build_bounded(ref.kdtree_ptr, ref.items_ptr, max_depth=14, leaf_size=48, max_growth=growth)
        t_kd = time.perf_counter() - t0
        assert rc in (0, 1)
        flat = ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, W, H, 128, 1, ref.get_bounds_ptr)
        hd = flat.header
        assert hd.n_items == N_OBJECTS
        o, v, pix = primary_rays(flat, step)
        n_prim = len(o)
        # the camera of random.c sees little of the cloud (its 6-D objects rarely cut the 3-D view): add rays
        # aimed at the objects, from the camera and from inside the cloud
        rng = np.random.default_rng(3)
        bs = np.frombuffer(flat.blob, np.float64, hd.n_objects * (hd.npad + 2), hd.off_bspheres).reshape(hd.n_objects, hd.npad + 2)
        cam_pos = o[0].copy()
        eo, ev = [], []
        for k in range(n_extra):
            i = rng.integers(hd.n_items)
            src = cam_pos if k % 2 == 0 else rng.uniform(2.0, 12.0, size=hd.n)
            tgt = bs[i, :hd.n] + rng.normal(size=hd.n) * abs(bs[i, hd.npad]) * 0.4
            d = tgt - src
            eo.append(src.copy()); ev.append(d / np.sqrt((d * d).sum()))
        o = np.vstack([o, np.array(eo)]); v = np.vstack([v, np.array(ev)])
        brute = [ref.trace_brute(o[k], v[k]) for k in range(len(o))]
        kd = [ref.trace_ray(o[k], v[k]) for k in range(len(o))]       # trace_kd on the bounded tree
    finally:
        ref.end_frame()
    with ndt_b200.Context(0) as ctx:
        ctx.upload(flat)
        frame = ctx.render_tile(0, 0, W, H)
        found, oid, t, hit, nrm = ctx.trace_rays(o, v)
    want_id = np.array([k_[3] for k_ in kd])
    want_found = np.array([1 if k_[0] else 0 for k_ in kd])
    diff = np.flatnonzero((oid != want_id) | ((found != 0) != (want_found != 0)))
    same = np.flatnonzero((oid == want_id) & (oid >= 0))
    hit_err = max((float(np.abs(hit[k] - kd[k][1]).max()) for k in same), default=0.0)
    brute_id = np.array([b[2] for b in brute])
    kd_vs_brute = int((brute_id != want_id).sum())
    n_hits = int((want_id >= 0).sum())
    print(f"\\nrandom.c 6-D, {N_OBJECTS} objects, max_growth {growth}: scene + bounds {t_scene:.1f} s, bounded kd build {t_kd*1e3:.0f} ms "
          f"({hd.n_nodes} nodes, {hd.n_leaf_refs} leaf refs, largest leaf {hd.max_leaf}, depth {hd.tree_depth}); "
          f"{len(o)} rays ({n_hits} hit something): {len(diff)} differ from the reference's trace_kd on the same tree, hit points within {hit_err:.2e}; "
          f"the reference's trace_kd and its tree-less trace() disagree on {kd_vs_brute}; "
          f"frame {W}x{H}: {frame.stats.rays_unique} rays in {frame.stats.device_ms:.2f} ms")
    assert len(diff) == 0, (len(diff), len(o))
    assert hit_err == 0.0
    # the frame's primary-ray buffers agree with the probe
    got_id = np.array([frame.obj_id[j, i] for j, i in pix])
    assert np.array_equal(got_id, oid[:n_prim])
    assert n_hits > min(500, n_extra // 2)


def test_staged_face_lists_equal_the_scalar_face_loop(ref, monkeypatch):
    """The face list nested in an hcube (472 orthotopes for a 6-cube) goes through the warp's second staging area
    (warp.cuh: warp_nested -- box / bundle culls, staged narrow phase) when many lanes ask for the cube, and
    across the lanes for one ray at a time when few do (hcube_one_ray); NDT_B200_NO_NESTED_STAGE=1 at upload
    keeps round 1's scalar loop (core.cuh: trace_list).  Same answers bit for bit on coherent bundles aimed at
    the cubes, on incoherent rays from inside the cloud, and on a frame."""
    n_obj = 400
    ref.open_scene("random")
    ref.begin_frame_nokd(6, 0, 300, str(n_obj))
    try:
        rc = ndt_b200.kd_tree_build_bounded(ref.kdtree_ptr, ref.items_ptr, max_depth=14, leaf_size=48, max_growth=1.5)
        assert rc in (0, 1)
        flat = ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, W, H, 128, 1, ref.get_bounds_ptr)
    finally:
        ref.end_frame()
    hd = flat.header
    obj_type = np.frombuffer(flat.blob, np.int32, hd.n_objects * 24, hd.off_objects).reshape(hd.n_objects, 24)[:, 0]
    cubes = np.flatnonzero(obj_type[:hd.n_items] == 4)
    assert len(cubes) > 10 and hd.n_objects > hd.n_items
    bs = np.frombuffer(flat.blob, np.float64, hd.n_objects * (hd.npad + 2), hd.off_bspheres).reshape(hd.n_objects, hd.npad + 2)
    rng = np.random.default_rng(11)
    o, v = [], []
    for k in range(60):                     # bundles of 32 nearly parallel rays through a cube: every lane asks for it
        c = bs[cubes[k % len(cubes)], :hd.n]
        src = rng.uniform(-20.0, 30.0, size=hd.n)
        for _ in range(32):
            d = c + rng.normal(size=hd.n) * 0.05 - src
            o.append(src.copy()); v.append(d / np.sqrt((d * d).sum()))
    for k in range(3000):                   # incoherent: from inside the cloud at an object
        i = rng.integers(hd.n_items)
        src = rng.uniform(2.0, 12.0, size=hd.n)
        d = bs[i, :hd.n] + rng.normal(size=hd.n) * abs(bs[i, hd.npad]) * 0.4 - src
        o.append(src); v.append(d / np.sqrt((d * d).sum()))
    o = np.array(o); v = np.array(v)
    a = ndt_b200.Context(0)
    b = ndt_b200.Context(0)
    try:
        a.upload(flat)
        monkeypatch.setenv("NDT_B200_NO_NESTED_STAGE", "1")
        b.upload(flat)
        monkeypatch.delenv("NDT_B200_NO_NESTED_STAGE")
        ra = a.trace_rays(o, v)
        rb = b.trace_rays(o, v)
        fa = a.render_tile(0, 0, W, H)
        fb = b.render_tile(0, 0, W, H)
    finally:
        a.close(); b.close()
    for x, y in zip(ra, rb):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    # (random.c's cubes have random, non-orthogonal edge directions: with orthotope.c's per-axis projections their
    # faces all but never report a hit -- the face lists are walked, nothing is accepted; the cube of
    # scenes/hypercube.c below is hit)
    assert int((ra[1] >= 0).sum()) > 1000
    assert np.array_equal(fa.obj_id, fb.obj_id)
    assert np.array_equal(fa.rgba_f64.view(np.uint8), fb.rgba_f64.view(np.uint8))


@pytest.mark.parametrize("key,w,h", [("hypercube5d_hcube", 640, 360), ("view_vr5d", 96, 54)])
def test_staged_face_lists_on_a_cube_that_is_hit(key, w, h, monkeypatch):
    """scenes/hypercube.c -u hcube: one 5-cube (130 faces of 2, 3 and 4 dimensions) in front of the camera.  The scene is too small for the box culls
    to be switched on; NDT_B200_FORCE_BOXES=1 at upload switches them on, and with them the staged face lists.
    The frame must not change (the plain render is checked against the oracle in test_gpu_parity.py)."""
    from conftest import load_flat
    flat = load_flat(key)
    if (w, h) != (flat.header.width, flat.header.height):
        flat = flat.retarget(w, h)
    a = ndt_b200.Context(0)
    b = ndt_b200.Context(0)
    try:
        a.upload(flat)
        monkeypatch.setenv("NDT_B200_FORCE_BOXES", "1")
        b.upload(flat)
        monkeypatch.delenv("NDT_B200_FORCE_BOXES")
        fa = a.render_tile(0, 0, w, h)
        fb = b.render_tile(0, 0, w, h)
    finally:
        a.close(); b.close()
    hd = flat.header
    obj_type = np.frombuffer(flat.blob, np.int32, hd.n_objects * 24, hd.off_objects).reshape(hd.n_objects, 24)[:, 0]
    cubes = np.flatnonzero(obj_type[:hd.n_items] == 4)
    bs = np.frombuffer(flat.blob, np.float64, hd.n_objects * (hd.npad + 2), hd.off_bspheres).reshape(hd.n_objects, hd.npad + 2)
    assert len(cubes) >= 1 and bs[cubes[0], hd.npad] > 0          # a bounding sphere: the cube gets a box, the scene the culls
    assert int(np.isin(fa.obj_id, cubes).sum()) >= 16
    assert np.array_equal(fa.hit, fb.hit) and np.array_equal(fa.obj_id, fb.obj_id)
    assert np.array_equal(fa.rgba_f64.view(np.uint8), fb.rgba_f64.view(np.uint8))
    assert fb.stats.launches != 0 and fa.stats.rays_unique == fb.stats.rays_unique
