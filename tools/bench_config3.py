"""tools/bench_config3.py [n_objects] [w h] [max_growth] [max_depth] [leaf_size] -- BASELINE config 3 at scale: scenes/random.c in 6-D with n random
spheres / orthotopes at 4K.  The reference's kd builder does not terminate on this scene (SURVEY note 8), so the
tree comes from ndt_b200_kd_tree_build_bounded (a valid kd_tree_t in host memory, parity against the reference's
own trace_kd on that tree: tests/test_config3_scale.py).  Prints one JSON line: GPU frames/s and Mrays/s (CUDA
events, frame alone), build times, and the reference's render_image on the SAME tree at 1/64 of the pixels."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ndt_b200
import refharness

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
W = int(sys.argv[2]) if len(sys.argv) > 3 else 3840
H = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
GROWTH = float(sys.argv[4]) if len(sys.argv) > 4 else 1.5
DEPTH = int(sys.argv[5]) if len(sys.argv) > 5 else 16
LEAF = int(sys.argv[6]) if len(sys.argv) > 6 else 64
R = refharness.RefHarness()
R.open_scene("random")
t0 = time.perf_counter()
R.begin_frame_nokd(6, 0, 300, str(n))
t_scene = time.perf_counter() - t0
t0 = time.perf_counter()
rc = ndt_b200.kd_tree_build_bounded(R.kdtree_ptr, R.items_ptr, max_depth=DEPTH, leaf_size=LEAF, max_growth=GROWTH)
t_kd = time.perf_counter() - t0
flat = ndt_b200.flatten(R.scene_ptr, R.kdtree_ptr, W, H, 128, 1, R.get_bounds_ptr)
hd = flat.header
ms, st = [], None
with ndt_b200.Context(0) as ctx:
    ctx.upload(flat)
    import torch
    out = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda:0")
    for i in range(6):
        ctx.launch_tile(0, 0, W, H, d_u8=out.data_ptr())
        st = ctx.sync()
        if i:
            ms.append(st.device_ms)
    hit_px = None
    small = flat.retarget(W // 8, H // 8)
    ctx.upload(small)
    fr = ctx.render_tile(0, 0, W // 8, H // 8)
# the reference on the same tree, all host cores, 1/64 of the pixels (NDT_C3_NO_REF=1 skips the minute it takes)
sec = float("nan")
if not os.environ.get("NDT_C3_NO_REF"):
    _, sec = R.render(W // 8, H // 8, threads=os.cpu_count())
hit, oid, dist = R.primary(W // 8, H // 8)
R.end_frame()
gpu_ms = float(np.median(ms))
print(json.dumps({
    "workload": f"BASELINE config 3: scenes/random.c -d 6 -u {n}, {W}x{H}, frame 0 (bounded kd tree: {hd.n_nodes} nodes, "
                f"{hd.n_leaf_refs} leaf refs, largest leaf {hd.max_leaf}, depth {hd.tree_depth})",
    "max_growth": GROWTH, "max_depth": DEPTH, "leaf_size": LEAF, "scene_setup_and_bounds_s": t_scene, "bounded_kd_build_ms": t_kd * 1e3, "flat_bytes": len(flat),
    "gpu_frame_ms": gpu_ms, "frames_per_s": 1e3 / gpu_ms, "rays_per_frame": int(st.rays_unique),
    "mrays_per_s": st.rays_unique / gpu_ms / 1e3, "generations": int(st.generations),
    "reference_same_tree": {"cores": os.cpu_count(), "sample": f"{W//8}x{H//8} (1/64 of the pixels), render_image only",
                            "seconds": sec, "frames_per_s_full_res": 1.0 / (sec * 64)},
    "parity_at_sample_size": {"hit_equal": bool(np.array_equal(fr.hit, hit)), "id_equal": bool(np.array_equal(fr.obj_id, oid)),
                              "hit_pixels": int(hit.sum())},
}))
