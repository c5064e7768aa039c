"""Generate tests/golden/anim_balls5d/frame_NN.ndsf.gz: the flat scenes of frames 0..31 of BASELINE config 4
(`ndt -s scenes/balls.so -d 5 -r 4k`), produced by the UNMODIFIED reference (oracle/_ref) running scene_setup
for every frame IN ORDER -- scenes/balls.c carries its physics state from frame to frame (balls.c:27,181) --
then kd_tree_build and camera_aim (ndt.c:1791-1925), flattened by ndt_b200_flatten.  These are the INPUT of
bench.py --workload config4_anim and of tests/test_anim.py; the reference cannot travel to the GPU box.

    python tests/golden/make_anim.py        (build container, needs oracle/_ref)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import ndt_b200  # noqa: E402
from oracle.refharness import RefHarness  # noqa: E402

N_FRAMES, W, H, DIMS = 32, 3840, 2160, 5


def main():
    out = os.path.join(HERE, "anim_balls5d")
    os.makedirs(out, exist_ok=True)
    R = RefHarness()
    R.open_scene("balls")
    frames = R.scene_frames(DIMS, None)
    if frames <= 0:
        frames = 300
    total = 0
    for f in range(N_FRAMES):
        R.begin_frame(DIMS, f, frames, None)          # runs scene_setup of every frame up to f exactly once, in order
        flat = ndt_b200.flatten(R.scene_ptr, R.kdtree_ptr, W, H, 128, 1, R.get_bounds_ptr)
        R.end_frame()
        p = os.path.join(out, f"frame_{f:02d}.ndsf.gz")
        flat.save(p)
        total += os.path.getsize(p)
        print(f, len(flat), flat.header.n_items, flat.header.n_nodes, flush=True)
    print("total bytes", total)


if __name__ == "__main__":
    main()
