"""perf_flat.py <flat scene file> [w h] [reps] -- device time of one frame of a flat scene read from a file
(tools/perf_frame.py for scenes that are not among the committed fixtures, e.g. BASELINE config 3 at 10 000 objects)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ndt_b200
flat = ndt_b200.FlatScene.load(sys.argv[1])
if len(sys.argv) > 3:
    flat = flat.retarget(int(sys.argv[2]), int(sys.argv[3]))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
W, H = flat.header.width, flat.header.height
ctx = ndt_b200.Context(0)
ctx.upload(flat)
import torch
out = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda:0")
torch.cuda.synchronize()
ms = []
for i in range(reps + 1):
    ctx.launch_tile(0, 0, W, H, d_u8=out.data_ptr())
    st = ctx.sync()
    if i:
        ms.append(st.device_ms)
print(f"{os.environ.get('NDT_B200_LIB','default')[-40:]:42s} {os.path.basename(sys.argv[1])} {W}x{H}: min {min(ms):9.3f} ms  median {np.median(ms):9.3f} ms  "
      f"rays {st.rays_unique} gens {st.generations} checksum {int(out.to(torch.int64).sum())}", flush=True)
