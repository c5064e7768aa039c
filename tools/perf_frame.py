"""perf_frame.py [workload] -- device time of one full frame, for A/B-ing library variants
(NDT_B200_LIB=path selects the .so).  Prints min/median kernel ms over a few repeats."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ndt_b200
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
key, W, H, *_ = bench.WORKLOADS[wl]
flat = bench.load_flats(wl)[int(os.environ.get('NDT_FRAME', '0'))]
ctx = ndt_b200.Context(0)
ctx.upload(flat)
if os.environ.get('NDT_OPTS'):
    ctx.set_options(int(os.environ['NDT_OPTS']))
import torch
out = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda:0")
torch.cuda.synchronize()
ms = []
for i in range(reps + 1):
    ctx.launch_tile(0, 0, W, H, d_u8=out.data_ptr())
    st = ctx.sync()
    if i:
        ms.append(st.device_ms)
print(f"{os.environ.get('NDT_B200_LIB','default')[-40:] + ' opts=' + os.environ.get('NDT_OPTS','0'):48s} {wl}: min {min(ms):8.3f} ms  median {np.median(ms):8.3f} ms  "
      f"rays {st.rays_unique}  gens {st.generations}  -> {st.rays_unique/min(ms)/1e3:.1f} Mrays/s", flush=True)
