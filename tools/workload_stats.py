"""tools/workload_stats.py <golden key> [w h] -- per-ray work statistics of a flat scene
(CPU emulation of the device core with the NDT_STAT hooks of core.cuh)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ndt_b200
so = "/tmp/libndt_emustats.so"
subprocess.run(["g++", "-m64", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-I" + ROOT + "/include",
                "-I" + ROOT + "/ndt_b200/csrc", "-shared", "-o", so, ROOT + "/tools/workload_stats.cpp", "-lm"], check=True)
L = C.CDLL(so)
key = sys.argv[1]
f = ndt_b200.FlatScene.load(key if os.path.exists(key) else os.path.join(ROOT, "tests", "golden", key + ".ndsf.gz"))
if len(sys.argv) > 3:
    f = f.retarget(int(sys.argv[2]), int(sys.argv[3]))
w, h = f.header.width, f.header.height
L.emu_render.argtypes = [C.c_char_p] + [C.c_int] * 4 + [C.c_void_p] * 6
st = (C.c_uint64 * 8)()
L.emu_stats_reset()
L.emu_render(f.blob, 0, 0, w, h, None, None, None, None, None, st)
n = L.emu_stats_words()
out = (C.c_ulonglong * n)()
L.emu_stats_get(out)
names = ["trace_kd", "aabb_hit", "nodes", "leaf_visits", "leaf_objs", "mb_skip", "bs_test", "bs_pass"] + \
        ["prim%d" % i for i in range(16)] + ["prim_hit", "accept", "hc_child", "hc_bs_pass", "hc_hit"]
d = dict(zip(names, list(out)))
rays = st[0] + st[1] + st[2]
print("frame %dx%d rays primary %d bounce %d shadow %d flops %d" % (w, h, st[0], st[1], st[2], st[5]))
for k, v in d.items():
    if v:
        print("%-12s %12d  per ray %8.2f  per aabb-hit ray %8.2f" % (k, v, v / rays, v / max(1, d["aabb_hit"])))
