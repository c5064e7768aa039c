#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric: Mrays/s = unique nearest-hit queries (trace_kd calls, object.c:683: primary +
reflection/refraction + shadow) per second; frames/s is reported next to it.  Workload:
BASELINE config 2 -- scenes/hypercube.c at 8 dimensions, 1920x1080, reflections on, frame 0
-- which is the configuration the metric is quoted on and fits one GPU.  The scene enters as
the flat blob the struct-ABI adapter produced from the reference's own scene + kd-tree
(tests/golden/, made by tests/golden/make_golden.py), so nothing under oracle/ runs on our arm.

A STEP renders FRAMES_PER_GPU frames per GPU ("weak": per-GPU work is fixed as N grows).  Each
rank renders its own frames with IN_FLIGHT contexts (four frames in flight per GPU, pulled from a
rank-local queue); the only exchange is "finished frames -> rank 0":

  value  device-resident: scene already in HBM, every frame ends up in rank 0's HBM.  N>1: each
         frame is sent with NCCL as soon as it is rendered, while the next one renders; rank 0 posts
         the receives up front (ndt_b200/multi.py).
  e2e    through the C ABI with HOST buffers: ndt_b200_upload (scene H2D) + ndt_b200_render_tile
         (render + D2H) per frame.  N>1: every rank writes its frames into ONE page-locked host
         buffer shared by the ranks of the box (POSIX shared memory), over its own PCIe link -- the
         host-side analogue of mpi_collect_image (ndt.c:1277-1309); rank 0 then holds every frame.
  multi_gpu_equal  after the timed region rank 0 renders every distinct frame alone and compares:
         every gathered frame (HBM and host) must be byte-identical.
  roofline  FP64 pipe.  `frac` = EXECUTED DADD + DMUL + DFMA thread instructions of a frame (ncu,
         profiles/r02_fp64_ops_<workload>.json, same build) / the CUDA-event time of one frame alone /
         the non-fused peak measured live (parity forbids FMA contraction, so one DMUL or DADD per lane
         slot is the pipe's real rate for this code).  `frac_algorithmic` keeps round 1's definition
         (flops of the REFERENCE algorithm per unique ray, counting build) for continuity.
  e2e_plugin  N=1: wall time per frame of the STOCK ndt command line with render_image and
         kd_tree_build bound to this library (integration/ndt_b200_demo): scene_setup, kd build,
         flatten, upload, render, read-back, image file.
  cpu_baseline / --impl reference: the UNMODIFIED reference's render_image (oracle/_ref) on all host
         cores, on a bounded sample of the same workload (same scene and frame at 1/16 of the pixels).
"""
import argparse
import ctypes as C
import glob
import gzip
import json
import os
import queue
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # key: (golden flat scene, width, height, description, reference scene plugin, dims, cfg, frame)
    "config2": ("config2_hypercube8d", 1920, 1080,
                "BASELINE config 2: scenes/hypercube.c -d 8, 1920x1080, reflections on, frame 0 "
                "(6561 objects: sphere/cylinder/orthotope/hcylinder + hplane floor, kd 513 nodes)",
                "hypercube", 8, None, 0),
    "config1": ("config1_default4d", 1920, 1080,
                "BASELINE config 1: built-in scene -d 4, 1920x1080, frame 0",
                None, 4, None, 0),
    "config4": ("config4_balls5d", 3840, 2160,
                "BASELINE config 4: scenes/balls.c -d 5, 4K, frame 2", "balls", 5, None, 2),
    "config4_anim": ("anim_balls5d", 3840, 2160,
                     "BASELINE config 4 as stated: scenes/balls.c -d 5, 4K, the animation's frames 0..31 (scene_setup run "
                     "in order by the reference, tests/golden/make_anim.py), frames sharded over the GPUs", "balls", 5, None, 2),
    "config5": ("config5_mixed10d", 1920, 1080,
                "BASELINE config 5 (C twin): mixed10d -d 10, 1920x1080, frame 0", "mixed10d", 10, None, 0),
    "config5_yaml": ("config5_yaml10d", 1920, 1080,
                     "BASELINE config 5: scenes/yaml.c -u tests/scenes/config5_mixed10d.yaml -d 10, 1920x1080 "
                     "(hplane, hcylinder, hdisk, hfacet; 6 lights, shadow rays), loaded by the reference's "
                     "scene_read_yaml over yaml_lite", "yaml", 10, "tests/scenes/config5_mixed10d.yaml", 0),
}
FRAMES_PER_GPU = 4
ANIM_FRAMES = 32
IN_FLIGHT = int(os.environ.get("NDT_IN_FLIGHT", "4"))   # frames in flight per GPU: one ndt_b200 context (own CUDA stream) and one host thread each; 4 = the whole step: the latency-bound late generations of one frame (config 1: ten generations of < 20 000 rays, 0.19 ms each) run under the bulk of the others (2 in flight: config 1 -7 %, config 2 -1 %)
SAMPLE_DIV = 4          # reference sample: width/4 x height/4 = 1/16 of the pixels


def read_blob(key):
    with gzip.open(os.path.join(ROOT, "tests", "golden", key + ".ndsf.gz"), "rb") as f:
        return f.read()


def load_flats(workload):
    """the flat scenes of the workload's frames (one for the single-frame workloads)"""
    import ndt_b200
    key, W, H = WORKLOADS[workload][:3]
    if workload == "config4_anim":
        return [ndt_b200.FlatScene.load(p) for p in
                sorted(glob.glob(os.path.join(ROOT, "tests", "golden", key, "frame_*.ndsf.gz")))[:ANIM_FRAMES]]
    return [ndt_b200.FlatScene.load(os.path.join(ROOT, "tests", "golden", key + ".ndsf.gz")).retarget(W, H)]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference on the host cores
# --------------------------------------------------------------------------------------
def reference_sample(workload, steps, warmup):
    """Times render_image (ndt.c:900) of oracle/_ref on the bounded sample; returns a dict.
    Nothing of ndt_b200 is imported here: the flat scene the oracle port counts rays on is read as bytes."""
    from oracle import refharness
    key, W, H, desc, scene, dims, cfg, frame = WORKLOADS[workload]
    sw, sh = W // SAMPLE_DIV, H // SAMPLE_DIV
    cores = os.cpu_count()
    if not refharness.available():
        raise RuntimeError("oracle/_ref missing: run `make -C oracle ref` in the build container")
    R = refharness.RefHarness()
    R.open_scene(scene)
    frames = R.scene_frames(dims, cfg) if scene else 300
    times = []
    R.begin_frame(dims, frame, frames if frames > 0 else 300, cfg)     # scene + kd build once (above the hot path)
    try:
        for i in range(warmup + steps):
            if i:
                R.reaim()           # render_image rescales cam.dirX in place (ndt.c:926)
            _, sec = R.render(sw, sh, threads=cores)
            if i >= warmup:
                times.append(sec)
    finally:
        R.end_frame()
    # unique / as-executed ray counts of the sample, from the oracle port (one trace per pixel)
    L = C.CDLL(os.path.join(ROOT, "oracle", "libndt_oracle.so"))
    L.ndo_render.argtypes = [C.c_char_p] + [C.c_int] * 5 + [C.c_void_p] * 6
    blob = bytearray(read_blob("anim_balls5d/frame_02" if workload == "config4_anim" else key))
    # the same scene for the sample's frame size: width / height live in the header (ndt_flat.h), equal aspect ratio
    hdr = np.frombuffer(blob, np.int32, 8, 0)     # magic, version, total_bytes (8), n, npad, width, height
    hdr[6], hdr[7] = sw, sh
    blob = bytes(blob)
    st = (C.c_uint64 * 5)()
    L.ndo_render(blob, 0, 0, sw, sh, cores, None, None, None, None, None, st)
    uniq = st[0] + st[1] + st[2]
    sec = float(np.mean(times))
    return {"seconds_per_step": sec, "rays_unique": int(uniq), "rays_ref": int(st[3]),
            "mrays_unique": uniq / sec / 1e6, "mrays_ref": st[3] / sec / 1e6,
            "frames_per_s_full": 1.0 / (sec * SAMPLE_DIV * SAMPLE_DIV),
            "cores": cores, "sample": f"same scene/frame at {sw}x{sh} (1/{SAMPLE_DIV*SAMPLE_DIV} of the pixels), "
                                      f"render_image only (kd build excluded), {cores} pthreads, mean of {steps}",
            "kd_build_seconds": R.kd_seconds}


def run_reference(args, rank, world):
    if rank != 0:
        return
    key, W, H, desc, *_ = WORKLOADS[args.workload]
    try:
        r = reference_sample(args.workload, args.steps, args.warmup)
    except Exception as e:  # the oracle always exists in a built tree; say why if not
        emit({"impl": "reference", "unavailable": str(e)[:200]})
        return
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": r["mrays_unique"], "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "note": "CPU: unmodified reference render_image on host cores"},
        "frames_per_s": r["frames_per_s_full"],
        "host_prepass": {"kd_tree_build_s": r["kd_build_seconds"],
                         "note": "the reference's serial kd builder (kd-tree.c:421), per frame, not in `value`"},
        "rays": {"unique_per_step": r["rays_unique"], "as_executed_by_reference_per_step": r["rays_ref"],
                 "mrays_as_executed": r["mrays_ref"]},
        "cpu_baseline": {"value": r["mrays_unique"], "unit": "Mrays/s", "cores": r["cores"],
                         "kind": "reference", "sample": r["sample"]},
        "e2e": {"value": r["mrays_unique"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------
# the stock command line with the library bound in (N = 1)
# --------------------------------------------------------------------------------------
def plugin_e2e(workload):
    """Wall time per frame of integration/ndt_b200_demo: the UNMODIFIED ndt main() (getopt, scene plugin, per
    frame scene_setup -> kd_tree_build -> camera_aim -> render_image -> image file) with kd_tree_build and
    render_image bound to libndt_b200, as ndt itself reports it."""
    key, W, H, desc, scene, dims, cfg, frame = WORKLOADS[workload]
    demo = os.path.join(ROOT, "integration", "ndt_b200_demo")
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(demo) or not os.path.exists(os.path.join(ref, "libndt_ref.so")):
        return {"unavailable": "integration/ndt_b200_demo not built (needs oracle/_ref)"}
    if workload == "config5_yaml":
        # the YAML file holds ONE document = one frame (scene.c:2067-2088); asked for frames 1..4 the reference's loader
        # never returns, and this measurement is a five-frame loop
        return {"unavailable": "one-document YAML scene: the five-frame loop of this measurement does not apply"}
    import tempfile
    base = [demo, "-d", str(dims), "-r", f"{W}x{H}", "-o", os.path.join(ref, "objects")]
    if workload == "config4_anim":
        frame = 0
    if scene:
        base += ["-s", os.path.join(ref, "scenes", scene + ".so")]
    if cfg:
        base += ["-u", os.path.join(ROOT, cfg) if not os.path.isabs(cfg) else cfg]
    out = {}
    import re
    try:
        with tempfile.TemporaryDirectory() as tmp:
            # five frames in ONE process; ndt prints the cumulative wall time of its frame loop after every frame
            # ("N frames took Xs", ndt.c:2018-2022, scene_setup included); the first frame also pays the CUDA
            # start-up and the graph build, so the per-frame time is taken over frames 2..5
            t0 = time.perf_counter()
            r = subprocess.run(base + ["-f", f"{frame}:{frame + 4}:300"], cwd=tmp, capture_output=True, text=True,
                               stdin=subprocess.DEVNULL, timeout=240, env=dict(os.environ, NDT_B200_TIMING="1", NDT_B200_KD_TIMING="1"))
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                return {"unavailable": (r.stdout[-300:] + r.stderr[-300:]).replace("\n", " ")}
            cum = [float(m.group(2)) for m in re.finditer(r"^\s*(\d+) frames? took ([0-9.]+)s", r.stdout, re.M)]
            if len(cum) < 5:
                return {"unavailable": "could not parse ndt's frame timings"}
            out["seconds_per_frame"] = (cum[4] - cum[0]) / 4.0
            out["frames_per_s"] = 4.0 / (cum[4] - cum[0]) if cum[4] > cum[0] else None
            out["first_frame_s"] = cum[0]
            out["process_wall_s_5_frames"] = wall
            stages = [ln.strip() for ln in r.stderr.splitlines() if ln.startswith("ndt_b200_")]
            out["stages_last_frame"] = stages[-2:]
    except Exception as e:
        return {"unavailable": str(e)[:200]}
    out["command"] = " ".join(os.path.relpath(x, ROOT) if x.startswith(ROOT) else x for x in base) + " -f a:b:300"
    out["path"] = ("unmodified ndt main(): scene_setup, ndt_b200_kd_tree_build (GPU), camera_aim, ndt_b200_render_image "
                   "(flatten with threaded bounding-sphere fits, upload, render, fp64 frame D2H), PPM file")
    return out


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def shared_host_frames(tag, shape, rank, world, dist):
    """One host buffer for the frames of a step, shared by the ranks of the box (POSIX shared memory) and
    page-locked in every rank, so that each GPU copies its frames into it over its own PCIe link."""
    import torch
    n = int(np.prod(shape))
    if world == 1:
        import ndt_b200
        return ndt_b200.pinned_empty(shape, np.uint8), "pinned (cudaMallocHost)"
    path = f"/dev/shm/ndtb200_{os.environ.get('MASTER_PORT', '0')}_{tag}"
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(n)
    dist.barrier()
    mm = np.memmap(path, dtype=np.uint8, mode="r+", shape=tuple(shape))
    kind = "POSIX shared memory, page-locked in every rank (cudaHostRegister)"
    try:
        err = torch.cuda.cudart().cudaHostRegister(mm.ctypes.data, n, 0)
        if int(err) != 0:
            kind = f"POSIX shared memory, pageable (cudaHostRegister -> {int(err)})"
    except Exception as e:          # noqa: BLE001
        kind = f"POSIX shared memory, pageable ({type(e).__name__})"
    dist.barrier()
    if rank == 0:
        os.unlink(path)             # the mappings keep it alive
    return mm, kind


def run_ours(args, rank, world, local_rank):
    import torch
    import ndt_b200
    from ndt_b200 import multi

    key, W, H, desc, *_ = WORKLOADS[args.workload]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # the frames go out over NVLink in the background of the render: two channels per peer carry the 8-33 MB of a
        # frame in a fraction of its render time and leave the SMs to the persistent trace kernels
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "2")
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    flats = load_flats(args.workload)
    n_distinct = len(flats)
    F = FRAMES_PER_GPU
    n_frames = F * world
    mine = list(range(rank * F, rank * F + F))                   # static first assignment: my own frames
    # IN_FLIGHT contexts per GPU: while one frame sits in the tail of a persistent k_trace launch, the other frame's
    # kernels fill the idle SMs (frames are independent, ndt.c:1771-1778 renders them on different MPI ranks)
    ctxs = [ndt_b200.Context(local_rank) for _ in range(IN_FLIGHT)]
    resident = {}                                                # ctx -> index of the flat scene it holds
    for c in ctxs:
        c.upload(flats[0]); resident[c] = 0
    ctx = ctxs[0]

    # rank 0 holds every frame of a step (two buffers: the receives of a step may start while the previous step's
    # frames are still being consumed); the others stage their own frames for the send
    nbuf = 2
    frames_dev = [torch.zeros((n_frames, H, W, 4), dtype=torch.uint8, device=dev) for _ in range(nbuf)] if rank == 0 else None
    stage = [torch.zeros((F, H, W, 4), dtype=torch.uint8, device=dev) for _ in range(nbuf)] if rank != 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    gathers = [multi.FrameGather(dist, rank, world, F) for _ in range(nbuf)] if world > 1 else None

    def barrier():
        for g in (gathers or []):
            g.end()                 # every frame of every step has arrived
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    totals = {"rays": 0, "dev_ms": 0.0, "launches": 0, "h2d": 0}
    lock = threading.Lock()

    def run_frames(work_one, after_start=None):
        """work_one(ctx, j) for my frames j = 0..F-1 on IN_FLIGHT host threads (ctypes releases the GIL inside the
        library); yields j as frames finish, in completion order.  after_start() runs in the calling thread once
        the workers are rendering."""
        todo = list(range(F))
        done = queue.Queue()

        def body(c):
            try:
                while True:
                    with lock:
                        if not todo:
                            break
                        j = todo.pop(0)
                    work_one(c, j)
                    done.put(j)
            except Exception as e:      # surface worker failures in the main thread
                done.put(e)
            done.put(None)
        ths = [threading.Thread(target=body, args=(c,)) for c in ctxs]
        for t in ths:
            t.start()
        if after_start is not None:
            after_start()
        live = len(ths)
        try:
            while live:
                x = done.get()
                if x is None:
                    live -= 1
                elif isinstance(x, Exception):
                    raise x
                else:
                    yield x
        finally:
            for t in ths:
                t.join()

    def ensure_scene(c, f, upload):
        k = f % n_distinct
        if upload or resident.get(c) != k:
            c.upload(flats[k])                                    # scene H2D from host memory
            resident[c] = k
            with lock:
                totals["h2d"] += len(flats[k])

    def device_step(i):
        b = i % nbuf
        flush.fill_(i & 0xFF)                                     # L2 flush between steps
        torch.cuda.current_stream().synchronize()
        g = gathers[b] if gathers else None
        if g is not None:
            g.end()                 # the transfers of the step that used this buffer last (two steps ago) are complete

        def post():                 # rank 0: the receives of this step, posted while its own frames already render
            if g is not None:
                g.begin(frames_dev[b] if rank == 0 else None)

        def one(c, j):
            f = mine[j]
            ensure_scene(c, f, False)
            dst = frames_dev[b][f] if rank == 0 else stage[b][j]
            c.launch_tile(0, 0, W, H, d_u8=dst.data_ptr())
            st = c.sync()
            with lock:
                totals["rays"] += st.rays_unique
                totals["dev_ms"] += st.device_ms
                totals["launches"] += st.launches
        for j in run_frames(one, post):
            if g is not None and rank != 0:
                g.send(stage[b][j], j)                            # NCCL, while the next frame renders

    def timed(fn, k, w):
        for i in range(w):
            fn(i)
        barrier()
        for kk in totals:
            totals[kk] = 0
        t0 = time.perf_counter()
        for i in range(k):
            fn(w + i)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt, float(totals["rays"]), totals["dev_ms"], float(totals["launches"]), float(totals["h2d"])],
                             dtype=torch.float64, device=dev)
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            totals["h2d_all"] = tsum[4].item()
            return tmax[0].item(), tsum[1].item(), tmax[2].item(), tsum[3].item()
        totals["h2d_all"] = float(totals["h2d"])
        return dt, float(totals["rays"]), totals["dev_ms"], float(totals["launches"])

    # ---- algorithmic flops of a frame (counting build, untimed) and FP64 peaks ------------
    flops_frame = 0
    peak_nf = peak_f = None
    scratch = frames_dev[0][0] if rank == 0 else stage[0][0]
    if rank == 0:
        ctx.set_options(ndt_b200.OPT_COUNT_FLOPS)
        ctx.launch_tile(0, 0, W, H, d_u8=scratch.data_ptr())
        flops_frame = ctx.sync().flops
        ctx.set_options(0)
        peak_nf = ctx.fp64_peak(False)
        peak_f = ctx.fp64_peak(True)
    # one frame alone on the GPU: CUDA-event time of its kernels on the launching stream (the roofline's clock)
    solo_ms = []
    for i in range(4):
        flush.fill_(i)
        torch.cuda.synchronize()
        ctx.launch_tile(0, 0, W, H, d_u8=scratch.data_ptr())
        st_solo = ctx.sync()
        if i:
            solo_ms.append(st_solo.device_ms)
    solo_ms = float(np.median(solo_ms))
    solo_launches = int(st_solo.launches)

    # ---- device-resident value ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dt, rays, dev_ms, launches = timed(device_step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    frames_total = n_frames * args.steps
    value = rays / dt / 1e6

    # ---- end to end: the C ABI with host buffers --------------------------------------------
    host_frames, host_kind = shared_host_frames("e2e", (n_frames, H, W, 4), rank, world, dist)
    host_views = [ndt_b200.Frame(W, H, ()) for _ in range(F)]
    for j, fr in enumerate(host_views):
        fr.rgba_u8 = host_frames[mine[j]]

    def e2e_step(i):
        first = set(ctxs)

        def one(c, j):
            f = mine[j]
            with lock:
                up = c in first
                first.discard(c)
            ensure_scene(c, f, up)                                # every context uploads its scene every step
            fr = host_views[j]
            c.render_tile(0, 0, W, H, out=fr)                     # C ABI, HOST buffer (shared by the ranks), D2H inside the call
            with lock:
                totals["rays"] += fr.stats.rays_unique
                totals["launches"] += fr.stats.launches
                totals["dev_ms"] += fr.stats.device_ms
        for _ in run_frames(one):
            pass
    edt, erays, _, _ = timed(e2e_step, args.steps, max(1, args.warmup // 2))
    e2e_value = erays / edt / 1e6
    h2d_total = int(totals["h2d_all"] / args.steps)              # counted from the blobs uploaded inside the timed steps
    d2h = n_frames * H * W * 4

    # ---- N GPUs == 1 GPU, byte for byte -------------------------------------------------------
    equal = None
    if rank == 0:
        check = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
        last_b = (args.warmup + args.steps - 1) % nbuf
        host_t = torch.from_numpy(np.asarray(host_frames))
        ok_dev = ok_host = True
        alone = {}
        for f in range(n_frames):
            k = f % n_distinct
            if k not in alone:
                ctx.upload(flats[k])
                ctx.launch_tile(0, 0, W, H, d_u8=check.data_ptr())
                ctx.sync()
                alone[k] = check.clone()
            ok_dev = ok_dev and bool(torch.equal(frames_dev[last_b][f], alone[k]))
            ok_host = ok_host and bool(torch.equal(host_t[f], alone[k].cpu()))
        equal = {"device_frames": ok_dev, "host_frames": ok_host, "frames_compared": n_frames,
                 "against": "rank 0's own single-GPU render of each distinct frame, torch.equal on the u8 RGBA"}

    if rank == 0:
        roofline = build_roofline(args.workload, solo_ms, flops_frame, peak_nf, peak_f, dt / args.steps)
        cpu = None
        plugin = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = reference_sample(args.workload, 1, 0)
                cpu = {"value": r["mrays_unique"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference",
                       "sample": r["sample"], "frames_per_s": r["frames_per_s_full"],
                       "mrays_as_executed_by_reference": r["mrays_ref"],
                       "kd_tree_build_s_per_frame": r["kd_build_seconds"]}
            except Exception as e:
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference",
                       "sample": "unavailable: " + str(e)[:160]}
        if world == 1 and not args.no_plugin_e2e:
            plugin = plugin_e2e(args.workload)
            if cpu and cpu.get("frames_per_s") and "seconds_per_frame" in plugin:
                ref_s = 1.0 / cpu["frames_per_s"] + (cpu.get("kd_tree_build_s_per_frame") or 0.0)
                plugin["reference_seconds_per_frame"] = ref_s
                plugin["reference_note"] = "reference render_image scaled from the 1/16 sample + its kd_tree_build, same cores"
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "frames_per_step": n_frames, "distinct_frames": n_distinct,
                       "parallelism": f"{F} frames per GPU and step, {IN_FLIGHT} in flight per GPU (one context + stream + host "
                                      f"thread each, rank-local queue); frames reach rank 0 over NCCL send/recv, each sent while "
                                      f"the next renders (value) / through one page-locked host buffer shared by the ranks (e2e)",
                       "l2": "256 MiB fill between steps (inside the bracket); the scene (3.4 MB for config 2) is meant to be "
                             "L2-resident within a step",
                       "rays": "unique trace_kd-equivalent queries (primary+bounce+shadow)"},
            "frames_per_s": frames_total / dt,
            "rays_per_frame": rays / frames_total,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h,
                    "frames_per_s": frames_total / edt, "host_buffer": host_kind,
                    "path": "ndt_b200_upload + ndt_b200_render_tile (HOST buffers) per frame; N>1: every rank into the one "
                            "host buffer of the box, rank 0 holds all frames"},
            "gpu_launches": int(launches),
            "launches_per_frame": solo_launches,
            "launches_note": "kernels; the host enqueues k_begin + ONE CUDA graph per frame (device-side generation loop)",
            "multi_gpu_equal": equal,
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if plugin is not None:
            line["e2e_plugin"] = plugin
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def build_roofline(workload, solo_ms, flops_frame, peak_nf, peak_f, step_s):
    """FP64 pipe.  Executed instruction counts come from the committed ncu pass of the same build
    (tools/ncu_fp64_ops.py -> profiles/r02_fp64_ops_<workload>.json); times are measured live."""
    peak = peak_nf / 1e3 if peak_nf else None          # T thread-instructions per second = non-fused TFLOP/s
    prof = None
    p = os.path.join(ROOT, "profiles", f"r02_fp64_ops_{'config4' if workload == 'config4_anim' else workload}.json")
    if os.path.exists(p):
        with open(p) as f:
            prof = json.load(f)
    out = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak,
           "peak_source": "measured live by ndt_b200_fp64_peak: DMUL+DADD chains on all SMs, no FMA contraction (MEASURED_PEAKS.json "
                          "has no FP64 entry); one FP64 thread instruction per lane slot is the rate parity allows",
           "peak_dfma_tflops": peak_f / 1e3 if peak_f else None,
           "frame_ms_alone": solo_ms}
    if prof and solo_ms > 0 and peak:
        ops = prof["fp64_thread_inst_dadd_dmul_dfma"]
        kt = prof["kernels"]
        tr = [v for k, v in kt.items() if "k_trace" in k or "k_pre<" in k]       # the query: k_pre (before the walk) + k_trace (the walk)
        tr_ops = sum(v["dadd"] + v["dmul"] + v["dfma"] for v in tr)
        tr_share = sum(v["ms_under_ncu"] for v in tr) / prof["frame_kernel_ms_under_ncu"]
        achieved = ops / (solo_ms * 1e-3) / 1e12
        out.update({
            "achieved": achieved, "frac": achieved / peak,
            "executed_fp64_thread_inst_per_frame": ops,
            "executed_source": os.path.relpath(p, ROOT) + ": smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum over "
                               "the kernels of one frame (ncu, this build); DFMA (division / sqrt / libm sequences) counted as ONE",
            "kernel": "the nearest-hit and shadow queries: k_pre<NP,0|1> (infinite objects + root box, every ray) + k_trace<NP,0|1> (the walk, "
                      "the rays that enter the root box)",
            "kernel_share_of_frame": tr_share,
            "kernel_achieved": tr_ops / (solo_ms * 1e-3 * tr_share) / 1e12 if tr_share > 0 else None,
            "kernel_frac": tr_ops / (solo_ms * 1e-3 * tr_share) / 1e12 / peak if tr_share > 0 else None,
            "kernel_fp64_pipe_active_pct_ncu": [v["fp64_pipe_active_pct"] for v in tr],
            "traffic": prof.get("dram_bytes_k_trace"),
            "traffic_unit": "bytes, dram__bytes_read.sum + dram__bytes_write.sum of the k_pre + k_trace launches of one frame (same ncu pass); "
                            "the scene is L2-resident, the traffic is ray queues and hit records",
        })
    else:
        out.update({"achieved": None, "frac": None, "executed_source": "profiles/r02_fp64_ops_<workload>.json missing"})
    if flops_frame and solo_ms > 0 and peak:
        alg = flops_frame / (solo_ms * 1e-3) / 1e12
        out.update({"algorithmic_flops_per_frame": flops_frame, "achieved_algorithmic": alg, "frac_algorithmic": alg / peak,
                    "algorithmic_note": "SURVEY 8(d): flops of the REFERENCE algorithm per unique ray (counting build); the kernels "
                                        "skip ~80 % of them with exact culls, so this is a throughput figure, not pipe utilisation",
                    "achieved_algorithmic_over_step": flops_frame * FRAMES_PER_GPU / step_s / 1e12 if step_s > 0 else None})
    return out


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints
    (NCCL's version banner, the reference's printf chatter) was sent to stderr."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # fd 1 -> stderr for C libraries and print() alike
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-plugin-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
