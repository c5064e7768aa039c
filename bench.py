#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric: Mrays/s = unique nearest-hit queries (trace_kd calls, object.c:683:
primary + reflection/refraction + shadow) per second; frames/s is reported next
to it.  Workload: BASELINE config 2 -- scenes/hypercube.c at 8 dimensions,
1920x1080, reflections on, frame 0 -- which is the configuration the metric is
quoted on and fits one GPU.  The scene enters as the flat blob the struct-ABI
adapter produced from the reference's own scene + kd-tree (tests/golden/, made
by tests/golden/make_golden.py), so nothing under oracle/ runs on our arm.

A STEP renders FRAMES_PER_GPU frames per GPU, each as TILES_PER_FRAME row bands
(work items).  With N>1 the items of a step are pulled from one shared counter
(dynamic tile queue) and finished tiles are gathered to rank 0 with NCCL
send/recv; per-GPU work is fixed as N grows ("weak").

  value  device-resident: scene already in HBM, outputs stay in HBM (rank 0's).
  e2e    N=1: ndt_b200_upload + ndt_b200_render_tile with HOST buffers (scene
         H2D and image D2H inside the timed region).  N>1: upload + device
         render + NCCL gather + rank 0's D2H of every frame.
  roofline  FP64 pipe: algorithmic flops of a step (counting build, untimed)
         / CUDA-event kernel time, against the NON-FUSED DMUL+DADD peak measured
         live (parity forbids FMA contraction); the DFMA peak is quoted too.
  cpu_baseline / --impl reference: the UNMODIFIED reference's render_image
         (oracle/_ref) on all host cores, on a bounded sample of the same
         workload (same scene and frame at 1/16 of the pixels).
"""
import argparse
import ctypes as C
import json
import os
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
    "config5": ("config5_mixed10d", 1920, 1080,
                "BASELINE config 5 (C twin): mixed10d -d 10, 1920x1080, frame 0", "mixed10d", 10, None, 0),
    "config5_yaml": ("config5_yaml10d", 1920, 1080,
                     "BASELINE config 5: scenes/yaml.c -u tests/scenes/config5_mixed10d.yaml -d 10, 1920x1080 "
                     "(hplane, hcylinder, hdisk, hfacet; 6 lights, shadow rays), loaded by the reference's "
                     "scene_read_yaml over yaml_lite", "yaml", 10, "tests/scenes/config5_mixed10d.yaml", 0),
}
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (both k_trace launches of generation 0), from the
# ncu --set full capture of the same frame (profiles/r01_ncu_k_trace_bundle_*): 1.1 + 68.7 MB and 238.8 + 86.0 MB
NCU_TRAFFIC_GB = {"config2": 0.3945}
FRAMES_PER_GPU = 4
TILES_PER_FRAME = 1
IN_FLIGHT = int(os.environ.get("NDT_IN_FLIGHT", "2"))           # frames in flight per GPU: one ndt_b200 context (own CUDA stream) and one host thread each
SAMPLE_DIV = 4          # reference sample: width/4 x height/4 = 1/16 of the pixels


def load_flat(key, w, h):
    import ndt_b200
    f = ndt_b200.FlatScene.load(os.path.join(ROOT, "tests", "golden", key + ".ndsf.gz"))
    return f.retarget(w, h)


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
    """Times render_image (ndt.c:900) of oracle/_ref on the bounded sample; returns a dict."""
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
    import ndt_b200
    L = C.CDLL(os.path.join(ROOT, "oracle", "libndt_oracle.so"))
    L.ndo_render.argtypes = [C.c_char_p] + [C.c_int] * 5 + [C.c_void_p] * 6
    flat = load_flat(key, sw, sh)
    st = (C.c_uint64 * 5)()
    L.ndo_render(flat.blob, 0, 0, sw, sh, cores, None, None, None, None, None, st)
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
        "rays": {"unique_per_step": r["rays_unique"], "as_executed_by_reference_per_step": r["rays_ref"],
                 "mrays_as_executed": r["mrays_ref"]},
        "cpu_baseline": {"value": r["mrays_unique"], "unit": "Mrays/s", "cores": r["cores"],
                         "kind": "reference", "sample": r["sample"]},
        "e2e": {"value": r["mrays_unique"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import ndt_b200
    from ndt_b200 import multi

    key, W, H, desc, *_ = WORKLOADS[args.workload]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    flat = load_flat(key, W, H)
    # IN_FLIGHT contexts per GPU: while one frame sits in the tail of a persistent k_trace launch or in the
    # host's per-generation hand-off, the other frame's kernels fill the idle SMs (frames are independent,
    # ndt.c:1771-1778 renders them on different MPI ranks)
    ctxs = [ndt_b200.Context(local_rank) for _ in range(IN_FLIGHT)]
    for c in ctxs:
        c.upload(flat)
    ctx = ctxs[0]

    n_frames = FRAMES_PER_GPU * world
    band = (H + TILES_PER_FRAME - 1) // TILES_PER_FRAME
    items = [(f, t * band, min(band, H - t * band)) for f in range(n_frames) for t in range(TILES_PER_FRAME)]
    queue = multi.TileQueue(dist, rank, world)
    # rank 0 holds every frame of the step; other ranks a staging area for their own tiles
    frames_dev = torch.zeros((n_frames, H, W, 4), dtype=torch.uint8, device=dev) if rank == 0 else None
    stage = torch.zeros((len(items), band, W, 4), dtype=torch.uint8, device=dev) if world > 1 and rank != 0 else None
    frames_host = torch.zeros((n_frames, H, W, 4), dtype=torch.uint8).pin_memory() if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    totals = {"rays": 0, "dev_ms": 0.0, "launches": 0}

    serial = [0]

    lock = threading.Lock()

    def run_workers(work):
        """work(ctx) on IN_FLIGHT host threads (ctypes releases the GIL inside the library)."""
        errs = []

        def body(c):
            try:
                work(c)
            except Exception as e:      # surface worker failures in the main thread
                errs.append(e)
        ths = [threading.Thread(target=body, args=(c,)) for c in ctxs]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]

    def device_step(_unused, upload, frames=None):
        frames = frames_dev if frames is None else frames
        step_id = serial[0]                               # queue keys must never repeat within a job
        serial[0] += 1
        if upload:
            for c in ctxs:
                c.upload(flat)
        flush.fill_(step_id & 0xFF)                       # L2 flush between steps
        torch.cuda.current_stream().synchronize()         # not the whole device: a read-back of the previous step may still run
        mine = []
        pull = queue.pull(step_id, len(items))

        def work(c):
            while True:
                with lock:
                    idx = next(pull, None)
                    if idx is None:
                        return
                    k = len(mine)
                    mine.append(idx)
                f, y0, th = items[idx]
                dst = frames[f, y0:y0 + th] if rank == 0 else stage[k, :th]
                c.launch_tile(0, y0, W, th, d_u8=dst.data_ptr())
                st = c.sync()
                with lock:
                    totals["rays"] += st.rays_unique
                    totals["dev_ms"] += st.device_ms
                    totals["launches"] += st.launches
        run_workers(work)
        if world > 1:
            multi.gather_tiles(dist, rank, world, items, mine, stage, frames, band)
        return mine

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
            t = torch.tensor([dt, float(totals["rays"]), totals["dev_ms"], float(totals["launches"])],
                             dtype=torch.float64, device=dev)
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return tmax[0].item(), tsum[1].item(), tmax[2].item(), tsum[3].item()
        return dt, float(totals["rays"]), totals["dev_ms"], float(totals["launches"])

    # ---- algorithmic flops of one step (counting build, untimed) and FP64 peaks ------------
    flops_frame = 0
    peak_nf = peak_f = None
    if rank == 0:
        ctx.set_options(ndt_b200.OPT_COUNT_FLOPS)
        for t in range(TILES_PER_FRAME):
            y0 = t * band
            th = min(band, H - y0)
            ctx.launch_tile(0, y0, W, th, d_u8=frames_dev[0, y0:y0 + th].data_ptr())
            flops_frame += ctx.sync().flops
        ctx.set_options(0)
        peak_nf = ctx.fp64_peak(False)
        peak_f = ctx.fp64_peak(True)
    # one frame alone on the GPU: CUDA-event time of its kernels on the launching stream (roofline numerator's clock)
    solo_ms = []
    for i in range(4):
        flush.fill_(i)
        torch.cuda.synchronize()
        ctx.launch_tile(0, 0, W, H, d_u8=(frames_dev[0] if rank == 0 else stage[0]).data_ptr())
        st_solo = ctx.sync()
        if i:
            solo_ms.append(st_solo.device_ms)
    solo_ms = float(np.median(solo_ms))

    # ---- device-resident value ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dt, rays, dev_ms, launches = timed(lambda i: device_step(i, False), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    frames_total = n_frames * args.steps
    value = rays / dt / 1e6

    # ---- end to end -------------------------------------------------------------------------
    h2d = len(flat) * IN_FLIGHT
    if world == 1:
        # one page-locked host frame buffer per frame of the step (ndt_b200_host_alloc)
        hosts = {(f, y0): ndt_b200.Frame(W, th, ("u8",), pinned=True) for f, y0, th in items}

        def e2e_step(i):
            todo = list(items)

            def work(c):
                c.upload(flat)                              # scene H2D from host memory, once per context
                while True:
                    with lock:
                        if not todo:
                            return
                        f, y0, th = todo.pop(0)
                    fr = hosts[(f, y0)]
                    c.render_tile(0, y0, W, th, out=fr)     # C ABI, HOST buffers, D2H inside the call
                    with lock:
                        totals["rays"] += fr.stats.rays_unique
                        totals["launches"] += fr.stats.launches
                        totals["dev_ms"] += fr.stats.device_ms
            run_workers(work)
        d2h = n_frames * H * W * 4
    else:
        # rank 0 reads every step's frames back to pinned host memory.  The read of step i runs on a copy stream
        # while step i+1 renders into the other of two frame buffers (at N = 8 the 265 MB per step over rank 0's
        # one PCIe link would otherwise serialise behind the render); every copy completes inside the timed region
        # (timed() ends with barrier + device synchronize).
        dev_bufs = [frames_dev, torch.zeros_like(frames_dev)] if rank == 0 else [None, None]
        host_bufs = [frames_host, torch.zeros_like(frames_host).pin_memory()] if rank == 0 else [None, None]
        copy_stream = torch.cuda.Stream(device=dev) if rank == 0 else None
        copy_done = [None, None]

        def e2e_step(i):
            b = i & 1
            if rank == 0 and copy_done[b] is not None:
                copy_done[b].synchronize()                  # the read-back of two steps ago still owns this buffer
            device_step(i, True, dev_bufs[b])
            if rank == 0:
                ready = torch.cuda.Event()
                ready.record()                              # after the gather on the current stream
                copy_stream.wait_event(ready)
                with torch.cuda.stream(copy_stream):
                    host_bufs[b].copy_(dev_bufs[b], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy_stream)
                copy_done[b] = done
        d2h = n_frames * H * W * 4
    edt, erays, _, _ = timed(e2e_step, args.steps, max(1, args.warmup // 2))
    e2e_value = erays / edt / 1e6

    if rank == 0:
        step_s = dt / args.steps                          # wall time of one step (IN_FLIGHT frames overlap)
        achieved = flops_frame / (solo_ms * 1e-3) / 1e12 if solo_ms > 0 else 0.0
        achieved_step = flops_frame * FRAMES_PER_GPU / step_s / 1e12 if step_s > 0 else 0.0
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": peak_nf / 1e3, "unit": "TFLOP/s",
            "frac": achieved / (peak_nf / 1e3) if peak_nf else None,
            "traffic": NCU_TRAFFIC_GB.get(args.workload),
            "traffic_unit": "GB of DRAM reads+writes of the two k_trace launches of generation 0 (ncu --set full, "
                            "profiles/r01_ncu_k_trace_bundle_details.txt): the scene is L2-resident, the traffic is "
                            "ray queues and hit records; against ~25 GFLOP of algorithmic work in the same launches",
            "kernel": "k_trace<NP,0> + k_trace<NP,1> (nearest-hit and shadow queries; 52 % of a frame's kernel time on "
                      "config 2, k_shade 38 %, k_finish 10 %: launch list in profiles/) timed with CUDA events as ONE "
                      "whole frame alone on the GPU, k_shade/k_resolve/k_finish included in the denominator",
            "frame_ms_alone": solo_ms,
            "algorithmic_flops_per_frame": flops_frame,
            "peak_source": "measured live by ndt_b200_fp64_peak: non-fused DMUL+DADD chains on all SMs "
                           "(MEASURED_PEAKS.json has no FP64 entry); parity forbids FMA contraction",
            "peak_dfma_tflops": peak_f / 1e3, "frac_of_dfma_peak": achieved / (peak_f / 1e3) if peak_f else None,
            "achieved_over_step": achieved_step,
            "frac_over_step": achieved_step / (peak_nf / 1e3) if peak_nf else None,
            "note_over_step": f"{IN_FLIGHT} frames in flight per GPU: algorithmic flops of a step / its wall time",
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = reference_sample(args.workload, 1, 0)
                cpu = {"value": r["mrays_unique"], "unit": "Mrays/s", "cores": r["cores"], "kind": "reference",
                       "sample": r["sample"], "frames_per_s": r["frames_per_s_full"],
                       "mrays_as_executed_by_reference": r["mrays_ref"]}
            except Exception as e:
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference",
                       "sample": "unavailable: " + str(e)[:160]}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "frames_per_step": n_frames, "tiles_per_frame": TILES_PER_FRAME,
                       "parallelism": f"dynamic tile queue over {world} GPU(s), {IN_FLIGHT} frames in flight per GPU "
                                      f"(one context + stream each), NCCL gather to rank 0",
                       "l2": "256 MiB fill between steps (inside the bracket); the 3.4 MB scene is meant to be "
                             "L2-resident within a step",
                       "rays": "unique trace_kd-equivalent queries (primary+bounce+shadow)"},
            "frames_per_s": frames_total / dt,
            "rays_per_frame": rays / frames_total,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "frames_per_s": frames_total / edt,
                    "path": ("ndt_b200_upload + ndt_b200_render_tile (host buffers)" if world == 1 else
                             "ndt_b200_upload + launch_tile + NCCL gather to rank 0 + read-back to rank 0's pinned host "
                             "memory on a copy stream, double buffered")},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
