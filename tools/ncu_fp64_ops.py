"""ncu_fp64_ops.py <ncu --csv log> [skip_frames] -- per-kernel executed FP64 work from an
`ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum,...`
launch list of tools/perf_frame.py <workload> 1 (two frames: the first is warm-up).  Prints, per kernel
name, launches, time, executed FP64 thread instructions (DADD + DMUL + DFMA) of the LAST frame and writes the
totals as JSON (second argument) for bench.py's roofline.executed."""
import csv, json, sys, re, collections

def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    return rows

def main():
    path = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    rows = load(path)
    per = collections.OrderedDict()       # launch id -> dict
    for r in rows:
        k = int(r["ID"])
        d = per.setdefault(k, {"name": r["Kernel Name"]})
        v = r["Metric Value"].replace(",", "")
        try:
            d[r["Metric Name"]] = float(v)
        except ValueError:
            pass
        d["unit:" + r["Metric Name"]] = r["Metric Unit"]
    launches = list(per.values())
    # the last frame = everything after the last k_finish but one
    fin = [i for i, d in enumerate(launches) if d["name"].startswith("k_finish")]
    first = fin[-2] + 1 if len(fin) >= 2 else 0
    frame = launches[first:fin[-1] + 1] if fin else launches
    agg = collections.OrderedDict()
    for d in frame:
        nm = re.sub(r"\(.*", "", d["name"])
        a = agg.setdefault(nm, collections.Counter())
        a["launches"] += 1
        t = d.get("gpu__time_duration.sum", 0.0)
        u = d.get("unit:gpu__time_duration.sum", "ns")
        a["ms"] += t * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(u, 1e-6)
        for m, key in (("dadd", "dadd"), ("dmul", "dmul"), ("dfma", "dfma")):
            a[key] += d.get(f"smsp__sass_thread_inst_executed_op_{m}_pred_on.sum", 0.0)
        a["fp64_thread_inst"] += d.get("smsp__thread_inst_executed_pipe_fp64_pred_on.sum", 0.0)
        a["fp64_warp_inst"] += d.get("smsp__inst_executed_pipe_fp64.sum", 0.0)
        a["warp_inst"] += d.get("smsp__inst_executed.sum", 0.0)
        a["pipe_pct_x_ms"] += d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * a["ms"] * 0
        a["pipe_pct_wsum"] += d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
        a["t_raw"] += t
        a["regs"] = max(a["regs"], d.get("launch__registers_per_thread", 0))
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            u2 = d.get("unit:" + m, "byte")
            a["dram"] += d.get(m, 0.0) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u2, 1.0)
    tot_ms = sum(a["ms"] for a in agg.values())
    print(f"{'kernel':28s} {'n':>4s} {'ms':>8s} {'share':>6s} {'DADD+DMUL+DFMA (G thread-inst)':>32s} {'Ginst/s':>9s} {'pipe%':>6s} {'regs':>5s}")
    res = {"frame_kernel_ms_under_ncu": tot_ms, "kernels": {}}
    for nm, a in agg.items():
        ops = a["dadd"] + a["dmul"] + a["dfma"]
        pipe = a["pipe_pct_wsum"] / a["t_raw"] if a["t_raw"] else 0.0
        print(f"{nm:28s} {a['launches']:4d} {a['ms']:8.3f} {a['ms']/tot_ms*100:5.1f}% {ops/1e9:32.3f} {ops/1e9/(a['ms']*1e-3) if a['ms'] else 0:9.1f} {pipe:6.1f} {int(a['regs']):5d}")
        res["kernels"][nm] = {"launches": a["launches"], "ms_under_ncu": a["ms"], "dadd": a["dadd"], "dmul": a["dmul"],
                              "dfma": a["dfma"], "fp64_thread_inst": a["fp64_thread_inst"], "fp64_warp_inst": a["fp64_warp_inst"],
                              "warp_inst": a["warp_inst"], "fp64_pipe_active_pct": pipe, "registers": int(a["regs"]),
                              "dram_bytes": a["dram"]}
    allops = sum(k["dadd"] + k["dmul"] + k["dfma"] for k in res["kernels"].values())
    res["fp64_thread_inst_dadd_dmul_dfma"] = allops
    res["dram_bytes_k_trace"] = sum(k["dram_bytes"] for n_, k in res["kernels"].items() if "k_trace" in n_ or "k_pre<" in n_) or None
    res["source"] = "ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum,... " \
                    "--clock-control none on `NDT_B200_NO_GRAPH=1 python tools/perf_frame.py <workload> 1` (second frame); " + path
    print(f"{'frame':28s} {'':4s} {tot_ms:8.3f} {'':6s} {allops/1e9:32.3f}")
    if out:
        json.dump(res, open(out, "w"), indent=1)

if __name__ == "__main__":
    main()
