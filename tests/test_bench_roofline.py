"""The measurement side on the CPU: what bench.py's roofline block is computed from can be recomputed from the
committed ncu launch lists, and the bench line keeps the driver's contract."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

WORKLOADS = ["config1", "config2", "config4", "config5_yaml"]


@pytest.mark.parametrize("wl", WORKLOADS)
def test_executed_fp64_counts_follow_from_the_committed_launch_list(wl, tmp_path):
    """profiles/r02_fp64_ops_<workload>.json is tools/ncu_fp64_ops.py applied to profiles/r02_launches_fp64_<workload>.csv"""
    csv = os.path.join(ROOT, "profiles", f"r02_launches_fp64_{wl}.csv")
    out = tmp_path / "ops.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_fp64_ops.py"), csv, str(out)], check=True, capture_output=True)
    got = json.load(open(out))
    want = json.load(open(os.path.join(ROOT, "profiles", f"r02_fp64_ops_{wl}.json")))
    assert got["fp64_thread_inst_dadd_dmul_dfma"] == want["fp64_thread_inst_dadd_dmul_dfma"] > 1e9
    assert got["kernels"].keys() == want["kernels"].keys()
    # the query (k_pre + k_trace) is there in both modes and is the largest group of a frame
    names = list(got["kernels"])
    for k in ("k_pre<", "k_trace<"):
        assert sum(k in n for n in names) == 2, names
    q = sum(v["ms_under_ncu"] for n, v in got["kernels"].items() if "k_trace" in n or "k_pre<" in n)
    assert 0.35 < q / got["frame_kernel_ms_under_ncu"] < 0.7


def test_roofline_fraction_is_executed_instructions_over_time_over_peak():
    prof = json.load(open(os.path.join(ROOT, "profiles", "r02_fp64_ops_config2.json")))
    ops = prof["fp64_thread_inst_dadd_dmul_dfma"]
    r = bench.build_roofline("config2", solo_ms=2.5, flops_frame=30960869430, peak_nf=18200.0, peak_f=35000.0, step_s=0.009)
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s"
    assert r["achieved"] == pytest.approx(ops / 2.5e-3 / 1e12)
    assert r["frac"] == pytest.approx(ops / 2.5e-3 / 1e12 / 18.2)
    assert 0.1 < r["frac"] < 0.25 and r["frac_algorithmic"] > r["frac"]          # the culls skip most of the reference's flops
    assert r["traffic"] == prof["dram_bytes_k_trace"] > 0
    assert 0.3 < r["kernel_share_of_frame"] < 0.7 and r["kernel_frac"] > 0


def test_committed_bench_lines_keep_the_contract():
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}
    for name, n in (("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n8.json", 8)):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert need <= d.keys(), need - d.keys()
        assert d["n_gpus"] == n and d["unit"] == "Mrays/s" and d["dtype"] == "f64" and d["scaling"] == "weak"
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        assert d["multi_gpu_equal"]["device_frames"] and d["multi_gpu_equal"]["host_frames"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert "workload" in d["config"] and "model" not in d["config"]
        if n == 1:
            assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
            # recomputable from profiles/: executed instructions / frame time alone / measured peak
            r = d["roofline"]
            assert r["frac"] == pytest.approx(r["executed_fp64_thread_inst_per_frame"] / (r["frame_ms_alone"] * 1e-3) / 1e12 / r["peak"], rel=1e-6)


def test_the_one_document_yaml_workload_skips_the_five_frame_plugin_loop():
    """the reference's YAML loader never returns for a frame the file has no document for: bench.py must not ask"""
    r = bench.plugin_e2e("config5_yaml")
    assert "unavailable" in r
