"""The binding of INTEGRATION.md as real code: integration/ndt_b200_demo is the UNMODIFIED reference
(oracle/_ref/libndt_ref.so: getopt, scene setup, kd build, camera aim, frame loop) with render_image
pre-empted by integration/render_image_b200.c, which forwards to ndt_b200_render_image[_aa].  The
stock command line must produce the frame the oracle predicts."""
import glob
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_flat, oracle_render

pytestmark = pytest.mark.gpu
DEMO = os.path.join(ROOT, "integration", "ndt_b200_demo")


def read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = map(int, f.readline().split())
        assert f.readline().strip() == b"255"
        return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)


def run_demo(tmp_path, *args):
    if not os.path.exists(DEMO) or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libndt_ref.so")):
        pytest.skip("integration/ndt_b200_demo or oracle/_ref not built (needs /root/reference at build time)")
    r = subprocess.run([DEMO, *args, "-o", os.path.join(ROOT, "oracle", "_ref", "objects")], cwd=tmp_path,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = glob.glob(os.path.join(tmp_path, "images", "**", "*.ppm"), recursive=True)
    assert len(out) == 1, out
    return read_ppm(out[0])


def test_stock_command_line_renders_on_the_gpu(tmp_path, oracle_lib):
    got = run_demo(tmp_path, "-d", "4", "-f", "0", "-r", "160x90")
    want = oracle_render(oracle_lib, load_flat("config1_default4d")).u8[..., :3]
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert float((d <= 1).mean()) >= 0.999


def test_stock_command_line_with_recursive_aa(tmp_path, oracle_lib):
    import ctypes as C
    got = run_demo(tmp_path, "-d", "4", "-f", "0", "-r", "96x54", "-a", "20,4")
    flat = load_flat("aa_default4d")
    oracle_lib.ndo_render_aa.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    want = np.zeros((54, 96, 4), np.uint8)
    assert oracle_lib.ndo_render_aa(flat.blob, os.cpu_count(), 20, 4, want.ctypes.data, None, None) == 0
    d = np.abs(got.astype(np.int16) - want[..., :3].astype(np.int16))
    assert float((d <= 1).mean()) >= 0.999


def test_stock_command_line_loads_a_yaml_scene(tmp_path, oracle_lib):
    """`ndt -s scenes/yaml.so -u file.yaml -d 10` (README.md:423-430 of the reference; BASELINE config 5):
    scenes/yaml.c + scene_read_yaml run unmodified over yaml_lite, the frame comes from the GPU."""
    got = run_demo(tmp_path, "-s", os.path.join(ROOT, "oracle", "_ref", "scenes", "yaml.so"),
                   "-u", os.path.join(ROOT, "tests", "scenes", "config5_mixed10d.yaml"),
                   "-d", "10", "-f", "0", "-r", "96x54")
    want = oracle_render(oracle_lib, load_flat("config5_yaml10d")).u8[..., :3]
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert float((d <= 1).mean()) >= 0.999
