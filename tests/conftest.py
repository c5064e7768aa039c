import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(args, target):
    if not os.path.exists(target):
        subprocess.run(["make"] + args, cwd=ROOT, check=True, capture_output=True)
    return target


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_lib():
    """oracle/libndt_oracle.so: the plain-C restatement (CPU checker)."""
    p = _make(["-C", "oracle", "oracle"], os.path.join(ROOT, "oracle", "libndt_oracle.so"))
    L = C.CDLL(p)
    L.ndo_render.argtypes = [C.c_char_p] + [C.c_int] * 5 + [C.c_void_p] * 6
    L.ndo_trace.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                            C.POINTER(C.c_int)]
    return L


@pytest.fixture(scope="session")
def emu_lib():
    """tests/emu/libndt_emu.so: the device core compiled for the CPU."""
    p = os.path.join(HERE, "emu", "libndt_emu.so")
    if not os.path.exists(p):
        subprocess.run(["g++", "-m64", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off",
                        "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "ndt_b200", "csrc"),
                        "-shared", "-o", p, os.path.join(HERE, "emu", "emu_driver.cpp"), "-lm"], check=True)
    L = C.CDLL(p)
    L.emu_render.argtypes = [C.c_char_p] + [C.c_int] * 4 + [C.c_void_p] * 6
    return L


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref).  Present in the build container
    and on the GPU box (the built .so files travel); skip when absent."""
    from oracle import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return refharness.RefHarness()


class Buffers:
    def __init__(self, w, h):
        self.f64 = np.zeros((h, w, 4), np.float64)
        self.u8 = np.zeros((h, w, 4), np.uint8)
        self.hit = np.zeros((h, w), np.uint8)
        self.id = np.zeros((h, w), np.int32)
        self.depth = np.zeros((h, w), np.float64)

    def ptrs(self):
        return [a.ctypes.data for a in (self.f64, self.u8, self.hit, self.id, self.depth)]


def oracle_render(L, flat, x0=0, y0=0, tw=None, th=None, threads=0):
    tw = tw or flat.header.width
    th = th or flat.header.height
    b = Buffers(tw, th)
    st = (C.c_uint64 * 5)()
    rc = L.ndo_render(flat.blob, x0, y0, tw, th, threads or os.cpu_count(), *b.ptrs(), st)
    assert rc == 0
    b.stats = dict(zip(("rays_primary", "rays_bounce", "rays_shadow", "rays_ref", "samples"), list(st)))
    return b


def emu_render(L, flat, x0=0, y0=0, tw=None, th=None):
    tw = tw or flat.header.width
    th = th or flat.header.height
    b = Buffers(tw, th)
    st = (C.c_uint64 * 8)()
    rc = L.emu_render(flat.blob, x0, y0, tw, th, *b.ptrs(), st)
    assert rc == 0
    b.stats = dict(zip(("rays_primary", "rays_bounce", "rays_shadow", "rays_ref", "samples", "flops",
                        "generations", "overflow"), list(st)))
    return b


def load_flat(key):
    import ndt_b200
    return ndt_b200.FlatScene.load(os.path.join(GOLDEN, key + ".ndsf.gz"))


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))
