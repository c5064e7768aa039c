"""The per-ray DEVICE code (ndt_b200/csrc/core.cuh + wave.cuh), compiled for
the CPU by tests/emu/emu_driver.cpp and driven through the same generation
loop / record fold as the CUDA kernels, must agree with the oracle bit for
bit (same libm here, so colour too).  This is the CPU-tier guard for the
wavefront formulation: queue order, child slots, resolve order, sample replay."""
import pytest

from conftest import bits_equal, emu_render, load_flat, oracle_render
from scenes import CASES


@pytest.mark.parametrize("case", CASES, ids=[c.key for c in CASES])
def test_device_core_on_cpu_equals_oracle(case, oracle_lib, emu_lib):
    flat = load_flat(case.key)
    a = oracle_render(oracle_lib, flat)
    b = emu_render(emu_lib, flat)
    assert bits_equal(a.hit, b.hit)
    assert bits_equal(a.id, b.id)
    assert bits_equal(a.depth, b.depth)
    assert bits_equal(a.f64, b.f64)
    assert bits_equal(a.u8, b.u8)
    for k in ("rays_primary", "rays_bounce", "rays_shadow", "rays_ref", "samples"):
        assert a.stats[k] == b.stats[k], k
    assert b.stats["overflow"] == 0 and b.stats["flops"] > 0


def test_device_core_partial_tile(oracle_lib, emu_lib):
    flat = load_flat("default5d_odd")
    a = oracle_render(oracle_lib, flat, x0=5, y0=3, tw=19, th=11)
    b = emu_render(emu_lib, flat, x0=5, y0=3, tw=19, th=11)
    assert bits_equal(a.f64, b.f64) and bits_equal(a.id, b.id)
