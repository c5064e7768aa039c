"""Known-answer vectors for trace_kd (object.c:683): ~190 rays per scene answered by
the UNMODIFIED reference (tests/golden/<key>.kat.json, made by make_golden.py):
rays towards bounding spheres, bounce rays leaving hit points, axis-aligned
directions with exact zeros, rays straight back, shadow-style rays with
dist_limit > 0 and == 0.  Between them the scenes cover every primitive type
(sphere, hplane, hdisk, orthotope, hcube, facet, hfacet, cylinder, hcylinder),
finite and infinite variants, and dimensions 3..12.

CPU tier: the oracle reproduces them bit for bit.  GPU tier: so does
ndt_b200_trace_rays (hit point and normal use only + - * / sqrt)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_flat
from scenes import BASE_CASES as CASES


def load_kats(key):
    with open(os.path.join(GOLDEN, key + ".kat.json")) as f:
        k = json.load(f)
    fh = float.fromhex
    o = np.array([[fh(x) for x in r["o"]] for r in k])
    v = np.array([[fh(x) for x in r["v"]] for r in k])
    lim = np.array([fh(r["limit"]) for r in k])
    found = np.array([r["found"] for r in k], np.int32)
    oid = np.array([r["id"] for r in k], np.int32)
    hit = np.array([[fh(x) for x in r["hit"]] for r in k])
    nrm = np.array([[fh(x) for x in r["normal"]] for r in k])
    return o, v, lim, found, oid, hit, nrm


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.mark.parametrize("case", CASES, ids=[c.key for c in CASES])
def test_oracle_answers_the_reference_kats(case, oracle_lib):
    flat = load_flat(case.key)
    o, v, lim, found, oid, hit, nrm = load_kats(case.key)
    assert len(o) > 100 and found.sum() > 30
    n = flat.header.n
    for k in range(len(o)):
        h = np.zeros(n); nr = np.zeros(n); i = C.c_int(-1)
        r = oracle_lib.ndo_trace(flat.blob, o[k].ctypes.data, v[k].ctypes.data, float(lim[k]),
                                 h.ctypes.data, nr.ctypes.data, C.byref(i))
        assert r == found[k] and i.value == oid[k], f"ray {k}"
        if oid[k] >= 0:
            assert same_bits(h, hit[k]) and same_bits(nr, nrm[k]), f"ray {k}"


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c.key for c in CASES])
def test_cuda_answers_the_reference_kats(case):
    import ndt_b200
    flat = load_flat(case.key)
    o, v, lim, found, oid, hit, nrm = load_kats(case.key)
    with ndt_b200.Context(0) as ctx:
        ctx.upload(flat)
        f, i, t, h, nr = ctx.trace_rays(o, v, lim)
    # ndt_b200_trace_rays runs the full trace_kd semantics for every dist_limit (the
    # "only the return value matters" shortcut is private to the DIRECTIONAL shadow test)
    assert np.array_equal(f, found)
    assert np.array_equal(i, oid)
    m = oid >= 0
    assert same_bits(h[m], hit[m]), f"{(h[m] != hit[m]).any(axis=1).sum()} hit points differ"
    assert same_bits(nr[m], nrm[m])
    d = np.sqrt(((h[m] - o[m]) ** 2).sum(axis=1))
    assert np.allclose(t[m], d, rtol=1e-12)


def _ref_replay(l):
    """get_pixel_color's sample loop for samples=1 (ndt.c:488-568), literally: six divisions per re-sample"""
    t = [0.0, 0.0, 0.0, 0.0]
    ts, clr_diff, i = 0, 256.0, 0
    while i < 1 or (i < 10000 and clr_diff > 1.0 / 256.0):
        if i > 1:
            d = [abs(t[k] / (i - 1) - (t[k] + l[k]) / i) for k in range(3)]
            clr_diff = d[0] if d[0] > (d[1] if d[1] > d[2] else d[2]) else (d[1] if d[1] > d[2] else d[2])
        for k in range(4):
            t[k] += l[k]
        ts += 1
        i += 1
    return [t[k] / ts for k in range(4)], ts


@pytest.mark.gpu
def test_sample_loop_replay_known_answers():
    """k_finish replays the reference's identical re-samples with three divisions per sample that share one
    reciprocal (core.cuh: replay_samples, divn_shared).  Bit-exact against the literal loop in IEEE double
    (Python floats) on colours across the range a frame can hold, incl. zeros, denormals, huge values, inf."""
    import ndt_b200
    rng = np.random.default_rng(7)
    cols = [rng.uniform(0, 1.5, size=(1500, 4)), rng.uniform(0, 400, size=(300, 4)),
            10.0 ** rng.uniform(-320, 300, size=(300, 4)), rng.uniform(-2, 2, size=(200, 4))]
    special = [0.0, -0.0, 1.0, 0.5, 1.0 / 256, 1.0 / 512, 2.0 / 256, 5e-324, 1e-310, 1.7e308, np.inf, 255.0, 1e-4, 3.0]
    cols.append(np.array([[a, b, c, 1.0] for a in special for b in (0.0, 0.3) for c in (0.0, 1.0 / 3)]))
    cols = np.vstack(cols)
    with ndt_b200.Context(0) as ctx:
        out, ns = ctx.replay_samples(cols)
    for k in range(len(cols)):
        want, ts = _ref_replay([float(x) for x in cols[k]])
        assert ns[k] == ts, (k, cols[k], ns[k], ts)
        got = out[k]
        for c in range(4):
            assert (got[c] == want[c]) or (got[c] != got[c] and want[c] != want[c]), (k, c, cols[k], got[c], want[c])
