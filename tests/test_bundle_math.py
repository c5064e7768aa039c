"""The exactness argument of the bundle cull (ndt_b200/csrc/warp.cuh: bundle_hit), checked in IEEE fp32 on
the CPU: numpy float32 arithmetic rounds like __fsub_rn / __fmul_rn and np.fmin / np.fmax drop a NaN operand
like fminf / fmaxf, so box_hit and bundle_hit are restated literally and the property

    bundle_hit(box, bundle of 32 rays) == False   ==>   box_hit(box, ray) == False for every ray of the bundle

is tested on random boxes and ray bundles, including axis-parallel rays (1/v = +-inf), mixed directions,
origins on a box face and degenerate boxes.  The GPU tier checks the same thing end to end (hit / id buffers
and 2244 known-answer rays stay bit-exact with the cull on)."""
import numpy as np

F = np.float32
FLT_MAX = np.finfo(np.float32).max


def box_hit(lo, hi, o, vi):
    with np.errstate(all="ignore"):
        t1 = (lo - o) * vi
        t2 = (hi - o) * vi
        tmin = F(0.0)
        tmax = F(FLT_MAX)
        for k in range(len(lo)):
            tmin = np.fmax(tmin, np.fmin(t1[k], t2[k]))
            tmax = np.fmin(tmax, np.fmax(t1[k], t2[k]))
        return bool(tmin <= tmax * F(1.000001) + F(1e-30))


def bundle_hit(lo, hi, o_lo, o_hi, v_lo, v_hi):
    with np.errstate(all="ignore"):
        tmin = F(0.0)
        tmax = F(FLT_MAX)
        for i in range(len(lo)):
            l, h = lo[i], hi[i]
            if v_lo[i] > 0 and v_hi[i] < FLT_MAX:
                a, d = l - o_hi[i], h - o_lo[i]
                tmin = np.fmax(tmin, np.fmin(a * v_lo[i], a * v_hi[i]))
                tmax = np.fmin(tmax, np.fmax(d * v_lo[i], d * v_hi[i]))
            elif v_hi[i] < 0 and v_lo[i] > -FLT_MAX:
                a, d = l - o_hi[i], h - o_lo[i]
                tmin = np.fmax(tmin, np.fmin(d * v_lo[i], d * v_hi[i]))
                tmax = np.fmin(tmax, np.fmax(a * v_lo[i], a * v_hi[i]))
            elif v_lo[i] == v_hi[i] and (v_lo[i] > FLT_MAX or v_lo[i] < -FLT_MAX):
                if o_lo[i] > h:
                    tmax = F(-np.inf)
                if o_hi[i] < l:
                    tmin = F(np.inf)
        return bool(tmin <= tmax * F(1.000001) + F(1e-30))


def make_bundle(rng, n, kind):
    """32 rays: origins and directions around a common ray; `kind` picks the awkward cases"""
    o0 = rng.uniform(-30, 30, n)
    d0 = rng.normal(size=n)
    spread_o = rng.choice([0.0, 1e-3, 0.5, 5.0])
    spread_d = rng.choice([0.0, 1e-3, 0.05, 1.0])
    o = o0 + rng.normal(size=(32, n)) * spread_o
    d = d0 + rng.normal(size=(32, n)) * spread_d
    if kind == 1:                       # some axes exactly parallel for every ray
        z = rng.random(n) < 0.4
        d[:, z] = 0.0
    elif kind == 2:                     # parallel for some rays only, with both signs of zero
        z = rng.random((32, n)) < 0.2
        d[z] = rng.choice([0.0, -0.0], size=int(z.sum()))
    elif kind == 3:                     # origins exactly on box faces come from the caller
        pass
    d /= np.sqrt((d * d).sum(axis=1, keepdims=True)) + 1e-300
    with np.errstate(all="ignore"):
        vi = (1.0 / d).astype(F)
    return o.astype(F), vi


def test_bundle_cull_never_drops_what_a_ray_would_keep():
    rng = np.random.default_rng(20261018)
    dropped = kept = 0
    for trial in range(4000):
        n = int(rng.choice([3, 4, 6, 8, 10]))
        kind = int(rng.integers(0, 4))
        o, vi = make_bundle(rng, n, kind)
        c = rng.uniform(-30, 30, n)
        half = rng.uniform(0, 6, n) * rng.choice([1.0, 0.0], size=n, p=[0.9, 0.1])
        lo = (c - half).astype(F)
        hi = (c + half).astype(F)
        if kind == 3:
            k = int(rng.integers(n))
            o[:, k] = rng.choice([lo[k], hi[k]])
        o_lo, o_hi = o.min(axis=0), o.max(axis=0)
        with np.errstate(all="ignore"):
            v_lo, v_hi = np.fmin.reduce(vi, axis=0), np.fmax.reduce(vi, axis=0)
        if not bundle_hit(lo, hi, o_lo, o_hi, v_lo, v_hi):
            dropped += 1
            for r in range(32):
                assert not box_hit(lo, hi, o[r], vi[r]), (trial, r, kind)
        else:
            kept += 1
    assert dropped > 500 and kept > 500, (dropped, kept)     # the property was exercised on both sides


def test_bundle_cull_is_as_sharp_as_the_ray_test_for_one_ray():
    """a bundle of one ray: the interval test degenerates to box_hit itself (the cull loses nothing there)"""
    rng = np.random.default_rng(5)
    for trial in range(2000):
        n = int(rng.choice([3, 6, 8]))
        o, vi = make_bundle(rng, n, int(rng.integers(0, 2)))
        o, vi = o[:1], vi[:1]
        c = rng.uniform(-30, 30, n)
        half = rng.uniform(0, 6, n)
        lo, hi = (c - half).astype(F), (c + half).astype(F)
        assert bundle_hit(lo, hi, o[0], o[0], vi[0], vi[0]) == box_hit(lo, hi, o[0], vi[0]) or \
            (bundle_hit(lo, hi, o[0], o[0], vi[0], vi[0]) and not box_hit(lo, hi, o[0], vi[0]))
