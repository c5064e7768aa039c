"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs oracle/_ref, i.e. /root/reference at build
time):   python tests/golden/make_golden.py

For every case of tests/scenes.py it stores
  <key>.ndsf.gz   the flat scene produced by ndt_b200_flatten from the
                  reference's own scene + kd-tree structures (the INPUT), and
  golden.json     SHA-256 digests of what the reference itself rendered for it:
                  fp64 RGBA framebuffer of render_image (ndt.c:900), 8-bit
                  image through pixel_d2c, and the primary-ray hit / object-id
                  buffers from camera_target_point + trace_kd; plus a few raw
                  sample pixels so a digest mismatch can be localised.
The reference has no golden vectors of its own; these
digests are the pin for oracle/ndt_oracle.c on boxes without /root/reference.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import ndt_b200  # noqa: E402
from oracle.refharness import RefHarness, rgba_f64_to_u8  # noqa: E402
from scenes import AA_CASES, AA_PARAMS, CASES, H_FOV, V_FOV  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexv(a):
    return [float(x).hex() for x in a]


def make_kats(R, flat, c, rng, n_primary=48):
    """Known-answer rays for trace_kd (object.c:683), answered by the reference itself:
    primary-like rays towards object bounding spheres, secondary rays leaving the
    hit points (random, axis-aligned with exact zeros, and straight back), and
    shadow-style rays with dist_limit > 0 and == 0 (ndt.c:172-183)."""
    h = flat.header
    n, npad = h.n, h.npad
    blob = flat.blob
    cam = np.frombuffer(blob, np.float64, 4 * npad, h.off_camera).reshape(4, npad)[:, :n]
    bs = np.frombuffer(blob, np.float64, h.n_objects * (npad + 2), h.off_bspheres).reshape(h.n_objects, npad + 2)
    kats = []

    def ask(o, v, lim):
        r, hit, nrm, oid = R.trace_ray(o, v, lim)
        kats.append({"o": hexv(o), "v": hexv(v), "limit": float(lim).hex(), "found": int(r), "id": int(oid),
                     "hit": hexv(hit), "normal": hexv(nrm)})
        return r, hit, nrm, oid

    def unit(v):
        return v / np.sqrt((v * v).sum())

    pos = cam[0].copy()
    finite = [i for i in range(h.n_items) if bs[i, npad] > 0]
    hits = []
    for k in range(n_primary):
        if finite and k % 4 != 3:
            i = finite[rng.integers(len(finite))]
            tgt = bs[i, :n] + rng.normal(size=n) * bs[i, npad] * 0.6
        else:
            tgt = cam[1] + (rng.random() - 0.5) * cam[2] + (rng.random() - 0.5) * cam[3]
        v = unit(tgt - pos)
        r, hit, nrm, oid = ask(pos, v, -1.0)
        if r and oid >= 0:
            hits.append((hit.copy(), nrm.copy(), v.copy()))
    for hit, nrm, v in hits[:24]:
        d = unit(rng.normal(size=n))
        ask(hit, d, -1.0)                                   # a bounce ray leaving the surface
        ax = np.zeros(n); ax[rng.integers(n)] = 1.0 if rng.random() < 0.5 else -1.0
        ask(hit, ax, -1.0)                                  # axis aligned: exact zeros in v
        ask(hit, -v, -1.0)                                  # straight back along the incoming ray
        lpos = hit + unit(rng.normal(size=n)) * (5 + 20 * rng.random())
        lv = unit(hit - lpos)
        dist = float(np.sqrt(((hit - lpos) ** 2).sum()))
        ask(lpos, lv, dist + 1e-4)                          # POINT-light style shadow ray
        ask(hit - 1e-4 * lv, -lv, 0.0)                      # DIRECTIONAL style: any hit
        ask(hit + 0.5 * nrm / max(1e-9, np.sqrt((nrm * nrm).sum())), -unit(nrm), -1.0)   # from just outside, straight in
    return kats


def main():
    R = RefHarness()
    rng = np.random.default_rng(20261018)
    out = {}
    for c in CASES:
        R.open_scene(c.scene)
        frames = R.scene_frames(c.dims, c.cfg) if c.scene else 300
        if frames <= 0:
            frames = 300
        R.begin_frame(c.dims, c.frame, frames, c.cfg)
        view = c.cam != 0 or c.stereo != 0
        if view:
            R.set_camera(c.cam, H_FOV, V_FOV)
        flat = ndt_b200.flatten(R.scene_ptr, R.kdtree_ptr, c.w, c.h, 128, 1, R.get_bounds_ptr,
                                stereo_mode=c.stereo, host_rotate2=R.rotate2_ptr)
        kats = [] if view else make_kats(R, flat, c, rng)
        img, _ = R.render(c.w, c.h, threads=os.cpu_count(), stereo=c.stereo)
        if view:
            # hit / id buffers of the new views: the oracle's, pinned through the fp64 image it shares with them
            hit = np.zeros((c.h, c.w), np.uint8); oid = np.zeros((c.h, c.w), np.int32)
            if c.stereo == 4:
                img[1080:1126, :, 3] = 0.0     # blanking rows: alpha is uninitialised stack in the reference (ndt.c:623)
        else:
            hit, oid, dist = R.primary(c.w, c.h)
        R.end_frame()
        if not view:
            with open(os.path.join(HERE, c.key + ".kat.json"), "w") as f:
                json.dump(kats, f)
        flat.save(os.path.join(HERE, c.key + ".ndsf.gz"))
        u8 = rgba_f64_to_u8(img)
        pts = [(0, 0), (c.h // 2, c.w // 2), (c.h - 1, c.w - 1), (c.h // 3, (2 * c.w) // 3)]
        out[c.key] = {
            "scene": c.scene, "dims": c.dims, "cfg": c.cfg, "frame": c.frame, "w": c.w, "h": c.h,
            "cam": c.cam, "stereo": c.stereo, "view": view,
            "flat_bytes": len(flat), "n_items": flat.header.n_items, "n_objects": flat.header.n_objects,
            "n_nodes": flat.header.n_nodes, "n_leaf_refs": flat.header.n_leaf_refs,
            "sha_f64": sha(img), "sha_u8": sha(u8), "sha_hit": sha(hit), "sha_id": sha(oid),
            "hit_pixels": int(hit.sum()),
            "samples": [{"y": y, "x": x, "rgba_hex": [float(v).hex() for v in img[y, x]],
                         "hit": int(hit[y, x]), "id": int(oid[y, x])} for y, x in pts],
        }
        print(c.key, "flat", len(flat), "hit px", int(hit.sum()), "kats", len(kats),
              "found", sum(k["found"] for k in kats), flush=True)
    # recursive anti-aliasing: the reference's 8-bit actual_img per (aa_diff, aa_depth)
    for c in AA_CASES:
        R.open_scene(c.scene)
        frames = R.scene_frames(c.dims, c.cfg) if c.scene else 300
        if frames <= 0:
            frames = 300
        shas = {}
        for diff, depth in AA_PARAMS:
            R.begin_frame(c.dims, c.frame, frames, c.cfg)
            flat = ndt_b200.flatten_aa(R.scene_ptr, R.kdtree_ptr, c.w, c.h, 128, 1, R.get_bounds_ptr)
            u8, _ = R.render_aa(c.w, c.h, diff, depth)
            R.end_frame()
            shas[f"{diff},{depth}"] = sha(u8)
        flat.save(os.path.join(HERE, c.key + ".ndsf.gz"))
        out[c.key] = {"scene": c.scene, "dims": c.dims, "cfg": c.cfg, "frame": c.frame, "w": c.w, "h": c.h,
                      "aa": True, "flat_bytes": len(flat), "sha_u8_by_params": shas}
        print(c.key, "flat", len(flat), flush=True)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
