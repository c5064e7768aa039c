"""tools/config3_cull_study.py [n_objects] [rays] -- ANALYSIS TOOL (numpy, CPU only): how much each exact cull of the broad
phase leaves of BASELINE config 3 (scenes/random.c, 6-D), for rays that start inside the cloud (the bounce generations).

Builds the scene with the reference (oracle/_ref), puts every finite object into ONE kd leaf (what the bounded builder
returns for this scene), flattens it, and counts per ray / per (ray, cube) pair:
  * top-level records passing the fp32-style slab test of their box, and box + bounding sphere;
  * hcube candidates with the box of the cube's bounding sphere, and with the union of its faces' boxes (what k_pack_leaf
    stores since round 2);
  * faces per candidate pair passing their world-space box + sphere (what warp_nested / hcube_one_ray intersect today);
  * the same with boxes in the CUBE'S OWN FRAME: the faces of a cube share its n edge vectors, so in the coordinates
    u = D^-1 (x - corner) every face is an axis-aligned box (thin in its fixed dimensions); the frame is rebuilt from the faces'
    bases alone (the flat scene does not carry the cube).  DESIGN.md section 10 quotes these numbers.
Nothing here is part of the product or of the tests."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ndt_b200      # noqa: E402  (flattener only: host code)
import refharness    # noqa: E402

ONELEAF = r'''
#include <stdlib.h>
#include "ndt_abi.h"
typedef struct { ndtabi_vec lower, upper; int id; void *obj_ptr; } kdb_item;
typedef struct { kdb_item **items; int n, cap; } kdb_item_list;
int oneleaf(ndtabi_kd_tree *tree, kdb_item_list *items)
{
    int dims = tree->bb_lower.n, n = items->n, nf = 0, ni = 0;
    if (!tree->root) tree->root = calloc(1, sizeof(ndtabi_kd_node));
    for (int i = 0; i < n; ++i) if (((ndtabi_object *)items->items[i]->obj_ptr)->bounds.radius >= 0.0) ++nf; else ++ni;
    tree->inf_obj_ptrs = calloc(ni ? ni : 1, sizeof(void *));
    tree->inf_obj_num = 0;
    ndtabi_kd_node *node = tree->root;
    node->obj_ids = calloc(nf ? nf : 1, sizeof(int *));
    node->objs = calloc(nf ? nf : 1, sizeof(void *));
    int k = 0;
    for (int i = 0; i < n; ++i) {
        kdb_item *it = items->items[i];
        it->id = i;
        if (((ndtabi_object *)it->obj_ptr)->bounds.radius >= 0.0) {
            node->obj_ids[k] = i; node->objs[k++] = it->obj_ptr;
            for (int d = 0; d < dims; ++d) {
                if (it->lower.v[d] < tree->bb_lower.v[d]) tree->bb_lower.v[d] = it->lower.v[d];
                if (it->upper.v[d] > tree->bb_upper.v[d]) tree->bb_upper.v[d] = it->upper.v[d];
            }
        } else tree->inf_obj_ptrs[tree->inf_obj_num++] = it->obj_ptr;
    }
    node->num = nf; node->dim = -1; node->boundary = 0; node->left = node->right = NULL;
    tree->obj_num = n;
    return 0;
}
'''


def build_flat(n_obj):
    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "oneleaf.c"); so = os.path.join(tmp, "liboneleaf.so")
    open(src, "w").write(ONELEAF)
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", so, src], check=True)
    R = refharness.RefHarness(); R.open_scene("random")
    R.begin_frame_nokd(6, 0, 300, str(n_obj))
    L = C.CDLL(so); L.oneleaf.argtypes = [C.c_void_p, C.c_void_p]
    L.oneleaf(R.kdtree_ptr, R.items_ptr)
    flat = ndt_b200.flatten(R.scene_ptr, R.kdtree_ptr, 3840, 2160, 128, 1, R.get_bounds_ptr)
    R.end_frame()
    return flat


def main():
    n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    f = build_flat(n_obj); h = f.header
    NP, n = h.npad, h.n
    dbl = lambda off, cnt: np.frombuffer(f.blob, dtype=np.float64, count=cnt, offset=off)
    nobj = h.n_objects
    obj = np.frombuffer(f.blob, dtype=np.dtype([('type', 'i4'), ('flags', 'i4'), ('rid', 'i4'), ('nax', 'i4'), ('cb', 'i4'), ('cc', 'i4'),
                                                ('goff', 'u4'), ('res', 'u4'), ('rgb', 'f8', 3), ('refl', 'f8', 3), ('ri', 'f8'), ('bsr', 'f8')]),
                        count=nobj, offset=h.off_objects)
    bs = dbl(h.off_bspheres, nobj * (NP + 2)).reshape(nobj, NP + 2)
    geom = dbl(h.off_geom, h.n_geom)

    def face(i):
        o_ = obj[i]; m = o_['nax']; g = geom[o_['goff']:]
        return g[:NP].copy(), g[NP:NP + m * NP].reshape(m, NP).copy(), g[NP + m * NP:NP + m * NP + m].copy()

    lo = np.empty((nobj, NP)); hi = np.empty((nobj, NP))          # k_pack_leaf's boxes (gen.cuh: orthotope_reach / sphere box)
    for i in range(nobj):
        if obj[i]['type'] == 3 and bs[i, NP] > 0:
            p0, B, ln = face(i)
            e0 = -2e-4 * B; e1 = (ln[:, None] + 2e-4) * B
            lo[i] = p0 + np.minimum(e0, e1).sum(0) - 0.03; hi[i] = p0 + np.maximum(e0, e1).sum(0) + 0.03
        else:
            lo[i] = bs[i, :NP] - bs[i, NP] - 0.03; hi[i] = bs[i, :NP] + bs[i, NP] + 0.03

    def slab(o, v, blo, bhi):
        with np.errstate(divide='ignore', invalid='ignore'):
            vi = 1.0 / v
            t1 = (blo[None] - o[:, None]) * vi[:, None]; t2 = (bhi[None] - o[:, None]) * vi[:, None]
            tmin = np.fmax.reduce(np.fmin(t1, t2)[..., :n], axis=2); tmax = np.fmin.reduce(np.fmax(t1, t2)[..., :n], axis=2)
        return np.fmax(tmin, 0) <= tmax

    def sph(o, v, idx):
        c = bs[idx, :NP]; r2 = bs[idx, NP + 1]
        oc = o[:, None, :] - c[None]; oc2 = (oc * oc).sum(2); voc = (v[:, None, :] * oc).sum(2); desc = voc * voc - oc2 + r2[None]
        return ~((desc < 0) | ((voc > 0) & (voc * voc > desc)))

    def cube_frame(ci):
        ob = obj[ci]; axes = []; fa = []
        for c in range(ob['cb'], ob['cb'] + ob['cc']):
            p0, B, ln = face(c); idx = []
            for b in B:
                k = next((j for j, a in enumerate(axes) if np.array_equal(a, b)), None)
                if k is None:
                    axes.append(b); k = len(axes) - 1
                idx.append(k)
            fa.append((p0, idx, ln))
        if len(axes) != n:
            return None
        D = np.array(axes)[:, :n].T
        if abs(np.linalg.det(D)) < 1e-9:
            return None
        Di = np.linalg.inv(D); ref = fa[0][0][:n]
        u0 = np.array([Di @ (p0[:n] - ref) for p0, _, _ in fa])
        shift = u0.min(0); u0 -= shift; ref = ref + D @ shift
        rown = np.sqrt((Di * Di).sum(1))
        ulo = u0.copy(); uhi = u0.copy()
        for k, (p0, idx, ln) in enumerate(fa):
            for a, j in enumerate(idx):
                ulo[k, j] = u0[k, j] - 2e-4; uhi[k, j] = u0[k, j] + ln[a] + 2e-4
        return Di, ref, ulo - (0.03 * rown + 1e-6), uhi + (0.03 * rown + 1e-6)

    rng = np.random.default_rng(5)
    top = np.arange(h.n_items)
    hc = np.flatnonzero(obj['type'][:h.n_items] == 4)
    tlo = np.array([lo[obj[i]['cb']:obj[i]['cb'] + obj[i]['cc']].min(0) for i in hc])
    thi = np.array([hi[obj[i]['cb']:obj[i]['cb'] + obj[i]['cc']].max(0) for i in hc])
    o = rng.uniform(2, 12, (n_rays, NP)); o[:, 4:] = rng.uniform(-1, 3, (n_rays, NP - 4))
    v = rng.normal(size=(n_rays, NP)); v /= np.linalg.norm(v, axis=1)[:, None]
    b_top = slab(o, v, lo[top], hi[top]); s_top = sph(o, v, top)
    print(f"{n_obj} objects ({len(hc)} hcubes, {nobj - h.n_items} nested faces), {n_rays} rays from inside the cloud")
    print(f"  top-level records per ray: box {b_top.sum() / n_rays:.1f}, box + sphere {(b_top & s_top).sum() / n_rays:.1f}")
    s_c = sph(o, v, hc)
    print(f"  hcube candidates per ray: sphere + sphere's box {(s_c & slab(o, v, lo[hc], hi[hc])).sum() / n_rays:.2f}, "
          f"sphere + union of the faces' boxes {(s_c & slab(o, v, tlo, thi)).sum() / n_rays:.2f}")
    pairs = np.argwhere(s_c & slab(o, v, tlo, thi))[:600]
    cache = {}; tw = tu = left = noframe = 0
    for r, j in pairs:
        ci = hc[j]
        if ci not in cache:
            cache[ci] = cube_frame(ci)
        fr = cache[ci]
        ch = np.arange(obj[ci]['cb'], obj[ci]['cb'] + obj[ci]['cc'])
        hw = slab(o[r:r + 1], v[r:r + 1], lo[ch], hi[ch])[0] & sph(o[r:r + 1], v[r:r + 1], ch)[0]
        tw += hw.sum()
        if fr is None:
            noframe += 1; tu += hw.sum(); left += 1
            continue
        Di, ref, ulo, uhi = fr
        uo = (Di @ (o[r, :n] - ref))[None]; uv = (Di @ v[r, :n])[None]
        with np.errstate(divide='ignore', invalid='ignore'):
            vi = 1.0 / uv
            t1 = (ulo - uo) * vi; t2 = (uhi - uo) * vi
            par = np.abs(uv) < 1e-300
            tmin = np.fmax.reduce(np.where(par, -np.inf, np.fmin(t1, t2)), axis=1)
            tmax = np.fmin.reduce(np.where(par, np.inf, np.fmax(t1, t2)), axis=1)
            out = (par & ((uo < ulo) | (uo > uhi))).any(1)
        hu = (np.fmax(tmin, 0) <= tmax) & ~out & hw
        tu += hu.sum(); left += bool(hu.any())
    m = max(1, len(pairs))
    print(f"  {len(pairs)} (ray, cube) pairs: faces per pair passing world box + sphere {tw / m:.2f}; also passing the box in the cube's "
          f"frame {tu / m:.3f}; pairs with any face left {left} ({noframe} cubes without a frame)")


if __name__ == "__main__":
    main()
