/*
 * emu_driver.cpp -- TEST INFRASTRUCTURE: runs the per-ray device code
 * (ndt_b200/csrc/core.cuh, wave.cuh) on the CPU, one ray at a time, with the
 * same generation loop / record fold the CUDA kernels use (kernels.cu).  It
 * lets the CPU-only test tier compare the wavefront formulation with the
 * oracle bit for bit; it is never part of the product library.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "wave.cuh"

using namespace ndt;

template <int NP>
static int run(const void *blob, int x0, int y0, int tw, int th, double *f64, uint8_t *u8,
               uint8_t *hit, int32_t *id, double *dep, uint64_t *stats, int eye_override = 0)
{
    const ndt_flat_header *h = (const ndt_flat_header *)blob;
    const char *b = (const char *)blob;
    Scene s;
    s.cam = (const double *)(b + h->off_camera);
    s.aabb = (const double *)(b + h->off_aabb);
    s.bs = (const double *)(b + h->off_bspheres);
    s.geom = (const double *)(b + h->off_geom);
    s.obj = (const ndt_flat_object *)(b + h->off_objects);
    s.nodes = (const ndt_flat_node *)(b + h->off_nodes);
    s.leaf = (const int32_t *)(b + h->off_leaf_refs);
    s.inf = (const int32_t *)(b + h->off_inf);
    s.lights = (const ndt_flat_light *)(b + h->off_lights);
    s.n = h->n; s.n_items = h->n_items; s.n_objects = h->n_objects; s.n_nodes = h->n_nodes;
    s.n_inf = h->n_inf; s.n_lights = h->n_lights;
    s.max_optic_depth = h->max_optic_depth; s.specular = h->specular; s.use_focal = h->use_focal;
    s.width = h->width; s.height = h->height;
    for (int k = 0; k < 4; ++k) s.bg[k] = h->bg[k];
    for (int k = 0; k < 3; ++k) s.ambient[k] = h->ambient[k];
    s.focal_scale = h->focal_scale;
    s.view = h->off_view ? (const double *)(b + h->off_view) : nullptr;
    s.cam_type = h->cam_type; s.stereo_mode = h->stereo_mode; s.view_eyes = h->view_eyes;
    s.eye_override = eye_override; s.cam_dist = h->cam_dist; s.any_boxed = 0;

    const int bpr = (tw + 7) / 8, brows = (th + 3) / 4, n0 = bpr * brows * 32;
    std::vector<RayRec> rec(n0);
    std::vector<RayIn<NP>> rays;
    uint32_t words = (h->n_items + 31) / 32; if (!words) words = 1;
    std::vector<uint32_t> bits(words, 0xffffffffu);
    Mailbox mb;
    mb.bits = bits.data(); mb.stride = 1; mb.slot = 0; mb.words = words;
    mb.group_shift = 0; while ((words >> mb.group_shift) >= 64) ++mb.group_shift;
    mb.dirty = ~0ull;
    Tally<true> tally;
    uint64_t shadow = 0;
    int ovf = 0;
    std::vector<int> gstart, gcount;
    int start = 0, count = n0;
    while (count > 0) {
        gstart.push_back(start); gcount.push_back(count);
        const int gen = (int)gstart.size() - 1;
        for (int r = 0; r < count; ++r) {
            double o[NP], v[NP], frac = 1.0; int depth = s.max_optic_depth;
            int tx = 0, ty = 0; bool active = true;
            if (gen == 0) {
                int blk = r >> 5, lane = r & 31;
                tx = (blk % bpr) * 8 + (lane & 7); ty = (blk / bpr) * 4 + (lane >> 3);
                active = tx < tw && ty < th;
                if (active) active = primary_ray<NP>(s, x0 + tx, y0 + ty, o, v);
            } else {
                const RayIn<NP> &in = rays[start + r - n0];
                for (int i = 0; i < NP; ++i) { o[i] = in.o[i]; v[i] = in.v[i]; }
                frac = in.frac; depth = in.depth;
            }
            RayRec rc; memset(&rc, 0, sizeof rc);
            rc.child_refl = rc.child_refr = CHILD_NONE;
            if (!active) {
                rc.flags = REC_UNTRACED;
                if (gen == 0 && tx < tw && ty < th) {     /* a pixel of the frame that is not traced */
                    size_t p = (size_t)ty * tw + tx;
                    if (hit) hit[p] = 0;
                    if (id) id[p] = -1;
                    if (dep) dep[p] = 0.0;
                }
            }
            if (active) {
                Spawn<NP> sp; int ph = 0, pid = -1; double pd = -1; uint32_t nsh = 0;
                process_ray<NP, true>(s, mb, o, v, frac, depth, rc, sp, ph, pid, pd, nsh, ovf, tally);
                rc.nrays = 1 + nsh; shadow += nsh;
                if (sp.want_refl == 2) rc.child_refl = CHILD_BLACK;
                if (sp.want_refr == 2) rc.child_refr = CHILD_BLACK;
                for (int k = 0; k < 2; ++k) {
                    int want = k ? sp.want_refr : sp.want_refl;
                    if (want != 1) continue;
                    RayIn<NP> q;
                    for (int i = 0; i < NP; ++i) { q.o[i] = sp.origin[i]; q.v[i] = k ? sp.refr_dir[i] : sp.refl_dir[i]; }
                    q.frac = k ? sp.refr_frac : sp.refl_frac; q.depth = depth - 1; q.pad = 0;
                    int slot = n0 + (int)rays.size();
                    rays.push_back(q);
                    if (k) rc.child_refr = slot; else rc.child_refl = slot;
                }
                if (gen == 0) {
                    size_t p = (size_t)ty * tw + tx;
                    if (hit) hit[p] = (uint8_t)ph;
                    if (id) id[p] = pid;
                    if (dep) dep[p] = (pid >= 0 && pd > EPS) ? 1.0 / pd : 0.0;
                }
            }
            rec[start + r] = rc;
        }
        start += count;
        count = n0 + (int)rays.size() - start;
        rec.resize(start + count);
    }
    for (int g = (int)gstart.size() - 1; g >= 1; --g)
        for (int i = 0; i < gcount[g]; ++i) {
            RayRec &r = rec[gstart[g] + i];
            resolve_rec(r, r.child_refl >= 0 ? &rec[r.child_refl] : nullptr,
                        r.child_refr >= 0 ? &rec[r.child_refr] : nullptr, h->specular);
        }
    uint64_t rays_ref = 0, samples = 0, traced = 0;
    for (int p = 0; p < tw * th; ++p) {
        int tx = p % tw, ty = p / tw;
        int slot = ((ty >> 2) * bpr + (tx >> 3)) * 32 + ((ty & 3) << 3) + (tx & 7);
        RayRec r = rec[slot];
        resolve_rec(r, r.child_refl >= 0 ? &rec[r.child_refl] : nullptr,
                    r.child_refr >= 0 ? &rec[r.child_refr] : nullptr, h->specular);
        double l[4] = { r.clr[0], r.clr[1], r.clr[2], r.alpha }, o[4] = { 0.0, 0.0, 0.0, 0.0 };
        int ns = (r.flags & REC_UNTRACED) ? 0 : replay_samples(l, o);
        if (f64) memcpy(f64 + 4 * (size_t)p, o, sizeof o);
        if (u8) for (int k = 0; k < 4; ++k) u8[4 * (size_t)p + k] = d2c(o[k]);
        rays_ref += (uint64_t)r.nrays * ns; samples += ns;
        traced += (r.flags & REC_UNTRACED) ? 0 : 1;
    }
    if (stats) {
        stats[0] = traced; stats[1] = rays.size(); stats[2] = shadow;
        stats[3] = rays_ref; stats[4] = samples; stats[5] = tally.f; stats[6] = gstart.size(); stats[7] = ovf;
    }
    return 0;
}

template <int NP>
static int run_view(const void *blob, int x0, int y0, int tw, int th, double *f64, uint8_t *u8,
                    uint8_t *hit, int32_t *id, double *dep, uint64_t *stats)
{
    const ndt_flat_header *h = (const ndt_flat_header *)blob;
    if (h->stereo_mode != NDT_ANAGLYPH_3D) return run<NP>(blob, x0, y0, tw, th, f64, u8, hit, id, dep, stats);
    /* ANAGLYPH_3D (ndt.c:634-646), like ndt_b200_launch_tile: one pass per eye, then the mix */
    const size_t px = (size_t)tw * th;
    std::vector<double> l(px * 4), r(px * 4);
    uint64_t s1[8] = {0}, s2[8] = {0};
    run<NP>(blob, x0, y0, tw, th, l.data(), nullptr, hit, id, dep, s1, 1);
    run<NP>(blob, x0, y0, tw, th, r.data(), nullptr, nullptr, nullptr, nullptr, s2, 2);
    for (size_t p = 0; p < px; ++p) {
        double o[4];
        o[0] = 0.299 * l[4 * p] + 0.587 * l[4 * p + 1] + 0.114 * l[4 * p + 2];
        o[1] = 0;
        o[2] = 0.299 * r[4 * p] + 0.587 * r[4 * p + 1] + 0.114 * r[4 * p + 2];
        o[3] = 1.0;
        if (f64) memcpy(f64 + 4 * p, o, sizeof o);
        if (u8) for (int k = 0; k < 4; ++k) u8[4 * p + k] = d2c(o[k]);
    }
    if (stats) {
        for (int k = 0; k < 6; ++k) stats[k] = s1[k] + s2[k];
        stats[6] = s1[6] > s2[6] ? s1[6] : s2[6];
        stats[7] = s1[7] | s2[7];
    }
    return 0;
}

extern "C" int emu_render(const void *blob, int x0, int y0, int tw, int th, double *f64, uint8_t *u8,
                          uint8_t *hit, int32_t *id, double *dep, uint64_t *stats)
{
    const ndt_flat_header *h = (const ndt_flat_header *)blob;
    switch (h->npad) {
    case 4: return run_view<4>(blob, x0, y0, tw, th, f64, u8, hit, id, dep, stats);
    case 6: return run_view<6>(blob, x0, y0, tw, th, f64, u8, hit, id, dep, stats);
    case 8: return run_view<8>(blob, x0, y0, tw, th, f64, u8, hit, id, dep, stats);
    case 10: return run_view<10>(blob, x0, y0, tw, th, f64, u8, hit, id, dep, stats);
    case 12: return run_view<12>(blob, x0, y0, tw, th, f64, u8, hit, id, dep, stats);
    }
    return -1;
}
