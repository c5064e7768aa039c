/*
 * wave.cuh -- one ray of one bounce generation: nearest hit, lighting with
 * shadow rays, and the rays it spawns.  Reference: get_ray_color (ndt.c:329-450)
 * and apply_lights (ndt.c:71-326).
 *
 * The reference is recursive: a ray's colour is finished only after its
 * reflection and refraction children return (ndt.c:398-428).  Here every ray
 * of a generation writes a RayRec with its LOCAL colour (apply_lights result,
 * background, or black), its reflectivity and the slots of the children it
 * queued; after the last generation the records are folded from the deepest
 * generation back to the primaries with exactly the reference's expressions
 * and order (resolve_rec below), so the fp64 result does not depend on which
 * thread traced what.
 */
#pragma once
#include "core.cuh"

namespace ndt {

constexpr int CHILD_NONE = -1;
constexpr int CHILD_BLACK = -2;   /* child cut by pixel_frac < 1/512 or depth 0 (ndt.c:336-341): colour (0,0,0), no trace */

struct RayRec {
    double clr[3];
    double alpha;
    double h[3];          /* get_reflect (ndt.c:383) */
    int32_t child_refl;
    int32_t child_refr;
    uint32_t nrays;       /* trace_kd calls of this ray's subtree (after resolve) */
    uint32_t flags;       /* bit0: shaded (hit an object farther than EPSILON) */
};                        /* 80 bytes */

template <int NP> struct RayIn {  /* queue entry of generation >= 1 */
    double o[NP];
    double v[NP];
    double frac;          /* pixel_frac (ndt.c:330) */
    int32_t depth;        /* max_depth (ndt.c:331) */
    int32_t pad;
};

/* what a traced ray asks the caller to enqueue */
template <int NP> struct Spawn {
    double origin[NP];    /* the hit point (ndt.c:400, 424) */
    double refl_dir[NP], refr_dir[NP];
    double refl_frac, refr_frac;
    int want_refl, want_refr;   /* 0 none, 1 queue it, 2 black */
};

/* get_ray_color for ONE ray without the recursion.  Precondition (checked by
 * whoever queued the ray): frac >= 1/512 and depth > 0.  prim_* are filled for
 * every ray; only generation 0 stores them. */
template <int NP, bool CNT>
NDT_FN void process_ray(const Scene &sc, Mailbox &mb, const double *src, const double *look,
                        double frac, int depth, RayRec &rec, Spawn<NP> &sp,
                        int &prim_hit, int &prim_id, double &prim_dist,
                        uint32_t &n_shadow, int &overflow, Tally<CNT> &tally)
{
    const int n = sc.n;
#ifdef NDT_EXP_NO_LIGHTS
    const int nl = 0;             /* experiment: extend-only kernel (register / time budget of the trace) */
#else
    const int nl = sc.n_lights;
#endif
    double Hp[NP], Hn[NP];        /* the ray's own hit point and normal */
    double clr0 = 0, clr1 = 0, clr2 = 0;
    double hr = 0, hg = 0, hb = 0;        /* colour */
    double rr = 0, rg = 0, rb = 0;        /* reflectivity used for specular */
    bool shaded = false, transparent = false;
    int oid = -1;
    n_shadow = 0;
    sp.want_refl = sp.want_refr = 0;
    rec.child_refl = rec.child_refr = CHILD_NONE;
    rec.flags = 0;

    /* per-light state that has to survive the shadow trace */
    double light_vec[NP], rev_light[NP];
    double ldist2 = 1.0;

    for (int it = -1; it < nl; ++it) {
        double ro[NP], rv[NP];
        double limit = -1.0;
        const ndt_flat_light *L = nullptr;
        int ltype = -1;
        if (it < 0) {
            vcopy<NP>(ro, src);
            vcopy<NP>(rv, look);
        } else {
            if (!shaded) break;
            L = sc.lights + it;
            ltype = NDT_LDG(&L->type);
            const double lr = NDT_LDG(&L->rgb[0]), lg = NDT_LDG(&L->rgb[1]), lb = NDT_LDG(&L->rgb[2]);
            if (ltype == NDT_L_AMBIENT) {                           /* ndt.c:105-111 */
                clr0 += hr * lr; clr1 += hg * lg; clr2 += hb * lb;
                continue;
            }
            const double *lv = sc.geom + NDT_LDG(&L->vec_off);
            double lgt_pos[NP];
            vload<NP>(lgt_pos, lv);
            /* side test, ndt.c:149-169 */
            if (ltype == NDT_L_DIRECTIONAL) {
                vload<NP>(rev_light, lv + 2 * NP);                  /* unit(-dir), hoisted */
            } else {
                vsub<NP>(lgt_pos, Hp, rev_light);
                vunit<NP>(rev_light);
            }
            double rev_view[NP];
            vsub<NP>(src, Hp, rev_view);
            double d1 = vdot<NP>(rev_light, Hn);
            double d2 = vdot<NP>(rev_view, Hn);
            tally.add(8 * n + 2);
            if ((d1 * d2) <= 0) continue;
            if (ltype == NDT_L_DIRECTIONAL) {                       /* ndt.c:230-240 */
                NDT_UNROLL
                for (int i = 0; i < NP; ++i) ro[i] = NDT_LDG(lv + 3 * NP + i) + Hp[i];
                vcopy<NP>(rv, rev_light);
                limit = 0.0;
                ldist2 = 1.0;
            } else {                                                /* ndt.c:184-211 */
                limit = vdist<NP>(Hp, lgt_pos);
                limit += EPS;
                vsub<NP>(Hp, lgt_pos, light_vec);
                ldist2 = vdot<NP>(light_vec, light_vec);
                vunit<NP>(light_vec);
                tally.add(9 * n + 2);
                if (ltype == NDT_L_SPOT) {
                    double ldir[NP];
                    vload<NP>(ldir, lv + NP);
                    double a = vangle<NP>(ldir, light_vec);
                    if ((a * 180.0 / PI) > NDT_LDG(&L->angle)) continue;
                }
                vcopy<NP>(ro, lgt_pos);
                vcopy<NP>(rv, light_vec);
            }
            ++n_shadow;
        }

        Hit T;
        trace_kd<NP, CNT>(sc, mb, ro, rv, limit, T, overflow, tally, ltype == NDT_L_DIRECTIONAL);

        if (it < 0) {
            /* ndt.c:357-376 */
            oid = T.id;
            double trace_dist = -1;
            if (oid >= 0) {
                materialise<NP>(sc, T.win, ro, rv, Hp, Hn);
                trace_dist = vdist<NP>(Hp, src);
                tally.add(3 * n);
            }
            prim_id = oid;
            prim_dist = (oid >= 0) ? trace_dist : -1.0;
            prim_hit = (oid >= 0 && trace_dist > EPS) ? 1 : 0;
            shaded = prim_hit != 0;
            if (!shaded) break;
            const ndt_flat_object *fo = sc.obj + oid;
            hr = NDT_LDG(&fo->rgb[0]); hg = NDT_LDG(&fo->rgb[1]); hb = NDT_LDG(&fo->rgb[2]);
            rec.h[0] = NDT_LDG(&fo->refl[0]); rec.h[1] = NDT_LDG(&fo->refl[1]); rec.h[2] = NDT_LDG(&fo->refl[2]);
            if (sc.specular) { rr = rec.h[0]; rg = rec.h[1]; rb = rec.h[2]; }
            transparent = (NDT_LDG(&fo->flags) & NDT_OF_TRANSPARENT) != 0;
            clr0 = hr * sc.ambient[0];                               /* ndt.c:88-92 */
            clr1 = hg * sc.ambient[1];
            clr2 = hb * sc.ambient[2];
            continue;
        }

        /* a shadow ray came back: ndt.c:212-310 */
        double lhn[NP];     /* light_hit_normal */
        if (ltype == NDT_L_DIRECTIONAL) {
            if (T.found) continue;
            vload<NP>(light_vec, sc.geom + NDT_LDG(&L->vec_off) + NP);   /* ndt.c:252 */
            vcopy<NP>(lhn, Hn);
        } else {
            if (!T.found || T.id != oid) continue;
            double lhp[NP];
            materialise<NP>(sc, T.win, ro, rv, lhp, lhn);
            double dist = vdist<NP>(Hp, lhp);
            tally.add(3 * n);
            if (dist > EPS) continue;
        }
        const double lr = NDT_LDG(&L->rgb[0]), lg = NDT_LDG(&L->rgb[1]), lb = NDT_LDG(&L->rgb[2]);
        double angle = vangle<NP>(Hn, light_vec);
        if (angle > PI / 2.0) angle = PI - angle;
        double light_scale = cos(angle) / ldist2;
        tally.add(6 * n + 8);
        if (!transparent) {
            clr0 += hr * lr * light_scale;
            clr1 += hg * lg * light_scale;
            clr2 += hb * lb * light_scale;
        }
        if (sc.specular) {                                           /* ndt.c:277-310 */
            double lref[NP], rev_look[NP];
            vreflect<NP>(light_vec, lhn, lref, 0.5);
            vunit<NP>(lref);
            vscale<NP>(look, -1, rev_look);
            vunit<NP>(rev_look);
            double rv_ = vdot<NP>(lref, rev_look);
            rv_ = ref_max(0, rv_);
            double rvn = pow(rv_, 50);
            double ml = NDT_LDG(&L->max_rgb);
            clr0 += rr * lr / ml * rvn;
            clr1 += rg * lg / ml * rvn;
            clr2 += rb * lb / ml * rvn;
            tally.add(16 * n + 14);
        }
    }

    if (!shaded) {
        /* ndt.c:436-442 */
        rec.clr[0] = sc.bg[0]; rec.clr[1] = sc.bg[1]; rec.clr[2] = sc.bg[2];
        rec.alpha = sc.bg[3];
        rec.h[0] = rec.h[1] = rec.h[2] = 0.0;
        return;
    }
    rec.flags = 1u;
    rec.clr[0] = clr0; rec.clr[1] = clr1; rec.clr[2] = clr2;
    rec.alpha = 1.0;

    /* children: ndt.c:383-429 */
    const double h0 = rec.h[0], h1 = rec.h[1], h2 = rec.h[2];
    const double contrib = ref_max(h0, ref_max(h1, h2));
    vcopy<NP>(sp.origin, Hp);
    if (contrib > 0) {
        if (h0 != 0.0 || h1 != 0.0 || h2 != 0.0) {
            sp.refl_frac = contrib * frac;
            if (sp.refl_frac < (1.0 / 512.0) || depth - 1 <= 0) {
                sp.want_refl = 2;
            } else {
                sp.want_refl = 1;
                vreflect<NP>(look, Hn, sp.refl_dir, 1.0);
                vunit<NP>(sp.refl_dir);
                tally.add(10 * n + 4);
            }
        }
    }
    if (transparent) {
        sp.refr_frac = (1 - contrib) * frac;
        if (sp.refr_frac < (1.0 / 512.0) || depth - 1 <= 0) {
            sp.want_refr = 2;
        } else {
            sp.want_refr = 1;
            vrefract<NP>(look, Hn, sp.refr_dir, NDT_LDG(&(sc.obj + oid)->refract_index));
            vunit<NP>(sp.refr_dir);
            tally.add(30 * n + 20);
        }
    }
}

/* fold the children of one record into it: ndt.c:398-428.  `child` maps a
 * slot to the (already resolved) record of the next generation. */
NDT_FN void resolve_rec(RayRec &r, const RayRec *child_refl, const RayRec *child_refr, int specular)
{
    if (r.child_refl != CHILD_NONE) {
        double ref0 = 0.0, ref1 = 0.0, ref2 = 0.0;
        if (child_refl) { ref0 = child_refl->clr[0]; ref1 = child_refl->clr[1]; ref2 = child_refl->clr[2]; r.nrays += child_refl->nrays; }
        if (specular) {
            r.clr[0] = (1 - r.h[0]) * (r.clr[0]) + (r.h[0]) * ref0;
            r.clr[1] = (1 - r.h[1]) * (r.clr[1]) + (r.h[1]) * ref1;
            r.clr[2] = (1 - r.h[2]) * (r.clr[2]) + (r.h[2]) * ref2;
        } else {
            r.clr[0] += r.h[0] * ref0;
            r.clr[1] += r.h[1] * ref1;
            r.clr[2] += r.h[2] * ref2;
        }
        r.alpha = 1.0;
    }
    if (r.child_refr != CHILD_NONE) {
        double ref0 = 0.0, ref1 = 0.0, ref2 = 0.0;
        if (child_refr) { ref0 = child_refr->clr[0]; ref1 = child_refr->clr[1]; ref2 = child_refr->clr[2]; r.nrays += child_refr->nrays; }
        r.clr[0] += (1.0 - r.h[0]) * ref0;
        r.clr[1] += (1.0 - r.h[1]) * ref1;
        r.clr[2] += (1.0 - r.h[2]) * ref2;
        r.alpha = 1.0;
    }
}

} /* namespace ndt */
