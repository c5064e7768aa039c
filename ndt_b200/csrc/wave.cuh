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

constexpr uint32_t REC_UNTRACED = 2u;
constexpr int CHILD_NONE = -1;
constexpr int CHILD_BLACK = -2;   /* child cut by pixel_frac < 1/512 or depth 0 (ndt.c:336-341): colour (0,0,0), no trace */

struct alignas(16) RayRec {
    double clr[3];
    double alpha;
    double h[3];          /* get_reflect (ndt.c:383) */
    int32_t child_refl;
    int32_t child_refr;
    uint32_t nrays;       /* trace_kd calls of this ray's subtree (after resolve) */
    uint32_t flags;       /* bit0: shaded (hit an object farther than EPSILON); REC_UNTRACED: no ray was traced
                             (tile padding, HIDEF_3D blanking rows ndt.c:619-626) */
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

/* get_ray_color (ndt.c:329-450) + apply_lights (ndt.c:71-326) for ONE ray, cut
 * into steps around the nearest-hit queries so that the same code serves the
 * scalar driver (process_ray below: emulation, counting build) and the warp
 * driver (warp.cuh), which runs the queries of 32 rays together:
 *   it = -1            the ray itself
 *   it = 0..n_lights-1 one shadow query per non-ambient light
 * shade_setup() says whether iteration `it` needs a query and leaves it in
 * ro/rv/limit; shade_after() consumes the answer; shade_finish() writes the
 * record and the rays to spawn.  Precondition (checked by whoever queued the
 * ray): frac >= 1/512 and depth > 0. */
template <int NP> struct Shade {
    double Hp[NP], Hn[NP];        /* the ray's own hit point and normal */
    double clr0, clr1, clr2;
    double hr, hg, hb;            /* colour */
    double rr, rg, rb;            /* reflectivity used for specular */
    bool shaded, transparent;
    int oid;
    /* per-light state that has to survive the shadow trace */
    double light_vec[NP], rev_light[NP];
    double ldist2;
    /* the query of the current iteration */
    double ro[NP], rv[NP];
    double limit;
    int ltype;
    const ndt_flat_light *L;
};

template <int NP> NDT_FN void shade_init(Shade<NP> &S, RayRec &rec, Spawn<NP> &sp)
{
    S.clr0 = S.clr1 = S.clr2 = 0;
    S.hr = S.hg = S.hb = 0;
    S.rr = S.rg = S.rb = 0;
    S.shaded = false; S.transparent = false;
    S.oid = -1;
    S.ldist2 = 1.0;
    S.limit = -1.0; S.ltype = -1; S.L = nullptr;
    sp.want_refl = sp.want_refr = 0;
    rec.child_refl = rec.child_refr = CHILD_NONE;
    rec.flags = 0;
}

/* true: iteration `it` wants trace_kd(S.ro, S.rv, S.limit) */
template <int NP, bool CNT>
NDT_FN bool shade_setup(const Scene &sc, Shade<NP> &S, int it, const double *src, const double *look,
                        uint32_t &n_shadow, Tally<CNT> &tally)
{
    const int n = sc.n;
    S.limit = -1.0;
    S.L = nullptr;
    S.ltype = -1;
    if (it < 0) {
        vcopy<NP>(S.ro, src);
        vcopy<NP>(S.rv, look);
        return true;
    }
    if (!S.shaded) return false;
    const ndt_flat_light *L = sc.lights + it;
    S.L = L;
    const int ltype = NDT_LDG(&L->type);
    S.ltype = ltype;
    const double lr = NDT_LDG(&L->rgb[0]), lg = NDT_LDG(&L->rgb[1]), lb = NDT_LDG(&L->rgb[2]);
    if (ltype == NDT_L_AMBIENT) {                           /* ndt.c:105-111 */
        S.clr0 += S.hr * lr; S.clr1 += S.hg * lg; S.clr2 += S.hb * lb;
        return false;
    }
    const double *lv = sc.geom + NDT_LDG(&L->vec_off);
    double lgt_pos[NP];
    vload<NP>(lgt_pos, lv);
    /* side test, ndt.c:149-169 */
    if (ltype == NDT_L_DIRECTIONAL) {
        vload<NP>(S.rev_light, lv + 2 * NP);                /* unit(-dir), hoisted */
    } else {
        vsub<NP>(lgt_pos, S.Hp, S.rev_light);
        vunit<NP>(S.rev_light);
    }
    double rev_view[NP];
    vsub<NP>(src, S.Hp, rev_view);
    double d1 = vdot<NP>(S.rev_light, S.Hn);
    double d2 = vdot<NP>(rev_view, S.Hn);
    tally.add(8 * n + 2);
    if ((d1 * d2) <= 0) return false;
    if (ltype == NDT_L_DIRECTIONAL) {                       /* ndt.c:230-240 */
        NDT_UNROLL
        for (int i = 0; i < NP; ++i) S.ro[i] = NDT_LDG(lv + 3 * NP + i) + S.Hp[i];
        vcopy<NP>(S.rv, S.rev_light);
        S.limit = 0.0;
        S.ldist2 = 1.0;
    } else {                                                /* ndt.c:184-211 */
        S.limit = vdist<NP>(S.Hp, lgt_pos);
        S.limit += EPS;
        vsub<NP>(S.Hp, lgt_pos, S.light_vec);
        S.ldist2 = vdot<NP>(S.light_vec, S.light_vec);
        vunit<NP>(S.light_vec);
        tally.add(9 * n + 2);
        if (ltype == NDT_L_SPOT) {
            double ldir[NP];
            vload<NP>(ldir, lv + NP);
            double a = vangle<NP>(ldir, S.light_vec);
            if ((a * 180.0 / PI) > NDT_LDG(&L->angle)) return false;
        }
        vcopy<NP>(S.ro, lgt_pos);
        vcopy<NP>(S.rv, S.light_vec);
    }
    ++n_shadow;
    return true;
}

/* ---- the lit part of apply_lights (ndt.c:212-310), cut where the wavefront cuts it ------------------------
 * light_geom   everything that needs N-vectors: is the shaded point the one the light sees (ndt.c:217-228,
 *              241-249), the argument of vectNd_angle's acos (vectNd.c:64-81) and the specular dot product
 *              (ndt.c:287-299).  `ro` / `rv` are the shadow query (origin = the light for POINT / SPOT),
 *              `light_vec` the unit vector from the light to the point for POINT / SPOT (it is overwritten with
 *              the light's direction for DIRECTIONAL, ndt.c:252).
 * light_libm   the scalar tail: acos, cos, pow (ndt.c:261-268, 300) -- no vectors, a few registers
 * light_apply  the colour update in the reference's order (ndt.c:269-273, 302-305) */
template <int NP, bool CNT>
NDT_FN bool light_geom(const Scene &sc, const ndt_flat_light *L, int ltype, int oid, const double *Hp, const double *Hn,
                       const double *ro, const double *rv, double *light_vec, const Hit &T, const double *look,
                       double &q, int &qok, double &rvdot, Tally<CNT> &tally)
{
    const int n = sc.n;
    double lhn[NP];     /* light_hit_normal */
    if (ltype == NDT_L_DIRECTIONAL) {
        if (T.found) return false;
        vload<NP>(light_vec, sc.geom + NDT_LDG(&L->vec_off) + NP);   /* ndt.c:252 */
        vcopy<NP>(lhn, Hn);
    } else {
        if (!T.found || T.id != oid) return false;
        double lhp[NP];
        materialise<NP>(sc, T.win, ro, rv, lhp, lhn);
        double dist = vdist<NP>(Hp, lhp);
        tally.add(3 * n);
        if (dist > EPS) return false;
    }
    {   /* vectNd_angle(Hn, light_vec) up to its acos (vectNd.c:64-81) */
        double dp = vdot<NP>(Hn, light_vec);
        double div = vnorm<NP>(Hn) * vnorm<NP>(light_vec);
        qok = fabs(div) > EPS;
        q = qok ? dp / div : 0.0;
    }
    rvdot = 0.0;
    if (sc.specular) {                                           /* ndt.c:277-299 */
        double lref[NP], rev_look[NP];
        vreflect<NP>(light_vec, lhn, lref, 0.5);
        vunit<NP>(lref);
        vscale<NP>(look, -1, rev_look);
        vunit<NP>(rev_look);
        double rv_ = vdot<NP>(lref, rev_look);
        rvdot = ref_max(0, rv_);
    }
    return true;
}

NDT_FN void light_libm(double q, int qok, double ldist2, double rvdot, int specular, double &light_scale, double &rvn)
{
    double angle = qok ? acos(q) : -1;
    if (angle > PI / 2.0) angle = PI - angle;
    light_scale = cos(angle) / ldist2;
    rvn = specular ? pow(rvdot, 50) : 0.0;
}

template <int NP> struct Shade;
template <int NP>
NDT_FN void light_apply(const Scene &sc, Shade<NP> &S, const ndt_flat_light *L, double light_scale, double rvn)
{
    const double lr = NDT_LDG(&L->rgb[0]), lg = NDT_LDG(&L->rgb[1]), lb = NDT_LDG(&L->rgb[2]);
    if (!S.transparent) {
        S.clr0 += S.hr * lr * light_scale;
        S.clr1 += S.hg * lg * light_scale;
        S.clr2 += S.hb * lb * light_scale;
    }
    if (sc.specular) {                                           /* ndt.c:300-310 */
        double ml = NDT_LDG(&L->max_rgb);
        S.clr0 += S.rr * lr / ml * rvn;
        S.clr1 += S.rg * lg / ml * rvn;
        S.clr2 += S.rb * lb / ml * rvn;
    }
}

/* the answer T to the query of iteration `it` */
template <int NP, bool CNT>
NDT_FN void shade_after(const Scene &sc, Shade<NP> &S, int it, const Hit &T, const double *src, const double *look,
                        RayRec &rec, int &prim_hit, int &prim_id, double &prim_dist, Tally<CNT> &tally)
{
    const int n = sc.n;
    if (it < 0) {
        /* ndt.c:357-376 */
        S.oid = T.id;
        double trace_dist = -1;
        if (S.oid >= 0) {
            materialise<NP>(sc, T.win, S.ro, S.rv, S.Hp, S.Hn);
            trace_dist = vdist<NP>(S.Hp, src);
            tally.add(3 * n);
        }
        prim_id = S.oid;
        prim_dist = (S.oid >= 0) ? trace_dist : -1.0;
        prim_hit = (S.oid >= 0 && trace_dist > EPS) ? 1 : 0;
        S.shaded = prim_hit != 0;
        if (!S.shaded) return;
        const ndt_flat_object *fo = sc.obj + S.oid;
        S.hr = NDT_LDG(&fo->rgb[0]); S.hg = NDT_LDG(&fo->rgb[1]); S.hb = NDT_LDG(&fo->rgb[2]);
        rec.h[0] = NDT_LDG(&fo->refl[0]); rec.h[1] = NDT_LDG(&fo->refl[1]); rec.h[2] = NDT_LDG(&fo->refl[2]);
        if (sc.specular) { S.rr = rec.h[0]; S.rg = rec.h[1]; S.rb = rec.h[2]; }
        S.transparent = (NDT_LDG(&fo->flags) & NDT_OF_TRANSPARENT) != 0;
        S.clr0 = S.hr * sc.ambient[0];                               /* ndt.c:88-92 */
        S.clr1 = S.hg * sc.ambient[1];
        S.clr2 = S.hb * sc.ambient[2];
        return;
    }

    /* a shadow ray came back: ndt.c:212-310, in three steps that the wavefront runs as separate kernels
     * (light_geom per shadow query, light_libm per shadow query at full occupancy, light_apply per ray) */
    double q, ldist2 = S.ldist2, rvdot;
    int qok;
    if (!light_geom<NP, CNT>(sc, S.L, S.ltype, S.oid, S.Hp, S.Hn, S.ro, S.rv, S.light_vec, T, look, q, qok, rvdot, tally)) return;
    double light_scale, rvn;
    light_libm(q, qok, ldist2, rvdot, sc.specular, light_scale, rvn);
    light_apply<NP>(sc, S, S.L, light_scale, rvn);
    tally.add(6 * n + 8);
    if (sc.specular) tally.add(16 * n + 14);
}

/* the record of the ray and the rays it spawns */
template <int NP, bool CNT>
NDT_FN void shade_finish(const Scene &sc, Shade<NP> &S, const double *look, double frac, int depth,
                         RayRec &rec, Spawn<NP> &sp, Tally<CNT> &tally)
{
    const int n = sc.n;
    if (!S.shaded) {
        /* ndt.c:436-442 */
        rec.clr[0] = sc.bg[0]; rec.clr[1] = sc.bg[1]; rec.clr[2] = sc.bg[2];
        rec.alpha = sc.bg[3];
        rec.h[0] = rec.h[1] = rec.h[2] = 0.0;
        return;
    }
    rec.flags = 1u;
    rec.clr[0] = S.clr0; rec.clr[1] = S.clr1; rec.clr[2] = S.clr2;
    rec.alpha = 1.0;

    /* children: ndt.c:383-429 */
    const double h0 = rec.h[0], h1 = rec.h[1], h2 = rec.h[2];
    const double contrib = ref_max(h0, ref_max(h1, h2));
    vcopy<NP>(sp.origin, S.Hp);
    if (contrib > 0) {
        if (h0 != 0.0 || h1 != 0.0 || h2 != 0.0) {
            sp.refl_frac = contrib * frac;
            if (sp.refl_frac < (1.0 / 512.0) || depth - 1 <= 0) {
                sp.want_refl = 2;
            } else {
                sp.want_refl = 1;
                vreflect<NP>(look, S.Hn, sp.refl_dir, 1.0);
                vunit<NP>(sp.refl_dir);
                tally.add(10 * n + 4);
            }
        }
    }
    if (S.transparent) {
        sp.refr_frac = (1 - contrib) * frac;
        if (sp.refr_frac < (1.0 / 512.0) || depth - 1 <= 0) {
            sp.want_refr = 2;
        } else {
            sp.want_refr = 1;
            vrefract<NP>(look, S.Hn, sp.refr_dir, NDT_LDG(&(sc.obj + S.oid)->refract_index));
            vunit<NP>(sp.refr_dir);
            tally.add(30 * n + 20);
        }
    }
}

/* scalar driver: one ray start to finish (CPU emulation, counting build).
 * prim_* are filled for every ray; only generation 0 stores them. */
template <int NP, bool CNT>
NDT_FN void process_ray(const Scene &sc, Mailbox &mb, const double *src, const double *look,
                        double frac, int depth, RayRec &rec, Spawn<NP> &sp,
                        int &prim_hit, int &prim_id, double &prim_dist,
                        uint32_t &n_shadow, int &overflow, Tally<CNT> &tally)
{
#ifdef NDT_EXP_NO_LIGHTS
    const int nl = 0;             /* experiment: extend-only kernel (register / time budget of the trace) */
#else
    const int nl = sc.n_lights;
#endif
    Shade<NP> S;
    shade_init<NP>(S, rec, sp);
    n_shadow = 0;
    for (int it = -1; it < nl; ++it) {
        if (it >= 0 && !S.shaded) break;
        if (!shade_setup<NP, CNT>(sc, S, it, src, look, n_shadow, tally)) continue;
        Hit T;
        trace_kd<NP, CNT>(sc, mb, S.ro, S.rv, S.limit, T, overflow, tally, S.ltype == NDT_L_DIRECTIONAL);
        shade_after<NP, CNT>(sc, S, it, T, src, look, rec, prim_hit, prim_id, prim_dist, tally);
    }
    shade_finish<NP, CNT>(sc, S, look, frac, depth, rec, sp, tally);
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
