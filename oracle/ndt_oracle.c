/*
 * ndt_oracle.c -- TEST INFRASTRUCTURE: plain-C restatement of ndt's render
 * path over the flat scene (include/ndt_flat.h).  It is the checker, never the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors (SURVEY.md
 * section 4), so the pin is the reference itself: tests/test_oracle.py
 * runs the unmodified reference (oracle/_ref, built from /root/reference by
 * oracle/Makefile) and this file on the same scenes and requires bit-identical
 * fp64 framebuffers and hit/id buffers; tests/golden/ holds the digests of
 * those runs for boxes without /root/reference.
 *
 * Structure follows the reference one function at a time (recursive, one ray
 * at a time, runtime N); vectors are npad-wide double arrays and every vector
 * op works on lane pairs like the SSE2 code in vectNd.h.  Compile with
 * -ffp-contract=off and no -march so nothing is fused.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "ndt_flat.h"

#define EPS   NDT_EPS
#define EPS2  NDT_EPS2
#define INV_EPS2 (1.0 / (EPS2))      /* kd-tree.c:480 */
#define MAXD  NDT_MAX_DIM
#define REF_MAX(x, y) (((x) > (y)) ? (x) : (y))   /* image.h:33 (included before ndt.c:33) */
#define REF_MIN(x, y) (((x) < (y)) ? (x) : (y))   /* image.h:30 */

typedef struct {
    const ndt_flat_header *h;
    int n, np;
    const double *cam, *aabb, *bs, *geom;
    const ndt_flat_object *obj;
    const ndt_flat_node *nodes;
    const int32_t *leaf, *inf;
    const ndt_flat_light *lights;
} view;

typedef struct { uint64_t primary, bounce, shadow; } raycnt;

static void view_init(view *w, const void *blob)
{
    const ndt_flat_header *h = blob;
    w->h = h; w->n = h->n; w->np = h->npad;
    w->cam = NDT_FLAT_PTR(blob, const double, h->off_camera);
    w->aabb = NDT_FLAT_PTR(blob, const double, h->off_aabb);
    w->bs = NDT_FLAT_PTR(blob, const double, h->off_bspheres);
    w->geom = NDT_FLAT_PTR(blob, const double, h->off_geom);
    w->obj = NDT_FLAT_PTR(blob, const ndt_flat_object, h->off_objects);
    w->nodes = NDT_FLAT_PTR(blob, const ndt_flat_node, h->off_nodes);
    w->leaf = NDT_FLAT_PTR(blob, const int32_t, h->off_leaf_refs);
    w->inf = NDT_FLAT_PTR(blob, const int32_t, h->off_inf);
    w->lights = NDT_FLAT_PTR(blob, const ndt_flat_light, h->off_lights);
}

/* ---- vectNd.h on npad-wide arrays ------------------------------------------ */
static double vdot(const double *a, const double *b, int np)      /* vectNd.h:215-227 */
{
    double s0 = a[0] * b[0], s1 = a[1] * b[1];
    for (int i = 2; i < np; i += 2) { s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1]; }
    return s0 + s1;
}
static void vadd(const double *a, const double *b, double *r, int np) { for (int i = 0; i < np; ++i) r[i] = a[i] + b[i]; }
static void vsub(const double *a, const double *b, double *r, int np) { for (int i = 0; i < np; ++i) r[i] = a[i] - b[i]; }
static void vscale(const double *a, double s, double *r, int np) { for (int i = 0; i < np; ++i) r[i] = a[i] * s; }
static void vcopy(double *d, const double *s, int np) { memcpy(d, s, (size_t)np * sizeof(double)); }
static void vzero(double *d, int np) { memset(d, 0, (size_t)np * sizeof(double)); }
static double vnorm(const double *a, int np) { return sqrt(vdot(a, a, np)); }
static void vunit(double *a, int np)                                /* vectNd.h:323-329 */
{
    double len = vnorm(a, np);
    if (len > EPS || len < -EPS) vscale(a, 1.0 / len, a, np);
}
static double vdist(const double *a, const double *b, int np)      /* vectNd.h:331-338 */
{
    double d[MAXD];
    vsub(a, b, d, np);
    return vnorm(d, np);
}
static double vangle(const double *a, const double *b, int np)     /* vectNd.c:64-81 */
{
    double dp = vdot(a, b, np);
    double div = vnorm(a, np) * vnorm(b, np);
    if (fabs(div) > EPS) return acos(dp / div);
    return -1;
}
static double vangle3(const double *p1, const double *p2, const double *p3, int np)  /* vectNd.c:83-99 */
{
    double a[MAXD], b[MAXD];
    vsub(p1, p2, a, np);
    vsub(p3, p2, b, np);
    return vangle(a, b, np);
}
static void vproj(const double *v, const double *onto, double *r, int np)           /* vectNd.h:355-363 */
{
    double bb = vdot(onto, onto, np);
    double ab = vdot(v, onto, np);
    vscale(onto, ab / bb, r, np);
}
static void vproj_unit(const double *v, const double *onto, double *r, int np)      /* vectNd.h:346-352 */
{
    vscale(onto, vdot(v, onto, np), r, np);
}
static void vreflect(const double *u, const double *nrm, double *r, double mag, int np) /* vectNd.c:101-117 */
{
    double nu = vdot(nrm, u, np), nn = vdot(nrm, nrm, np), t[MAXD];
    vscale(nrm, (1 + mag) * nu / nn, t, np);
    vsub(u, t, r, np);
}
/* vectNd.c:119-188; unitizes `nrm` in place like the reference (line 155) */
static void vrefract(const double *u, double *nrm, double *res, double index, int np)
{
    double rev_u[MAXD], rev_n[MAXD], un[MAXD], perp[MAXD], ref_n[MAXD], ref_p[MAXD];
    vscale(u, -1, rev_u, np);
    vscale(nrm, -1, rev_n, np);
    double un_dot = vdot(rev_u, nrm, np);
    double theta_in;
    if (un_dot < 0) {
        index = 1 / index;
        theta_in = vangle(rev_u, rev_n, np);
    } else {
        theta_in = vangle(rev_u, nrm, np);
    }
    double theta_out, sin_out = sin(theta_in) / index;
    if (sin_out <= 1.0) theta_out = asin(sin_out);
    else theta_out = M_PI - theta_in;
    vunit(rev_n, np);
    vunit(nrm, np);
    vproj_unit(u, rev_n, un, np);
    vsub(u, un, perp, np);
    vunit(perp, np);
    double rn = cos(theta_out), rp = sin(theta_out);
    if (un_dot < 0) vscale(nrm, rn, ref_n, np);
    else vscale(rev_n, rn, ref_n, np);
    vscale(perp, rp, ref_p, np);
    vadd(ref_n, ref_p, res, np);
}

/* ---- bounding sphere pre-test: bounding.c:34-85 ---------------------------- */
static int bsphere_test(const view *w, int id, const double *o, const double *v, double min_dist)
{
    const int np = w->np;
    const double *b = w->bs + (size_t)id * (np + 2);
    double oc[MAXD];
    vsub(o, b, oc, np);
    double oc2 = vdot(oc, oc, np);
    if (min_dist > 0) {
        double mr = min_dist + b[np];
        if (oc2 > mr * mr) return 0;
    }
    double voc = vdot(v, oc, np);
    double voc2 = voc * voc;
    double desc = voc2 - oc2 + b[np + 1];
    if (desc < 0.0 || (voc > 0.0 && voc2 > desc)) return 0;
    return 1;
}

/* ---- per-type intersect functions (objects/<type>.c) ------------------------
 * Contract as in the reference: return 1 and fill res/normal on a hit; on a
 * miss res/normal may hold scratch values (trace() ignores them). */

static int isect_sphere(const view *w, const ndt_flat_object *fo, const double *o, const double *v,
                        double *res, double *nrm)                       /* sphere.c:57-112 */
{
    const int np = w->np;
    const double *c = w->geom + fo->geom_off;
    double r2 = c[np];
    vsub(o, c, res, np);
    double oc2 = vdot(res, res, np);
    double voc = vdot(v, res, np);
    double desc = (voc * voc) - oc2 + r2;
    if (desc < 0.0) return 0;
    double root = sqrt(desc);
    double d = -(voc + root);
    if (d < EPS) {
        d = root - voc;
        if (d < EPS) { vzero(res, w->n); vzero(nrm, w->n); return 0; }
    }
    vscale(v, d, res, np);
    vadd(o, res, res, np);
    vsub(res, c, nrm, np);
    return 1;
}

static int isect_hplane_raw(const double *p, const double *pn, const double *o, const double *v,
                            double *res, double *nrm, int n, int np)     /* hplane.c:39-75 */
{
    double pl[MAXD], d = -1;
    memcpy(nrm, pn, (size_t)n * sizeof(double));   /* vectNd_copy: n lanes, pad of nrm kept */
    vsub(p, o, pl, np);
    double pln = vdot(pl, nrm, np);
    double ln = vdot(v, nrm, np);
    if (ln > EPS || ln < -EPS) d = pln / ln;
    if (d >= EPS) {
        memcpy(res, o, (size_t)n * sizeof(double));
        vscale(v, d, pl, np);
        vadd(res, pl, res, np);
    }
    if (d < EPS) return 0;
    return 1;
}

static int isect_hdisk(const view *w, const ndt_flat_object *fo, const double *o, const double *v,
                       double *res, double *nrm)                        /* hdisk.c:61-85 */
{
    const int np = w->np;
    const double *g = w->geom + fo->geom_off;
    if (!isect_hplane_raw(g, g + np, o, v, res, nrm, w->n, np)) return 0;
    double dist = vdist(res, g, np);
    if (dist > g[2 * np] || dist < 0) return 0;
    return 1;
}

/* orthotope.c:122-148 and hcylinder.c:102-130 share this shape */
static int within_axes(const double *pt, const double *p0, const double *basis, const double *len,
                       const double *ada, int m, int np)
{
    double bc[MAXD];
    vsub(pt, p0, bc, np);
    for (int i = 0; i < m; ++i) {
        double s = vdot(bc, basis + (size_t)i * np, np);
        s = s / ada[i];
        if (s < -EPS || s > len[i] + EPS) return 0;
    }
    return 1;
}

/* the sum over axes both orthotope.c:170-193 and hcylinder.c:160-179 start with */
static void axes_PQ(const double *o, const double *v, const double *p0, const double *basis,
                    const double *ada, const double *bda, int m, int np, int n, double *P, double *Q)
{
    double sum[MAXD], sA[MAXD];
    vzero(sum, np);
    for (int i = 0; i < m; ++i) {
        double vda = vdot(v, basis + (size_t)i * np, np);
        vscale(basis + (size_t)i * np, vda / ada[i], sA, np);
        vadd(sum, sA, sum, np);
    }
    vsub(sum, v, P, np);
    vzero(sum, n);                                   /* vectNd_reset: n lanes */
    for (int i = 0; i < m; ++i) {
        double oda = vdot(o, basis + (size_t)i * np, np);
        vscale(basis + (size_t)i * np, (oda - bda[i]) / ada[i], sA, np);
        vadd(sum, sA, sum, np);
    }
    vsub(p0, o, Q, np);
    vadd(Q, sum, Q, np);
}

static void axes_normal(const double *res, const double *p0, const double *basis, int m, int np, int n,
                        double *Q, double *nrm)      /* orthotope.c:277-294, hcylinder.c:217-236 */
{
    double P[MAXD], sA[MAXD];
    vsub(res, p0, P, np);
    vzero(Q, n);
    for (int i = 0; i < m; ++i) {
        vproj(P, basis + (size_t)i * np, sA, np);
        vadd(Q, sA, Q, np);
    }
    vsub(P, Q, nrm, np);
}

static int isect_orthotope(const view *w, const ndt_flat_object *fo, const double *o, const double *v,
                           double *res, double *nrm)                    /* orthotope.c:150-302 */
{
    const int np = w->np, m = fo->n_axes;
    const double *p0 = w->geom + fo->geom_off, *basis = p0 + np;
    const double *len = basis + (size_t)m * np, *bdb = len + m, *bdp = bdb + m;
    double P[MAXD], Q[MAXD], sA[MAXD];
    int ret = 0;
    axes_PQ(o, v, p0, basis, bdb, bdp, m, np, w->n, P, Q);
    double qa = vdot(P, P, np);
    double qb = vdot(P, Q, np);
    qb *= 2;
    double qc = vdot(Q, Q, np);
    qc -= EPS;
    double det = qb * qb - 4 * qa * qc;
    if (det >= 0.0 && fabs(qa) > EPS) {
        double root = sqrt(det);
        double hiq = 0.5 / qa;
        double t1 = (-qb + root) * hiq;
        double t2 = (-qb - root) * hiq;
        if (t2 > EPS) {
            vscale(v, t2, sA, np);
            vadd(o, sA, res, np);
            if (within_axes(res, p0, basis, len, bdb, m, np)) ret = 1;
        }
        if (ret == 0 && t1 > EPS) {
            vscale(v, t1, sA, np);
            vadd(o, sA, res, np);
            if (within_axes(res, p0, basis, len, bdb, m, np)) ret = 1;
        }
    }
    if (ret == 0) {
        double t = -1.0;
        if (fabs(qa) < EPS) {
            if (fabs(qb) < EPS) t = -qc / qb;        /* sic: orthotope.c:236-242 */
            else t = -1.0;
        } else {
            t = -qb / (2 * qa);
        }
        if (t < EPS) return 0;
        double dist = qa * t * t + qb * t + qc;
        if (fabs(dist) > EPS) return 0;
        vscale(v, t, sA, np);
        vadd(o, sA, res, np);
        if (within_axes(res, p0, basis, len, bdb, m, np)) ret = 1;
    }
    if (ret) axes_normal(res, p0, basis, m, np, w->n, Q, nrm);
    return ret;
}

static int isect_hcylinder(const view *w, const ndt_flat_object *fo, const double *o, const double *v,
                           double *res, double *nrm)                    /* hcylinder.c:132-244 */
{
    const int np = w->np, m = fo->n_axes;
    const double *p0 = w->geom + fo->geom_off, *axes = p0 + np;
    const double *len = axes + (size_t)m * np, *ada = len + m, *bda = ada + m;
    double radius = bda[m];
    int no_end = (fo->flags & NDT_OF_NO_END_TEST) != 0;
    double P[MAXD], Q[MAXD], sA[MAXD];
    int ret = 0;
    axes_PQ(o, v, p0, axes, ada, bda, m, np, w->n, P, Q);
    double qa = vdot(P, P, np);
    double qb = vdot(P, Q, np);
    qb *= 2;
    double qc = vdot(Q, Q, np);
    qc -= radius * radius;
    double det = qb * qb - 4 * qa * qc;
    if (det < 0.0) return 0;
    double root = sqrt(det);
    double t1 = (-qb + root) / (2 * qa);
    double t2 = (-qb - root) / (2 * qa);
    if (t2 > EPS) {
        vscale(v, t2, sA, np);
        vadd(o, sA, res, np);
        if (no_end || within_axes(res, p0, axes, len, ada, m, np)) ret = 1;
    }
    if (ret == 0 && t1 > EPS) {
        vscale(v, t1, sA, np);
        vadd(o, sA, res, np);
        if (no_end || within_axes(res, p0, axes, len, ada, m, np)) ret = 1;
    }
    if (ret) axes_normal(res, p0, axes, m, np, w->n, Q, nrm);
    return ret;
}

static int isect_cylinder(const view *w, const ndt_flat_object *fo, const double *o, const double *v,
                          double *res, double *nrm)                     /* cylinder.c:104-210 */
{
    const int np = w->np;
    const double *p0 = w->geom + fo->geom_off, *A = p0 + np, *sc = p0 + 2 * np;
    double length = sc[0], AdA = sc[1], BdA = sc[2], r = sc[3];
    int no_end = (fo->flags & NDT_OF_NO_END_TEST) != 0;
    double sA[MAXD], X[MAXD], Y[MAXD], tmp[MAXD];
    int ret = 0;
    double VdA = vdot(v, A, np);
    double OdA = vdot(o, A, np);
    double Vaaa = VdA / AdA;
    double BOaa = (BdA - OdA) / AdA;
    vscale(A, Vaaa, sA, np);
    vsub(v, sA, Y, np);
    vsub(o, p0, tmp, np);
    vscale(A, BOaa, sA, np);
    vadd(tmp, sA, X, np);
    double qa = vdot(Y, Y, np);
    double qb = vdot(Y, X, np);
    qb *= 2;
    double qc = vdot(X, X, np);
    qc -= r * r;
    double det = qb * qb - 4 * qa * qc;
    if (det <= 0) return 0;
    double root = sqrt(det);
    double t1 = (-qb + root) / (2 * qa);
    double t2 = (-qb - root) / (2 * qa);
    for (int pass = 0; pass < 2 && ret == 0; ++pass) {
        double t = pass ? t1 : t2;
        if (!(t > EPS)) continue;
        vscale(v, t, sA, np);
        vadd(o, sA, res, np);
        if (no_end) { ret = 1; break; }
        double bc[MAXD];                              /* between_ends, cylinder.c:85-102 */
        vsub(res, p0, bc, np);
        double s = vdot(bc, A, np);
        if (s > 0 && s < length) ret = 1;
    }
    if (ret) {
        vsub(res, p0, X, np);
        double ncda = vdot(A, X, np);
        vscale(A, ncda / AdA, Y, np);
        vsub(X, Y, nrm, np);
    }
    return ret;
}

static int isect_facet(const view *w, const ndt_flat_object *fo, const double *o, const double *v,
                       double *res, double *nrm)                        /* facet.c:166-269 */
{
    const int np = w->np, n = w->n;
    const double *p = w->geom + fo->geom_off, *basis = p + 3 * np, *fn = p + 5 * np, *sc = p + 6 * np;
    const double *ada = sc, *bda = sc + 2, *ang = sc + 4;
    double P[MAXD], Q[MAXD], sA[MAXD], sum[MAXD];
    vzero(sum, np);
    for (int i = 0; i < 2; ++i) {
        double vda = vdot(v, basis + (size_t)i * np, np);
        vscale(basis + (size_t)i * np, vda / ada[i], sA, np);
        vadd(sum, sA, sum, np);
    }
    vsub(sum, v, P, np);
    vzero(sum, n);
    for (int i = 0; i < 2; ++i) {
        double oda = vdot(o, basis + (size_t)i * np, np);
        vscale(basis + (size_t)i * np, (oda - bda[i]) / ada[i], sA, np);
        vadd(sum, sA, sum, np);
    }
    vsub(p + np, o, Q, np);
    vadd(Q, sum, Q, np);
    double qa = vdot(P, P, np);
    double qb = vdot(P, Q, np);
    qb *= 2;
    double qc = vdot(Q, Q, np);
    double t = -1.0;
    if (fabs(qa) < EPS) {
        if (fabs(qb) < EPS) t = -qc / qb;            /* sic: facet.c:216-222 */
        else t = -1.0;
    } else {
        t = -qb / (2 * qa);
    }
    if (t < EPS) return 0;
    double dist = qa * t * t + qb * t + qc;
    if (fabs(dist) > EPS) return 0;
    vscale(v, t, sA, np);
    vadd(o, sA, res, np);
    int ret = 1;
    for (int i = 0; i < 3; ++i) {                    /* inside_edges, facet.c:149-164 */
        int j = (i + 1) % 3;
        double a = vangle3(res, p + (size_t)i * np, p + (size_t)j * np, np);
        if (a > ang[i]) { ret = 0; break; }
    }
    memcpy(nrm, fn, (size_t)n * sizeof(double));
    return ret;
}

static int isect_hfacet(const view *w, const ndt_flat_object *fo, const double *o, const double *v,
                        double *res, double *nrm)                       /* hfacet.c:211-310 */
{
    const int np = w->np, n = w->n;
    const double *v0 = w->geom + fo->geom_off, *ue0 = v0 + np, *eperp = v0 + 2 * np;
    const double *normals = v0 + 3 * np, *sc = v0 + 6 * np;
    double ones[MAXD];
    for (int i = 0; i < np; ++i) ones[i] = 1.0;
    if (np > n) ones[n] = sc[4];
    double R[MAXD], vE0[MAXD], vE2[MAXD], Q[MAXD], oP0[MAXD];
    vproj_unit(v, ue0, vE0, np);
    vproj_unit(v, eperp, vE2, np);
    vadd(vE0, vE2, R, np);
    vsub(R, v, R, np);
    double Rv = vdot(R, ones, np);
    if (fabs(Rv) < EPS) return 0;
    vsub(o, v0, oP0, np);
    vproj_unit(oP0, ue0, vE0, np);
    vproj_unit(oP0, eperp, vE2, np);
    vadd(vE0, vE2, Q, np);
    vsub(Q, oP0, Q, np);
    double Qv = vdot(Q, ones, np);
    double t = -Qv / Rv;
    double lam[3];
    int ret = 0;
    if (t > EPS) {
        vscale(v, t, res, np);
        vadd(o, res, res, np);
        /* get_barycentric, hfacet.c:147-188 */
        double C[MAXD];
        vsub(res, v0, C, np);
        double xp = vdot(ue0, C, np), yp = vdot(eperp, C, np);
        double x1 = 0, y1 = 0, x2 = sc[0], y2 = sc[1], x3 = sc[2], y3 = sc[3];
        lam[0] = ((y2 - y3) * (xp - x3) + (x3 - x2) * (yp - y3)) /
                 ((y2 - y3) * (x1 - x3) + (x3 - x2) * (y1 - y3));
        lam[1] = ((y3 - y1) * (xp - x3) + (x1 - x3) * (yp - y3)) /
                 ((y2 - y3) * (x1 - x3) + (x3 - x2) * (y1 - y3));
        lam[2] = 1 - lam[0] - lam[1];
        ret = 1;
        for (int i = 0; i < 3; ++i)
            if (lam[i] < -EPS || lam[i] > 1 + EPS) { ret = 0; break; }
    }
    if (ret) {
        if (fo->flags & NDT_OF_USE_NORMALS) {
            vzero(nrm, n);
            for (int i = 0; i < 3; ++i) {
                vscale(normals + (size_t)i * np, lam[i], R, np);
                vadd(nrm, R, nrm, np);
            }
        } else {
            double D[MAXD], U[MAXD], V[MAXD];    /* hfacet_point_in_plane, hfacet.c:120-144 */
            vsub(o, v0, D, np);
            vproj_unit(D, ue0, U, np);
            vproj_unit(D, eperp, V, np);
            vadd(U, V, R, np);
            vadd(R, v0, R, np);
            vsub(o, R, nrm, np);
            vunit(nrm, np);
        }
    }
    return ret;
}

static int trace_list(const view *w, const int32_t *ids, int base, int cnt, unsigned char *mask,
                      const double *pos, const double *look, double *hit, double *hit_n,
                      int *obj_id, double *t_ptr, double dist_limit);

/* object.c:605-630 followed by the plugin's intersect */
static int object_intersect(const view *w, int id, const double *o, const double *v,
                            double *res, double *nrm, int *obj_id, double min_dist)
{
    const ndt_flat_object *fo = &w->obj[id];
    int ret = 0;
    if (fo->bs_radius > 0) {
        if (bsphere_test(w, id, o, v, min_dist) <= 0) { *obj_id = -1; return 0; }
    }
    switch (fo->type) {
    case NDT_T_SPHERE:    ret = isect_sphere(w, fo, o, v, res, nrm); break;
    case NDT_T_HPLANE: {
        const double *g = w->geom + fo->geom_off;
        ret = isect_hplane_raw(g, g + w->np, o, v, res, nrm, w->n, w->np);
        break;
    }
    case NDT_T_HDISK:     ret = isect_hdisk(w, fo, o, v, res, nrm); break;
    case NDT_T_ORTHOTOPE: ret = isect_orthotope(w, fo, o, v, res, nrm); break;
    case NDT_T_FACET:     ret = isect_facet(w, fo, o, v, res, nrm); break;
    case NDT_T_HFACET:    ret = isect_hfacet(w, fo, o, v, res, nrm); break;
    case NDT_T_CYLINDER:  ret = isect_cylinder(w, fo, o, v, res, nrm); break;
    case NDT_T_HCYLINDER: ret = isect_hcylinder(w, fo, o, v, res, nrm); break;
    case NDT_T_HCUBE: {                               /* hcube.c:236-250 */
        int sub = -1;
        ret = trace_list(w, NULL, fo->child_begin, fo->child_count, NULL, o, v, res, nrm, &sub, NULL, -1.0);
        break;
    }
    default: ret = 0;
    }
    if (ret) *obj_id = fo->report_id;
    return ret;
}

/* object.c:692-747 */
static int trace_list(const view *w, const int32_t *ids, int base, int cnt, unsigned char *mask,
                      const double *pos, const double *look, double *hit, double *hit_n,
                      int *obj_id, double *t_ptr, double dist_limit)
{
    const int n = w->n, np = w->np;
    double min_dist = -1;
    double res[MAXD], nrm[MAXD];
    vzero(res, np);
    vzero(nrm, np);
    if (obj_id) *obj_id = -1;
    for (int i = 0; i < cnt; ++i) {
        int id = ids ? ids[i] : base + i;
        if (mask && ids) {
            if (mask[id]) continue;
            mask[id] = 1;
        }
        int tmp = -1;
        int ret = object_intersect(w, id, pos, look, res, nrm, &tmp, min_dist);
        if (ret > 0) {
            double dist = vdist(pos, res, np);
            if (dist > EPS && (dist + EPS < min_dist || min_dist < 0)) {
                min_dist = dist;
                memcpy(hit, res, (size_t)n * sizeof(double));
                memcpy(hit_n, nrm, (size_t)n * sizeof(double));
                if (obj_id) *obj_id = tmp;
            }
            if (dist_limit == 0.0 || dist < dist_limit) break;
        }
    }
    if (t_ptr != NULL && min_dist > EPS) *t_ptr = min_dist;
    if (min_dist < 0) return 0;
    return 1;
}

/* kd-tree.c:84-127 */
static int aabb_hit(const view *w, const double *o, const double *v, double *tl_out, double *tu_out)
{
    const double *lo = w->aabb, *hi = w->aabb + w->np;
    double tl = -DBL_MAX, tu = DBL_MAX;
    for (int i = 0; i < w->n; ++i) {
        if (fabs(v[i]) < EPS2) continue;
        double a = (lo[i] - o[i]) / v[i];
        double b = (hi[i] - o[i]) / v[i];
        if (a > b) { double t = a; a = b; b = t; }
        if (a > tl) tl = a;
        if (b < tu) tu = b;
        if (tu < -EPS) return 0;
    }
    tl -= EPS;
    tu += EPS;
    *tl_out = tl; *tu_out = tu;
    return (tu >= -EPS) && (tl <= tu);
}

typedef struct {
    const view *w;
    const double *o, *v;
    double v_inv[MAXD];
    double *hit, *hit_n;
    unsigned char *mask;
    int *obj_id;
    double *t_ptr;
    double dist_limit;
} kdq;

/* kd-tree.c:482-568 */
static int kd_visit(kdq *q, int ni, double tl, double tu)
{
    if (ni < 0) return 0;
    if (tu < 0.0) return 0;
    const view *w = q->w;
    const ndt_flat_node *nd = &w->nodes[ni];
    int ret = 0;
    if (nd->leaf_count > 0) {
        double t, lhit[MAXD], lnrm[MAXD];
        int oid = -1;
        vzero(lhit, w->np); vzero(lnrm, w->np);
        ret = trace_list(w, w->leaf + nd->leaf_begin, 0, nd->leaf_count, q->mask, q->o, q->v,
                         lhit, lnrm, &oid, &t, q->dist_limit);
        if (ret && t < *q->t_ptr) {
            *q->t_ptr = t;
            *q->obj_id = oid;
            memcpy(q->hit, lhit, (size_t)w->n * sizeof(double));
            memcpy(q->hit_n, lnrm, (size_t)w->n * sizeof(double));
        }
        if (nd->dim < 0) return ret;
    }
    if (nd->dim < 0) return 0;   /* empty leaf: both children are NULL in the reference */
    int nr = nd->left, fr = nd->right;
    double vi = q->v_inv[nd->dim], oi = q->o[nd->dim], b = nd->boundary;
    if (vi < EPS2) { int t = nr; nr = fr; fr = t; }
    if (-INV_EPS2 <= vi && vi <= INV_EPS2) {
        double tp = (b - oi) * vi;
        if (tu < tp - EPS && *q->t_ptr > tl) {
            ret |= kd_visit(q, nr, tl, tu);
        } else if (tl > tp + EPS && *q->t_ptr > tl) {
            ret |= kd_visit(q, fr, tl, tu);
        } else {
            if (*q->t_ptr > tl) ret |= kd_visit(q, nr, tl, tp + EPS);
            if (*q->t_ptr > tp) ret |= kd_visit(q, fr, tp - EPS, tu);
        }
    } else {
        if (oi < b + EPS && *q->t_ptr > tl) ret |= kd_visit(q, nr, tl, tu);
        if (oi > b - EPS && *q->t_ptr > tl) ret |= kd_visit(q, fr, tl, tu);
    }
    return ret;
}

/* kd-tree.c:570-625 (trace_kd, object.c:683) */
static int trace_kd(const view *w, unsigned char *mask, const double *o, const double *v,
                    double *hit, double *hit_n, int *obj_id, double dist_limit)
{
    kdq q;
    q.w = w; q.o = o; q.v = v;
    for (int i = 0; i < w->n; ++i) {
        double vi = v[i];
        if (vi < EPS2 && vi >= 0.0) q.v_inv[i] = INV_EPS2;
        else if (vi > -EPS2 && vi <= 0.0) q.v_inv[i] = -INV_EPS2;
        else q.v_inv[i] = 1.0 / vi;
    }
    double t = DBL_MAX;
    int ret = trace_list(w, w->inf, 0, w->h->n_inf, NULL, o, v, hit, hit_n, obj_id, &t, dist_limit);
    double tl, tu;
    if (aabb_hit(w, o, v, &tl, &tu)) {
        double lt = DBL_MAX, lhit[MAXD], lnrm[MAXD];
        int lid = -1;
        memset(mask, 0, (size_t)w->h->n_items);
        vzero(lhit, w->np); vzero(lnrm, w->np);
        q.hit = lhit; q.hit_n = lnrm; q.mask = mask; q.obj_id = &lid; q.t_ptr = &lt; q.dist_limit = dist_limit;
        int lret = w->h->n_nodes > 0 ? kd_visit(&q, 0, tl, tu) : 0;
        if (lret) {
            if (!ret || (lt > EPS && lt + EPS < t)) {
                memcpy(hit, lhit, (size_t)w->n * sizeof(double));
                memcpy(hit_n, lnrm, (size_t)w->n * sizeof(double));
                *obj_id = lid;
                ret |= lret;
            }
        }
    }
    return ret;
}

/* ---- shading: ndt.c:71-326 --------------------------------------------------- */
static void apply_lights(const view *w, unsigned char *mask, raycnt *rc, int oid, const double *src,
                         const double *look, const double *hit, const double *hit_n, double *clr)
{
    const int np = w->np, n = w->n;
    const ndt_flat_object *fo = &w->obj[oid];
    double hr = fo->rgb[0], hg = fo->rgb[1], hb = fo->rgb[2];
    double rr = 0.0, rg = 0.0, rb = 0.0;
    if (w->h->specular) { rr = fo->refl[0]; rg = fo->refl[1]; rb = fo->refl[2]; }
    clr[0] = hr * w->h->ambient[0];
    clr[1] = hg * w->h->ambient[1];
    clr[2] = hb * w->h->ambient[2];
    clr[3] = 1.0;
    double rev_view[MAXD], rev_light[MAXD], light_vec[MAXD], light_hit[MAXD], light_hit_n[MAXD];
    double lgt_pos[MAXD], near_pos[MAXD];
    vzero(rev_view, np); vzero(rev_light, np); vzero(light_vec, np); vzero(light_hit, np);
    vzero(light_hit_n, np); vzero(lgt_pos, np); vzero(near_pos, np);
    for (int i = 0; i < w->h->n_lights; ++i) {
        const ndt_flat_light *L = &w->lights[i];
        const double *lv = w->geom + L->vec_off;
        const double *lpos = lv, *ldir = lv + np, *lrev = lv + 2 * np, *lnear = lv + 3 * np;
        if (L->type == NDT_L_AMBIENT) {
            clr[0] += hr * L->rgb[0];
            clr[1] += hg * L->rgb[1];
            clr[2] += hb * L->rgb[2];
            continue;
        }
        memcpy(lgt_pos, lpos, (size_t)n * sizeof(double));
        if (L->type == NDT_L_POINT || L->type == NDT_L_SPOT) {
            vsub(lgt_pos, hit, rev_light, np);
            vunit(rev_light, np);
        } else {
            vcopy(rev_light, lrev, np);          /* unit(-dir), hoisted: ndt.c:156-158 */
        }
        vsub(src, hit, rev_view, np);
        double d1 = vdot(rev_light, hit_n, np);
        double d2 = vdot(rev_view, hit_n, np);
        if ((d1 * d2) <= 0) continue;

        int light_obj = -1, got;
        double ldist2 = 1.0;
        if (L->type == NDT_L_POINT || L->type == NDT_L_SPOT) {
            double dist_limit = vdist(hit, lgt_pos, np);
            dist_limit += EPS;
            vsub(hit, lgt_pos, light_vec, np);
            ldist2 = vdot(light_vec, light_vec, np);
            vunit(light_vec, np);
            if (L->type == NDT_L_SPOT) {
                double a = vangle(ldir, light_vec, np);
                if ((a * 180.0 / M_PI) > L->angle) continue;
            }
            rc->shadow++;
            got = trace_kd(w, mask, lgt_pos, light_vec, light_hit, light_hit_n, &light_obj, dist_limit);
            if (!got || light_obj != oid) continue;
            double dist = vdist(hit, light_hit, np);
            if (dist > EPS) continue;
        } else {
            vadd(lnear, hit, near_pos, np);      /* ndt.c:234-237 */
            vscale(ldir, -1.0, light_vec, np);
            rc->shadow++;
            got = trace_kd(w, mask, near_pos, rev_light, light_hit, light_hit_n, &light_obj, 0.0);
            if (got) continue;
            memcpy(light_vec, ldir, (size_t)n * sizeof(double));
            memcpy(light_hit, hit, (size_t)n * sizeof(double));
            memcpy(light_hit_n, hit_n, (size_t)n * sizeof(double));
            light_obj = oid;
            ldist2 = 1;
        }
        double angle = vangle(hit_n, light_vec, np);
        if (angle > M_PI / 2.0) angle = M_PI - angle;
        double light_scale = cos(angle) / ldist2;
        if (!(fo->flags & NDT_OF_TRANSPARENT)) {
            clr[0] += hr * L->rgb[0] * light_scale;
            clr[1] += hg * L->rgb[1] * light_scale;
            clr[2] += hb * L->rgb[2] * light_scale;
        }
        if (w->h->specular) {                     /* ndt.c:277-310 */
            double lref[MAXD], rev_look[MAXD];
            vreflect(light_vec, light_hit_n, lref, 0.5, np);
            vunit(lref, np);
            vscale(look, -1, rev_look, np);
            vunit(rev_look, np);
            double rv = vdot(lref, rev_look, np);
            rv = REF_MAX(0, rv);
            double rvn = pow(rv, 50);
            double ml = L->max_rgb;
            clr[0] += rr * L->rgb[0] / ml * rvn;
            clr[1] += rg * L->rgb[1] / ml * rvn;
            clr[2] += rb * L->rgb[2] / ml * rvn;
            clr[3] = 1.0;
        }
    }
}

/* ndt.c:329-450.  Returns through px[4]; prim_* describe this very ray. */
static void ray_color(const view *w, unsigned char *mask, raycnt *rc, const double *src, const double *look,
                      double *px, double frac, int max_depth, int is_primary,
                      int *prim_hit, int *prim_id, double *prim_dist)
{
    const int np = w->np;
    px[0] = px[1] = px[2] = 0.0; px[3] = 1.0;
    if (frac < (1.0 / 512.0)) return;
    if (max_depth <= 0) return;
    double hit[MAXD], hit_n[MAXD], clr[4] = {0, 0, 0, 0};
    vzero(hit, np); vzero(hit_n, np);
    int oid = -1;
    if (is_primary) rc->primary++; else rc->bounce++;
    trace_kd(w, mask, src, look, hit, hit_n, &oid, -1.0);
    double trace_dist = -1;
    if (oid >= 0 || is_primary) trace_dist = vdist(hit, src, np);
    if (is_primary) {
        *prim_id = oid;
        *prim_hit = (oid >= 0 && trace_dist > EPS);
        *prim_dist = (oid >= 0) ? trace_dist : -1.0;
    }
    if (oid >= 0 && trace_dist > EPS) {
        const ndt_flat_object *fo = &w->obj[oid];
        apply_lights(w, mask, rc, oid, src, look, hit, hit_n, clr);
        double hr = fo->refl[0], hg = fo->refl[1], hb = fo->refl[2];
        double ref[4], nr[MAXD];
        vzero(nr, np);
        double contrib = REF_MAX(hr, REF_MAX(hg, hb));
        if (contrib > 0) {
            if (hr != 0.0 || hg != 0.0 || hb != 0.0) {
                vreflect(look, hit_n, nr, 1.0, np);
                vunit(nr, np);
                ray_color(w, mask, rc, hit, nr, ref, contrib * frac, max_depth - 1, 0, NULL, NULL, NULL);
                if (w->h->specular) {
                    clr[0] = (1 - hr) * (clr[0]) + (hr) * ref[0];
                    clr[1] = (1 - hg) * (clr[1]) + (hg) * ref[1];
                    clr[2] = (1 - hb) * (clr[2]) + (hb) * ref[2];
                    clr[3] = 1.0;
                } else {
                    clr[0] += hr * ref[0];
                    clr[1] += hg * ref[1];
                    clr[2] += hb * ref[2];
                    clr[3] = 1.0;
                }
            }
        }
        if (fo->flags & NDT_OF_TRANSPARENT) {
            vrefract(look, hit_n, nr, fo->refract_index, np);
            vunit(nr, np);
            ray_color(w, mask, rc, hit, nr, ref, (1 - contrib) * frac, max_depth - 1, 0, NULL, NULL, NULL);
            clr[0] += (1.0 - hr) * ref[0];
            clr[1] += (1.0 - hg) * ref[1];
            clr[2] += (1.0 - hb) * ref[2];
            clr[3] = 1.0;
        }
    } else {
        clr[0] = w->h->bg[0]; clr[1] = w->h->bg[1]; clr[2] = w->h->bg[2]; clr[3] = w->h->bg[3];
    }
    memcpy(px, clr, sizeof clr);
}

/* the sample loop of get_pixel_color (ndt.c:488-568) for samples==1: every
 * iteration re-traces the SAME ray, so the colour is traced once and only the
 * scalar accumulate / convergence arithmetic is replayed */
static int replay_samples(const double *l, double *out)
{
    const int min_samples = 1, max_samples = 10000;
    const double max_diff = 1.0 / 256.0;
    double clr_diff = 256;
    double t[4] = {0, 0, 0, 0};
    int ts = 0, i;
    for (i = 0; i < min_samples || (i < max_samples && clr_diff > max_diff); ++i) {
        if (i > 1) {
            clr_diff = REF_MAX(fabs(t[0] / (i - 1) - (t[0] + l[0]) / i),
                       REF_MAX(fabs(t[1] / (i - 1) - (t[1] + l[1]) / i),
                               fabs(t[2] / (i - 1) - (t[2] + l[2]) / i)));
        }
        t[0] += l[0]; t[1] += l[1]; t[2] += l[2]; t[3] += l[3];
        ts += 1;
    }
    for (int k = 0; k < 4; ++k) out[k] = t[k] / ts;
    return ts;
}

/* image.h:36-39 */
static unsigned char d2c(double d)
{
    double c = REF_MAX(0.0, REF_MIN(1.0, d));
    double s = sqrt(c) * 255;
    if (!(s == s)) return 0;
    return (unsigned char)s;
}

typedef struct {
    const void *blob;
    int x0, y0, tw, th, row0, rows_step;
    double *rgba; uint8_t *u8; uint8_t *hit; int32_t *id; double *inv_depth;
    int eye_override;   /* 0, or 1 / 2: the left / right pass of ANAGLYPH_3D */
    uint64_t st[6]; /* primary, bounce, shadow, rays_ref, samples, reserved */
} job;

static void *render_rows(void *arg)
{
    job *J = arg;
    view W, *w = &W;
    view_init(w, J->blob);
    const int np = w->np;
    const double *cpos = w->cam, *corig = w->cam + np, *cdx = w->cam + 2 * np, *cdy = w->cam + 3 * np;
    unsigned char *mask = malloc((size_t)(w->h->n_items ? w->h->n_items : 1));
    const int W_ = w->h->width, H_ = w->h->height;
    for (int ty = J->row0; ty < J->th; ty += J->rows_step) {
        int j = J->y0 + ty;
        for (int tx = 0; tx < J->tw; ++tx) {
            int i = J->x0 + tx;
            double pixel[MAXD], tmp[MAXD], look[MAXD], eye[MAXD];
            int blank = 0;
            if (!w->h->off_view) {
                double x = (double)i / (double)W_ - 0.5;          /* ndt.c:632 */
                double y = -((double)j / (double)H_ - 0.5);       /* ndt.c:633 */
                /* camera_target_point, camera.c:557-575 */
                vcopy(pixel, corig, np);
                vscale(cdx, x, tmp, np); vadd(pixel, tmp, pixel, np);
                vscale(cdy, y, tmp, np); vadd(pixel, tmp, pixel, np);
                if (w->h->use_focal) {
                    vsub(pixel, cpos, tmp, np);
                    vscale(tmp, w->h->focal_scale, tmp, np);
                    vadd(cpos, tmp, pixel, np);
                }
                vcopy(eye, cpos, np);
            } else {
                /* render_pixel (ndt.c:578-653) + camera_target_point (camera.c:504-581) + the eye of
                 * get_pixel_color (ndt.c:488-525), from the per-column / per-row tables the flattener
                 * filled with the host's libm (ndt_flat.h, version 5) */
                const double *ext = NDT_FLAT_PTR(J->blob, const double, w->h->off_view);
                const double *col = ext + 5 * np + (size_t)i * 4;
                const double *row = ext + 5 * np + (size_t)W_ * 4 + (size_t)j * 6;
                const double x = col[0], y = row[0];
                blank = row[4] != 0.0;
                if (w->h->cam_type == NDT_CAM_NORMAL) {
                    vcopy(pixel, corig, np);
                    vscale(cdx, x, tmp, np); vadd(pixel, tmp, pixel, np);
                    vscale(cdy, y, tmp, np); vadd(pixel, tmp, pixel, np);
                    if (w->h->use_focal) {
                        vsub(pixel, cpos, tmp, np);
                        vscale(tmp, w->h->focal_scale, tmp, np);
                        vadd(cpos, tmp, pixel, np);
                    }
                } else {
                    const double dist = w->h->cam_dist;
                    double vx, vy, vz;
                    if (w->h->cam_type == NDT_CAM_VR) {           /* camera.c:507-529 */
                        vx = dist * col[1] * row[2];
                        vy = dist * row[1];
                        vz = dist * col[2] * row[2];
                    } else {                                      /* camera.c:530-556 */
                        vx = dist * col[1];
                        vy = row[3];
                        vz = dist * col[2];
                    }
                    vcopy(pixel, cpos, np);
                    vscale(ext + 2 * np, vx, tmp, np); vadd(pixel, tmp, pixel, np);
                    vscale(ext + 3 * np, vy, tmp, np); vadd(pixel, tmp, pixel, np);
                    vscale(ext + 4 * np, vz, tmp, np); vadd(pixel, tmp, pixel, np);
                }
                const int e = J->eye_override ? J->eye_override : ((int)col[3] | (int)row[5]);
                if (e == 0) vcopy(eye, cpos, np);
                else if (w->h->view_eyes)
                    vcopy(eye, ext + 5 * np + (size_t)W_ * 4 + (size_t)H_ * 6 + ((size_t)i * 2 + (size_t)(e - 1)) * np, np);
                else vcopy(eye, ext + (size_t)(e - 1) * np, np);
            }
            size_t p = (size_t)ty * J->tw + tx;
            if (blank) {                                      /* ndt.c:619-626: black, nothing traced */
                if (J->rgba) memset(J->rgba + 4 * p, 0, 32);
                if (J->u8) memset(J->u8 + 4 * p, 0, 4);
                if (J->hit) J->hit[p] = 0;
                if (J->id) J->id[p] = -1;
                if (J->inv_depth) J->inv_depth[p] = 0.0;
                continue;
            }
            vsub(pixel, eye, look, np);
            vunit(look, np);
            raycnt rc = {0, 0, 0};
            double l[4], out[4];
            int ph = 0, pid = -1; double pd = -1;
            ray_color(w, mask, &rc, eye, look, l, 1.0, w->h->max_optic_depth, 1, &ph, &pid, &pd);
            int ns = replay_samples(l, out);
            if (J->rgba) memcpy(J->rgba + 4 * p, out, sizeof out);
            if (J->u8) for (int k = 0; k < 4; ++k) J->u8[4 * p + k] = d2c(out[k]);
            if (J->hit) J->hit[p] = (uint8_t)ph;
            if (J->id) J->id[p] = pid;
            if (J->inv_depth) J->inv_depth[p] = (pid >= 0 && pd > EPS) ? 1.0 / pd : 0.0;
            uint64_t tree = rc.primary + rc.bounce + rc.shadow;
            J->st[0] += rc.primary; J->st[1] += rc.bounce; J->st[2] += rc.shadow;
            J->st[3] += tree * (uint64_t)ns; J->st[4] += (uint64_t)ns;
        }
    }
    free(mask);
    return NULL;
}

/* stats[5]: rays_primary, rays_bounce, rays_shadow, rays_ref, samples */
static int render_pass(const void *blob, int x0, int y0, int tw, int th, int threads, int eye_override,
                       double *rgba, uint8_t *u8, uint8_t *hit, int32_t *id, double *inv_depth,
                       uint64_t *stats)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    job *jobs = calloc((size_t)threads, sizeof *jobs);
    pthread_t *thr = calloc((size_t)threads, sizeof *thr);
    for (int t = 0; t < threads; ++t) {
        job *J = &jobs[t];
        J->blob = blob; J->x0 = x0; J->y0 = y0; J->tw = tw; J->th = th;
        J->row0 = t; J->rows_step = threads;
        J->rgba = rgba; J->u8 = u8; J->hit = hit; J->id = id; J->inv_depth = inv_depth;
        J->eye_override = eye_override;
        if (threads > 1) pthread_create(&thr[t], NULL, render_rows, J);
        else render_rows(J);
    }
    for (int t = 0; t < threads; ++t) {
        if (threads > 1) pthread_join(thr[t], NULL);
        if (stats) for (int k = 0; k < 5; ++k) stats[k] += jobs[t].st[k];
    }
    free(jobs); free(thr);
    return 0;
}

/* stats[5]: rays_primary, rays_bounce, rays_shadow, rays_ref, samples */
int ndo_render(const void *blob, int x0, int y0, int tw, int th, int threads,
               double *rgba, uint8_t *u8, uint8_t *hit, int32_t *id, double *inv_depth,
               uint64_t *stats)
{
    const ndt_flat_header *h = blob;
    if (!h || h->magic != NDT_FLAT_MAGIC || h->version != NDT_FLAT_VERSION) return -1;
    if (stats) memset(stats, 0, 5 * sizeof *stats);
    if (h->stereo_mode != NDT_ANAGLYPH_3D)
        return render_pass(blob, x0, y0, tw, th, threads, 0, rgba, u8, hit, id, inv_depth, stats);
    /* ANAGLYPH_3D, ndt.c:634-646: one render per eye, mixed into red (left) and blue (right);
     * depth (and our hit / id buffers) from the left eye */
    const size_t px = (size_t)tw * th;
    double *l = malloc(px * 32), *r = malloc(px * 32);
    if (!l || !r) { free(l); free(r); return -2; }
    render_pass(blob, x0, y0, tw, th, threads, 1, l, NULL, hit, id, inv_depth, stats);
    render_pass(blob, x0, y0, tw, th, threads, 2, r, NULL, NULL, NULL, NULL, stats);
    for (size_t p = 0; p < px; ++p) {
        double out[4];
        out[0] = 0.299 * l[4 * p] + 0.587 * l[4 * p + 1] + 0.114 * l[4 * p + 2];
        out[1] = 0;
        out[2] = 0.299 * r[4 * p] + 0.587 * r[4 * p + 1] + 0.114 * r[4 * p + 2];
        out[3] = 1.0;
        if (rgba) memcpy(rgba + 4 * p, out, sizeof out);
        if (u8) for (int k = 0; k < 4; ++k) u8[4 * p + k] = d2c(out[k]);
    }
    free(l); free(r);
    return 0;
}

/* ---- recursive anti-aliasing: resample_pixel / recursive_resample, ndt.c:655-733 ------------------- */
typedef struct { double c[4]; } px4;

static void avg4(const px4 *p1, const px4 *p2, const px4 *p3, const px4 *p4, px4 *avg, double *var)
{   /* image_avg_dbl_pixels4, image.c:1175-1197 */
    for (int k = 0; k < 4; ++k) avg->c[k] = (p1->c[k] + p2->c[k] + p3->c[k] + p4->c[k]) / 4;
    if (var) {
        double v = 0;
        for (int k = 0; k < 4; ++k)
            v += fabs(avg->c[k] - p1->c[k]) + fabs(avg->c[k] - p2->c[k]) +
                 fabs(avg->c[k] - p3->c[k]) + fabs(avg->c[k] - p4->c[k]);
        *var = v;
    }
}

typedef struct {
    view *w; unsigned char *mask; raycnt rc; uint64_t rays_ref, samples;
    int aa_diff, aa_depth;
} aa_ctx;

/* render_pixel at a fractional pixel position (MONO, CAMERA_NORMAL): ndt.c:628-650 */
static void aa_sample(aa_ctx *A, double ip, double jp, px4 *out)
{
    view *w = A->w;
    const int np = w->np;
    const double *cpos = w->cam, *corig = w->cam + np, *cdx = w->cam + 2 * np, *cdy = w->cam + 3 * np;
    double x = ip / (double)w->h->width - 0.5;
    double y = -(jp / (double)w->h->height - 0.5);
    double pixel[MAXD], tmp[MAXD], look[MAXD];
    vcopy(pixel, corig, np);
    vscale(cdx, x, tmp, np); vadd(pixel, tmp, pixel, np);
    vscale(cdy, y, tmp, np); vadd(pixel, tmp, pixel, np);
    if (w->h->use_focal) {
        vsub(pixel, cpos, tmp, np);
        vscale(tmp, w->h->focal_scale, tmp, np);
        vadd(cpos, tmp, pixel, np);
    }
    vsub(pixel, cpos, look, np);
    vunit(look, np);
    raycnt rc = {0, 0, 0};
    double l[4];
    int ph = 0, pid = -1; double pd = -1;
    ray_color(w, A->mask, &rc, cpos, look, l, 1.0, w->h->max_optic_depth, 1, &ph, &pid, &pd);
    int ns = replay_samples(l, out->c);
    A->rc.primary += rc.primary; A->rc.bounce += rc.bounce; A->rc.shadow += rc.shadow;
    A->rays_ref += (rc.primary + rc.bounce + rc.shadow) * (uint64_t)ns; A->samples += (uint64_t)ns;
}

static void aa_recurse(aa_ctx *A, double x, double y, double step,
                       const px4 *p1, const px4 *p2, const px4 *p3, const px4 *p4, px4 *res)
{
    if (A->aa_depth <= 0 || step < 1.0 / (2 << (A->aa_depth - 1))) {      /* ndt.c:663-666 */
        avg4(p1, p2, p3, p4, res, NULL);
        return;
    }
    double hs = step / 2;
    px4 p5, p6, p7, p8, p9;
    aa_sample(A, x + hs, y + hs, &p5);
    aa_sample(A, x + hs, y, &p6);
    aa_sample(A, x, y + hs, &p7);
    aa_sample(A, x + step, y + hs, &p8);
    aa_sample(A, x + hs, y + step, &p9);
    px4 sp1, sp2, sp3, sp4;
    double var1 = 0, var2 = 0, var3 = 0, var4 = 0;
    double threshold = A->aa_diff / 255.0;
    avg4(p1, &p6, &p7, &p5, &sp1, &var1);
    if (var1 > threshold) aa_recurse(A, x, y, hs, p1, &p6, &p7, &p5, &sp1);
    avg4(p2, &p6, &p8, &p5, &sp2, &var2);
    if (var2 > threshold) aa_recurse(A, x + hs, y, hs, &p6, p2, &p5, &p8, &sp2);
    avg4(p3, &p9, &p7, &p5, &sp3, &var3);
    if (var3 > threshold) aa_recurse(A, x, y + hs, hs, &p7, &p5, p3, &p9, &sp3);
    avg4(p4, &p9, &p8, &p5, &sp4, &var4);
    if (var4 > threshold) aa_recurse(A, x + hs, y + hs, hs, &p5, &p8, &p9, p4, &sp4);
    avg4(&sp1, &sp2, &sp3, &sp4, res, NULL);
}

typedef struct {
    const void *blob; const double *img; int W, H, row0, rows_step, aa_diff, aa_depth;
    uint8_t *u8; double *f64; uint64_t st[6];
} aa_job;

static void *aa_rows(void *arg)
{
    aa_job *J = arg;
    view Wv;
    view_init(&Wv, J->blob);
    aa_ctx A;
    memset(&A, 0, sizeof A);
    A.w = &Wv; A.aa_diff = J->aa_diff; A.aa_depth = J->aa_depth;
    A.mask = malloc((size_t)(Wv.h->n_items ? Wv.h->n_items : 1));
    const int W = J->W, H = J->H;
    for (int j = J->row0; j < H; j += J->rows_step)
        for (int i = 0; i < W; ++i) {
            const px4 *p1 = (const px4 *)(J->img + 4 * ((size_t)(W + 1) * j + i)), *p2 = p1 + 1;
            const px4 *p3 = (const px4 *)(J->img + 4 * ((size_t)(W + 1) * (j + 1) + i)), *p4 = p3 + 1;
            px4 clr;
            if (J->aa_depth >= 0 && J->aa_diff < 256) {
                double var = 0.0;
                avg4(p1, p2, p3, p4, &clr, &var);                 /* resample_pixel, ndt.c:709-733 */
                if (var > J->aa_diff / 255.0) {
                    J->st[5]++;
                    aa_recurse(&A, i, j, 1.0, p1, p2, p3, p4, &clr);
                }
            } else {
                clr = *p1;                                         /* ndt.c:1089-1100 */
            }
            size_t p = (size_t)j * W + i;
            if (J->f64) memcpy(J->f64 + 4 * p, clr.c, 32);
            if (J->u8) for (int k = 0; k < 4; ++k) J->u8[4 * p + k] = d2c(clr.c[k]);
        }
    J->st[0] = A.rc.primary; J->st[1] = A.rc.bounce; J->st[2] = A.rc.shadow; J->st[3] = A.rays_ref; J->st[4] = A.samples;
    free(A.mask);
    return NULL;
}

/* render_image with recursive_aa set (ndt.c:921-926, 1039-1100) over a scene from ndt_b200_flatten_aa.
 * u8: W x H RGBA as stored in the 8-bit actual_img; f64: the colours before quantisation.
 * stats[6]: rays_primary, rays_bounce, rays_shadow, rays_ref, samples, pixels resampled */
int ndo_render_aa(const void *blob, int threads, int aa_diff, int aa_depth, uint8_t *u8, double *f64, uint64_t *stats)
{
    const ndt_flat_header *h = blob;
    if (!h || h->magic != NDT_FLAT_MAGIC || h->version != NDT_FLAT_VERSION || !h->aa_pad) return -1;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    const int W = h->width - 1, H = h->height - 1;
    double *img = malloc((size_t)(W + 1) * (H + 1) * 32);
    if (!img) return -2;
    uint64_t st0[5] = {0};
    render_pass(blob, 0, 0, W + 1, H + 1, threads, 0, img, NULL, NULL, NULL, NULL, st0);
    aa_job *jobs = calloc((size_t)threads, sizeof *jobs);
    pthread_t *thr = calloc((size_t)threads, sizeof *thr);
    for (int t = 0; t < threads; ++t) {
        aa_job *J = &jobs[t];
        J->blob = blob; J->img = img; J->W = W; J->H = H; J->row0 = t; J->rows_step = threads;
        J->aa_diff = aa_diff; J->aa_depth = aa_depth; J->u8 = u8; J->f64 = f64;
        if (threads > 1) pthread_create(&thr[t], NULL, aa_rows, J);
        else aa_rows(J);
    }
    if (stats) { memset(stats, 0, 6 * sizeof *stats); for (int k = 0; k < 5; ++k) stats[k] = st0[k]; }
    for (int t = 0; t < threads; ++t) {
        if (threads > 1) pthread_join(thr[t], NULL);
        if (stats) for (int k = 0; k < 6; ++k) stats[k] += jobs[t].st[k];
    }
    free(jobs); free(thr); free(img);
    return 0;
}

/* one nearest-hit query, for the per-primitive known-answer tests */
int ndo_trace(const void *blob, const double *o_in, const double *v_in, double dist_limit,
              double *hit_out, double *normal_out, int *obj_id)
{
    view W, *w = &W;
    view_init(w, blob);
    double o[MAXD], v[MAXD], hit[MAXD], nrm[MAXD];
    vzero(o, MAXD); vzero(v, MAXD); vzero(hit, MAXD); vzero(nrm, MAXD);
    memcpy(o, o_in, (size_t)w->n * sizeof(double));
    memcpy(v, v_in, (size_t)w->n * sizeof(double));
    unsigned char *mask = malloc((size_t)(w->h->n_items ? w->h->n_items : 1));
    int id = -1;
    int r = trace_kd(w, mask, o, v, hit, nrm, &id, dist_limit);
    free(mask);
    memcpy(hit_out, hit, (size_t)w->n * sizeof(double));
    memcpy(normal_out, nrm, (size_t)w->n * sizeof(double));
    *obj_id = id;
    return r;
}
