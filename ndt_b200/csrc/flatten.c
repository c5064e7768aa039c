/*
 * flatten.c -- host side of the drop-in: ndt host structures -> flat scene.
 *
 * Runs where render_image (ndt.c:900) used to start.  It walks the host's
 * `scene` and the global `kd_tree_t` through the struct ABI (ndt_abi.h),
 * forces the state the reference computes lazily on the first ray
 * (bounding spheres: object.c:608-615; hcube faces: hcube.c:155-170) and
 * recomputes each plugin's file-local `prepped` data with the same
 * expressions, in the same order, as the plugin's prepare() -- IEEE-754
 * makes that bit-identical as long as nothing here is fused or re-associated,
 * hence: no -march, -ffp-contract=off, and the two-lane dot product below.
 *
 * Plain C, no CUDA: this file is also linked into the CPU-only test build.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>
#include <unistd.h>
#include <string.h>
#include "ndt_abi.h"
#include "ndt_b200.h"
#include "ndt_internal.h"

#define EPS NDT_EPS

/* ---- vector helpers mirroring vectNd.h on npad-wide arrays ---------------- */

/* vectNd.h:215-227: lane pairs, even and odd sums kept apart, added last */
static double v_dot(const double *a, const double *b, int np)
{
    double s0 = a[0] * b[0], s1 = a[1] * b[1];
    for (int i = 2; i < np; i += 2) {
        s0 = s0 + a[i] * b[i];
        s1 = s1 + a[i + 1] * b[i + 1];
    }
    return s0 + s1;
}
static void v_sub(const double *a, const double *b, double *r, int np)
{
    for (int i = 0; i < np; ++i) r[i] = a[i] - b[i];
}
static void v_scale(const double *a, double s, double *r, int np)
{
    for (int i = 0; i < np; ++i) r[i] = a[i] * s;
}
/* vectNd_copy (vectNd.h:340) moves n lanes only; the pad lane of dst stays */
static void v_copy_n(double *dst, const double *src, int n)
{
    memcpy(dst, src, (size_t)n * sizeof(double));
}
static double v_norm(const double *a, int np) { return sqrt(v_dot(a, a, np)); }
/* vectNd.h:323-329 */
static void v_unitize(double *a, int np)
{
    double len = v_norm(a, np);
    if (len > EPS || len < -EPS)
        v_scale(a, 1.0 / len, a, np);
}
/* vectNd.h:331-338 */
static double v_dist(const double *a, const double *b, int np)
{
    double d[NDT_MAX_DIM];
    v_sub(a, b, d, np);
    return v_norm(d, np);
}
/* vectNd.c:64-81 */
static double v_angle(const double *a, const double *b, int np)
{
    double dp = v_dot(a, b, np);
    double l1 = v_norm(a, np), l2 = v_norm(b, np);
    double div = l1 * l2;
    if (fabs(div) > EPS)
        return acos(dp / div);
    return -1;
}
/* image.h:33 */
#define REF_MAX(x, y) (((x) > (y)) ? (x) : (y))

/* ---- growable pools -------------------------------------------------------- */

typedef struct { double *p; size_t n, cap; } dpool;
typedef struct { int32_t *p; size_t n, cap; } ipool;

static int dpool_need(dpool *d, size_t extra)
{
    if (d->n + extra <= d->cap) return 0;
    size_t nc = d->cap ? d->cap * 2 : 1024;
    while (nc < d->n + extra) nc *= 2;
    double *t = realloc(d->p, nc * sizeof(double));
    if (!t) return -1;
    d->p = t; d->cap = nc;
    return 0;
}
static double *dpool_take(dpool *d, size_t cnt)
{
    if (dpool_need(d, cnt)) return NULL;
    double *r = d->p + d->n;
    memset(r, 0, cnt * sizeof(double));
    d->n += cnt;
    return r;
}
static int ipool_push(ipool *d, int32_t v)
{
    if (d->n == d->cap) {
        size_t nc = d->cap ? d->cap * 2 : 1024;
        int32_t *t = realloc(d->p, nc * sizeof(int32_t));
        if (!t) return -1;
        d->p = t; d->cap = nc;
    }
    d->p[d->n++] = v;
    return 0;
}

/* ---- flattening state ------------------------------------------------------ */

typedef struct { const ndtabi_object *ptr; int id; } ptr_id;

typedef struct {
    int n, np;
    const ndt_b200_host_api *host;
    void *host_module;            /* dli_fbase of the host's object.c */
    /* objects */
    const ndtabi_object **item;   /* id -> host object */
    int n_items, cap_items;
    ndt_flat_object *obj;
    int n_obj, cap_obj;
    dpool bs;                     /* (np+2) per object */
    dpool geom;
    /* kd */
    ndt_flat_node *node;
    int n_node, cap_node;
    ipool leaf;
    int max_leaf, depth;
    ptr_id *map;
} fstate;

static int cmp_ptr(const void *a, const void *b)
{
    const ptr_id *x = a, *y = b;
    if (x->ptr < y->ptr) return -1;
    if (x->ptr > y->ptr) return 1;
    return x->id - y->id;
}
static int id_of(const fstate *st, const void *p)
{
    int lo = 0, hi = st->n_items - 1, ans = -1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        if ((const void *)st->map[mid].ptr < p) lo = mid + 1;
        else { if ((const void *)st->map[mid].ptr == p) ans = st->map[mid].id; hi = mid - 1; }
    }
    return ans;
}

/* host vectNd -> np-wide array; the pad lane is read from the host's memory,
 * which always holds one (vectNd.h:128-148) */
static int load_vec(const fstate *st, const ndtabi_vec *v, double *dst)
{
    if (v == NULL || v->v == NULL || v->n != st->n)
        return -1;
    for (int i = 0; i < st->n; ++i) dst[i] = v->v[i];
    if (st->np > st->n) dst[st->n] = v->v[st->n];
    return 0;
}

static int type_of(const ndtabi_object *o, char *buf, size_t len)
{
    memset(buf, 0, len);
    if (!o->type_name) return -1;
    o->type_name(buf, (int)len - 1);
    return 0;
}

static int is_host_default(const fstate *st, void *fn)
{
    Dl_info di;
    if (fn == NULL) return 0;
    if (st->host_module == NULL) return 1; /* cannot tell: accept */
    if (!dladdr(fn, &di)) return 0;
    return di.dli_fbase == st->host_module;
}

static int items_add(fstate *st, const ndtabi_object *o)
{
    char tn[64];
    if (type_of(o, tn, sizeof tn))
        return ndt_set_error(NDT_B200_E_ARG, "object without type_name");
    if (!strcmp(tn, "cluster")) { /* object.c:636-643 */
        for (int i = 0; i < o->n_obj; ++i) {
            int r = items_add(st, o->obj[i]);
            if (r) return r;
        }
        return 0;
    }
    if (st->n_items == st->cap_items) {
        int nc = st->cap_items ? st->cap_items * 2 : 256;
        const ndtabi_object **t = realloc(st->item, (size_t)nc * sizeof *t);
        if (!t) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
        st->item = t; st->cap_items = nc;
    }
    st->item[st->n_items++] = o;
    return 0;
}

static ndt_flat_object *obj_new(fstate *st)
{
    if (st->n_obj == st->cap_obj) {
        int nc = st->cap_obj ? st->cap_obj * 2 : 256;
        ndt_flat_object *t = realloc(st->obj, (size_t)nc * sizeof *t);
        if (!t) return NULL;
        st->obj = t; st->cap_obj = nc;
    }
    ndt_flat_object *fo = &st->obj[st->n_obj++];
    memset(fo, 0, sizeof *fo);
    return fo;
}

/* make sure obj->bounds holds what the reference would have on its first ray
 * (object.c:608-615) */
static int force_bounds(const fstate *st, ndtabi_object *o)
{
    if (o->bounds.radius == 0) {
        if (!st->host || !st->host->object_get_bounds)
            return ndt_set_error(NDT_B200_E_ARG,
                "object '%s' has no bounding sphere yet and no host object_get_bounds was given", o->name);
        st->host->object_get_bounds(o);
    }
    return 0;
}

/* The same for many objects at once, on a pool of host threads.  object_get_bounds (object.c:582-603) works on
 * its object alone -- bounding_points of every shipped plugin only reads the object, bounds_list_optimal
 * (bounding.c:177-240) and the Nelder-Mead state behind it are local -- and each fit is deterministic, so which
 * thread runs it does not change a bit of the result.  The reference serialises these calls behind the mutex of
 * vect_object_intersect (object.c:608-615); BASELINE config 2 has 6561 of them, 3.1 s on one core. */
typedef struct {
    const fstate *st;
    ndtabi_object **objs;
    int n;
    volatile int next;
} bounds_job;

static void *bounds_worker(void *arg)
{
    bounds_job *j = arg;
    for (;;) {
        const int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        j->st->host->object_get_bounds(j->objs[i]);
    }
    return NULL;
}

static int force_bounds_many(const fstate *st, ndtabi_object **objs, int n)
{
    int todo = 0;
    for (int i = 0; i < n; ++i)
        if (objs[i]->bounds.radius == 0) objs[todo++] = objs[i];      /* compacts in place: the caller's array is scratch */
    if (todo == 0) return 0;
    if (!st->host || !st->host->object_get_bounds)
        return ndt_set_error(NDT_B200_E_ARG,
            "object '%s' has no bounding sphere yet and no host object_get_bounds was given", objs[0]->name);
    int nt = (int)sysconf(_SC_NPROCESSORS_ONLN);
    const char *e = getenv("NDT_B200_HOST_THREADS");
    if (e && atoi(e) > 0) nt = atoi(e);
    if (nt > 64) nt = 64;
    if (nt > todo / 8) nt = todo / 8;           /* a fit is ~0.5 ms: not worth a thread for a handful */
    bounds_job job = { st, objs, todo, 0 };
    pthread_t th[64];
    int started = 0;
    for (int t = 0; t + 1 < nt; ++t) {
        if (pthread_create(&th[started], NULL, bounds_worker, &job) != 0) break;
        ++started;
    }
    bounds_worker(&job);                        /* the calling thread works too */
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    return 0;
}

static ndtabi_vec *tmp_vec(int n)
{
    ndtabi_vec *v = NULL;
    if (posix_memalign((void **)&v, 16, sizeof *v)) return NULL;
    memset(v, 0, sizeof *v);
    v->n = n;
    if (n > 4) {
        void *p = NULL;
        if (posix_memalign(&p, 16, (size_t)(n + (n & 1)) * sizeof(double))) { free(v); return NULL; }
        v->v = p;
    } else {
        v->v = v->inl;
    }
    memset(v->v, 0, (size_t)(n + (n & 1)) * sizeof(double));
    return v;
}
static void tmp_vec_free(ndtabi_vec *v)
{
    if (!v) return;
    if (v->n > 4) free(v->v);
    free(v);
}

/* hcube builds its faces inside prepare() on the first intersect call
 * (hcube.c:155-170, 238-240); make that call */
static int force_hcube(const fstate *st, ndtabi_object *o)
{
    if (o->prepared && o->n_obj > 0)
        return 0;
    if (!o->intersect || o->n_pos < 1)
        return ndt_set_error(NDT_B200_E_ARG, "hcube '%s' cannot be prepared", o->name);
    ndtabi_vec *a = tmp_vec(st->n), *b = tmp_vec(st->n), *c = tmp_vec(st->n), *d = tmp_vec(st->n);
    if (!a || !b || !c || !d)
        return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    for (int i = 0; i < st->n; ++i) a->v[i] = o->pos[0].v[i];
    b->v[0] = 1.0;
    ndtabi_object *hit = NULL;
    o->intersect(o, a, b, c, d, &hit);
    tmp_vec_free(a); tmp_vec_free(b); tmp_vec_free(c); tmp_vec_free(d);
    if (o->n_obj <= 0)
        return ndt_set_error(NDT_B200_E_ARG, "hcube '%s' produced no faces", o->name);
    return 0;
}

#define NEED(cond, ...) do { if (!(cond)) return ndt_set_error(NDT_B200_E_ARG, __VA_ARGS__); } while (0)
#define TAKE(ptr, cnt) do { (ptr) = dpool_take(&st->geom, (size_t)(cnt)); \
        if (!(ptr)) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory"); } while (0)

/* one host object -> objects[slot] (+ bsphere + geometry).  Nested objects
 * of an hcube are appended by the caller afterwards. */
static int emit_object(fstate *st, ndtabi_object *o, int slot, int report_id)
{
    const int np = st->np, n = st->n;
    char tn[64];
    type_of(o, tn, sizeof tn);
    ndt_flat_object *fo = &st->obj[slot];
    int r;

    NEED(o->dimensions == n, "object '%s' has %d dimensions, scene has %d", o->name, o->dimensions, n);
    if (!is_host_default(st, (void *)o->get_color) || !is_host_default(st, (void *)o->get_reflect))
        return ndt_set_error(NDT_B200_E_UNSUPPORTED,
            "object '%s' (%s) overrides get_color/get_reflect; the device path shades with the "
            "default material only (object.c:23-43)", o->name, tn);

    if ((r = force_bounds(st, o))) return r;

    fo->report_id = report_id;
    fo->flags = o->transparent ? NDT_OF_TRANSPARENT : 0;
    for (int k = 0; k < 3; ++k) { fo->rgb[k] = o->rgb[k]; fo->refl[k] = o->refl[k]; }
    fo->refract_index = o->refract_index;
    fo->child_begin = -1;
    fo->child_count = 0;

    /* bounding sphere: bounding.c:16-28 computes radius_sqr = radius*radius */
    double *bs = st->bs.p + (size_t)slot * (np + 2);
    memset(bs, 0, (size_t)(np + 2) * sizeof(double));
    if (o->bounds.radius > 0) {
        if (load_vec(st, &o->bounds.center, bs))
            return ndt_set_error(NDT_B200_E_ARG, "object '%s': bad bounds centre", o->name);
    }
    bs[np] = o->bounds.radius;
    bs[np + 1] = o->bounds.radius * o->bounds.radius;
    fo->bs_radius = o->bounds.radius;

    /* every geometry block starts 16-byte aligned (even double index): the device stages it
     * into shared memory with one TMA bulk copy (cp.async.bulk needs 16-byte addresses) */
    if (st->geom.n & 1) { double *pad; TAKE(pad, 1); pad[0] = 0.0; }
    size_t g0 = st->geom.n;
    NEED(g0 < 0xffffffffu, "geometry pool too large");
    fo->geom_off = (uint32_t)g0;
    double *g;

    if (!strcmp(tn, "sphere")) {                       /* sphere.c:18-32 */
        fo->type = NDT_T_SPHERE;
        NEED(o->n_pos >= 1 && o->n_size >= 1, "sphere '%s' incomplete", o->name);
        TAKE(g, np + 1);
        NEED(!load_vec(st, &o->pos[0], g), "sphere '%s': bad vector", o->name);
        g[np] = pow(o->size[0], 2.0);
    } else if (!strcmp(tn, "hplane")) {                /* hplane.c:39-52 */
        fo->type = NDT_T_HPLANE;
        NEED(o->n_pos >= 1 && o->n_dir >= 1, "hplane '%s' incomplete", o->name);
        TAKE(g, 2 * np);
        NEED(!load_vec(st, &o->pos[0], g) && !load_vec(st, &o->dir[0], g + np), "hplane '%s': bad vector", o->name);
    } else if (!strcmp(tn, "hdisk")) {                 /* hdisk.c:15-34: inner hplane = copies of pos[0], dir[0] */
        fo->type = NDT_T_HDISK;
        NEED(o->n_pos >= 1 && o->n_dir >= 1 && o->n_size >= 1, "hdisk '%s' incomplete", o->name);
        TAKE(g, 2 * np + 1);
        const ndtabi_vec *pp = &o->pos[0], *pd = &o->dir[0];
        if (o->prepared && o->n_obj > 0 && o->obj[0]->n_pos > 0 && o->obj[0]->n_dir > 0) {
            pp = &o->obj[0]->pos[0];   /* the plane that was frozen at prepare time */
            pd = &o->obj[0]->dir[0];
        }
        NEED(!load_vec(st, pp, g) && !load_vec(st, pd, g + np), "hdisk '%s': bad vector", o->name);
        /* the inside test measures from the disk's own pos[0] (hdisk.c:77) */
        if (pp != &o->pos[0]) {
            double own[NDT_MAX_DIM];
            NEED(!load_vec(st, &o->pos[0], own), "hdisk '%s': bad vector", o->name);
            NEED(!memcmp(own, g, (size_t)n * sizeof(double)),
                 "hdisk '%s' moved after it was prepared; unsupported", o->name);
        }
        g[2 * np] = o->size[0];
    } else if (!strcmp(tn, "orthotope")) {             /* orthotope.c:23-54 */
        fo->type = NDT_T_ORTHOTOPE;
        NEED(o->n_flag >= 1 && o->n_pos >= 1, "orthotope '%s' incomplete", o->name);
        int m = o->flag[0];
        NEED(m >= 1 && m <= n && o->n_dir >= m, "orthotope '%s': %d axes in %d-D", o->name, m, n);
        fo->n_axes = m;
        TAKE(g, np + (size_t)m * np + 3 * (size_t)m);
        double *p0 = g, *basis = g + np, *len = basis + (size_t)m * np, *bdb = len + m, *bdp = bdb + m;
        NEED(!load_vec(st, &o->pos[0], p0), "orthotope '%s': bad vector", o->name);
        for (int i = 0; i < m; ++i) {
            double dir[NDT_MAX_DIM], *b = basis + (size_t)i * np;
            NEED(!load_vec(st, &o->dir[i], dir), "orthotope '%s': bad vector", o->name);
            v_copy_n(b, dir, n);              /* fresh vector: pad lane 0 (vectNd.h:146) */
            v_unitize(b, np);
            len[i] = v_norm(dir, np);
            bdb[i] = v_dot(b, b, np);
            bdp[i] = v_dot(p0, b, np);
        }
    } else if (!strcmp(tn, "hcube")) {                 /* hcube.c:155-170 */
        fo->type = NDT_T_HCUBE;
        if ((r = force_hcube(st, o))) return r;
        fo = &st->obj[slot];
    } else if (!strcmp(tn, "facet")) {                 /* facet.c:42-83 */
        fo->type = NDT_T_FACET;
        NEED(o->n_pos >= 3 && o->n_dir >= 1, "facet '%s' incomplete", o->name);
        TAKE(g, 6 * np + 7);
        double *p = g, *basis = g + 3 * np, *nrm = g + 5 * np, *sc = g + 6 * np;
        double edge[3][NDT_MAX_DIM];
        for (int i = 0; i < 3; ++i)
            NEED(!load_vec(st, &o->pos[i], p + (size_t)i * np), "facet '%s': bad vector", o->name);
        for (int i = 0; i < 3; ++i) {
            int j = (i + 1) % 3, k = (i + 2) % 3;
            double a[NDT_MAX_DIM], b[NDT_MAX_DIM];
            v_sub(p + (size_t)j * np, p + (size_t)i * np, edge[i], np);
            /* vectNd_angle3(pos[k], pos[i], pos[j]) (vectNd.c:83-99) */
            v_sub(p + (size_t)k * np, p + (size_t)i * np, a, np);
            v_sub(p + (size_t)j * np, p + (size_t)i * np, b, np);
            sc[4 + i] = v_angle(a, b, np);
        }
        /* vectNd_orthogonalize(edge0, edge1 -> basis0, basis1) (vectNd.c:35-58) */
        {
            double bb = v_dot(edge[1], edge[1], np);
            double ab = v_dot(edge[0], edge[1], np);
            double t[NDT_MAX_DIM];
            v_scale(edge[1], ab / bb, t, np);
            v_sub(edge[0], t, basis, np);
            v_copy_n(basis + np, edge[1], n);
            v_unitize(basis, np);
            v_unitize(basis + np, np);
        }
        NEED(!load_vec(st, &o->dir[0], nrm), "facet '%s': bad vector", o->name);
        for (int i = 0; i < 2; ++i) {                  /* facet.c:189-201: ray-invariant dots */
            sc[i] = v_dot(basis + (size_t)i * np, basis + (size_t)i * np, np);
            sc[2 + i] = v_dot(p + np, basis + (size_t)i * np, np);
        }
    } else if (!strcmp(tn, "hfacet")) {                /* hfacet.c:43-92 */
        fo->type = NDT_T_HFACET;
        NEED(o->n_pos >= 3 && o->n_flag >= 1, "hfacet '%s' incomplete", o->name);
        int use_normals = o->flag[0] != 0;
        NEED(!use_normals || o->n_dir >= 3, "hfacet '%s' needs three normals", o->name);
        if (use_normals) fo->flags |= NDT_OF_USE_NORMALS;
        TAKE(g, 6 * np + 5);
        double *v0 = g, *ue0 = g + np, *eperp = g + 2 * np, *nrm = g + 3 * np, *sc = g + 6 * np;
        double vtx[3][NDT_MAX_DIM], edge[3][NDT_MAX_DIM];
        for (int i = 0; i < 3; ++i)
            NEED(!load_vec(st, &o->pos[i], vtx[i]), "hfacet '%s': bad vector", o->name);
        memcpy(v0, vtx[0], (size_t)np * sizeof(double));
        for (int i = 0; i < 3; ++i)
            v_sub(vtx[(i + 1) % 3], vtx[i], edge[i], np);
        v_copy_n(ue0, edge[0], n);
        v_unitize(ue0, np);
        v_scale(edge[2], -1.0, edge[2], np);
        {   /* edge_perp = unit(edge2 - proj(edge2 onto edge0)) (hfacet.c:78-85) */
            double bb = v_dot(edge[0], edge[0], np);
            double ab = v_dot(edge[2], edge[0], np);
            double t[NDT_MAX_DIM];
            v_scale(edge[0], ab / bb, t, np);
            v_sub(edge[2], t, eperp, np);
            v_unitize(eperp, np);
        }
        if (use_normals)
            for (int i = 0; i < 3; ++i)
                NEED(!load_vec(st, &o->dir[i], nrm + (size_t)i * np), "hfacet '%s': bad vector", o->name);
        sc[0] = v_dot(ue0, edge[0], np);               /* x2 (hfacet.c:168) */
        sc[1] = v_dot(eperp, edge[0], np);             /* y2 */
        sc[2] = v_dot(ue0, edge[2], np);               /* x3 (hfacet.c:172) */
        sc[3] = v_dot(eperp, edge[2], np);             /* y3 */
        sc[4] = (n & 1) ? 1.0 : 0.0;                   /* pad lane of the all-ones vector (hfacet.c:52-54) */
    } else if (!strcmp(tn, "cylinder")) {              /* cylinder.c:22-41 */
        fo->type = NDT_T_CYLINDER;
        NEED(o->n_pos >= 2 && o->n_size >= 1, "cylinder '%s' incomplete", o->name);
        if (o->n_flag > 1 && o->flag[1] != 0) fo->flags |= NDT_OF_NO_END_TEST;
        TAKE(g, 2 * np + 4);
        double *p0 = g, *axis = g + np, *sc = g + 2 * np, p1[NDT_MAX_DIM];
        NEED(!load_vec(st, &o->pos[0], p0) && !load_vec(st, &o->pos[1], p1), "cylinder '%s': bad vector", o->name);
        v_sub(p1, p0, axis, np);
        v_unitize(axis, np);
        sc[0] = v_dist(p1, p0, np);
        sc[1] = v_dot(axis, axis, np);
        sc[2] = v_dot(p0, axis, np);
        sc[3] = o->size[0];
    } else if (!strcmp(tn, "hcylinder")) {             /* hcylinder.c:23-54 */
        fo->type = NDT_T_HCYLINDER;
        int a = n - 2;
        NEED(a >= 1 && o->n_pos >= a + 1 && o->n_size >= 1, "hcylinder '%s' incomplete", o->name);
        if (o->n_flag != 0 && o->flag[0] != 0) fo->flags |= NDT_OF_NO_END_TEST;
        fo->n_axes = a;
        TAKE(g, np + (size_t)a * np + 3 * (size_t)a + 1);
        double *p0 = g, *axes = g + np, *len = axes + (size_t)a * np, *ada = len + a, *bda = ada + a;
        NEED(!load_vec(st, &o->pos[0], p0), "hcylinder '%s': bad vector", o->name);
        for (int i = 0; i < a; ++i) {
            double pi[NDT_MAX_DIM], *ax = axes + (size_t)i * np;
            NEED(!load_vec(st, &o->pos[i + 1], pi), "hcylinder '%s': bad vector", o->name);
            v_sub(pi, p0, ax, np);
            v_unitize(ax, np);
            len[i] = v_dist(pi, p0, np);
            ada[i] = v_dot(ax, ax, np);
            bda[i] = v_dot(p0, ax, np);
        }
        bda[a] = o->size[0];
    } else {
        return ndt_set_error(NDT_B200_E_UNSUPPORTED,
            "object type '%s' ('%s') has no device kernel", tn, o->name);
    }
    return 0;
}

static int reserve_slots(fstate *st, int count)
{
    for (int i = 0; i < count; ++i)
        if (!obj_new(st)) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    if (dpool_need(&st->bs, (size_t)count * (st->np + 2)))
        return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    st->bs.n += (size_t)count * (st->np + 2);
    return 0;
}

static int walk_node(fstate *st, const ndtabi_kd_node *hn, int depth, int *out_idx)
{
    if (hn == NULL) { *out_idx = -1; return 0; }
    if (st->n_node == st->cap_node) {
        int nc = st->cap_node ? st->cap_node * 2 : 256;
        ndt_flat_node *t = realloc(st->node, (size_t)nc * sizeof *t);
        if (!t) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
        st->node = t; st->cap_node = nc;
    }
    int me = st->n_node++;
    *out_idx = me;
    if (depth > st->depth) st->depth = depth;
    ndt_flat_node fn;
    memset(&fn, 0, sizeof fn);
    fn.dim = hn->dim;
    fn.boundary = hn->boundary;
    fn.leaf_begin = (int32_t)st->leaf.n;
    fn.leaf_count = hn->num > 0 ? hn->num : 0;
    fn.left = fn.right = -1;
    for (int i = 0; i < fn.leaf_count; ++i) {
        int id = hn->obj_ids[i];
        NEED(id >= 0 && id < st->n_items && (const void *)st->item[id] == hn->objs[i],
             "kd leaf refers to object id %d which is not item %d of the scene walk", id, id);
        if (ipool_push(&st->leaf, id)) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    }
    if (fn.leaf_count > st->max_leaf) st->max_leaf = fn.leaf_count;
    NEED(fn.dim < st->n, "kd node splits dimension %d of %d", fn.dim, st->n);
    if (fn.dim >= 0) {   /* kd-tree.c:515-519: only non-leaves descend */
        int r;
        if ((r = walk_node(st, hn->left, depth + 1, &fn.left))) return r;
        if ((r = walk_node(st, hn->right, depth + 1, &fn.right))) return r;
    }
    st->node[me] = fn;
    return 0;
}

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

int ndt_b200_flatten(const void *scene_v, const void *kdtree_v, int width, int height,
                     int max_optic_depth, int specular,
                     const ndt_b200_host_api *host, ndt_flat_scene **out)
{
    return ndt_b200_flatten_view(scene_v, kdtree_v, width, height, max_optic_depth, specular, NDT_MONO, host, out);
}

/* temporary host vectNd (vectNd.h:42-51 layout) around a plain array, for calling vectNd_rotate2 */
static void tmpvec_wrap(ndtabi_vec *v, double *store, int n)
{
    memset(v, 0, sizeof *v);
    v->v = store;
    v->n = n;
}

static int flatten_impl(const void *scene_v, const void *kdtree_v, int width, int height,
                        int max_optic_depth, int specular, int stereo_mode, int aa_pad,
                        const ndt_b200_host_api *host, ndt_flat_scene **out);

int ndt_b200_flatten_view(const void *scene_v, const void *kdtree_v, int width, int height,
                          int max_optic_depth, int specular, int stereo_mode,
                          const ndt_b200_host_api *host, ndt_flat_scene **out)
{
    return flatten_impl(scene_v, kdtree_v, width, height, max_optic_depth, specular, stereo_mode, 0, host, out);
}

/* recursive anti-aliasing (-w / -a, ndt.c:921-926): the initial image is one sample larger in each
 * direction and normalised by width+1 / height+1, while cam.dirX is scaled by width/height */
int ndt_b200_flatten_aa(const void *scene_v, const void *kdtree_v, int width, int height,
                        int max_optic_depth, int specular,
                        const ndt_b200_host_api *host, ndt_flat_scene **out)
{
    return flatten_impl(scene_v, kdtree_v, width, height, max_optic_depth, specular, NDT_MONO, 1, host, out);
}

static int flatten_impl(const void *scene_v, const void *kdtree_v, int out_width, int out_height,
                        int max_optic_depth, int specular, int stereo_mode, int aa_pad,
                        const ndt_b200_host_api *host, ndt_flat_scene **out)
{
    const int width = out_width + (aa_pad ? 1 : 0), height = out_height + (aa_pad ? 1 : 0);
    const ndtabi_scene *scn = scene_v;
    const ndtabi_kd_tree *kd = kdtree_v;
    fstate S, *st = &S;
    int r = 0;
    memset(st, 0, sizeof S);
    if (out) *out = NULL;
    if (!scn || !kd || !out || width <= 0 || height <= 0)
        return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_flatten: NULL scene/kdtree/out or empty frame");
    st->n = scn->dimensions;
    if (st->n < 3 || st->n > NDT_MAX_DIM - 2)
        return ndt_set_error(NDT_B200_E_UNSUPPORTED, "%d dimensions: device kernels cover 3..%d", st->n, NDT_MAX_DIM - 2);
    st->np = st->n + (st->n & 1);
    st->host = host;
    if (host && host->object_get_bounds) {
        Dl_info di;
        if (dladdr((void *)host->object_get_bounds, &di)) st->host_module = di.dli_fbase;
    }
    const int n = st->n, np = st->np;

    if (scn->cam.type < 0 || scn->cam.type > 2) {
        r = ndt_set_error(NDT_B200_E_UNSUPPORTED, "camera type %d (camera.h:16-20 knows NORMAL, VR, PANO)", scn->cam.type);
        goto done;
    }
    if (stereo_mode < NDT_MONO || stereo_mode > NDT_HIDEF_3D) {
        r = ndt_set_error(NDT_B200_E_UNSUPPORTED, "stereo mode %d (ndt.c:46-48)", stereo_mode);
        goto done;
    }
    const int cam_type = scn->cam.type;
    if (aa_pad && (cam_type != NDT_CAM_NORMAL || stereo_mode != NDT_MONO)) {
        r = ndt_set_error(NDT_B200_E_UNSUPPORTED, "recursive anti-aliasing is on the device path for CAMERA_NORMAL + MONO only");
        goto done;
    }
    if (aa_pad && scn->cam.aperture_radius != 0.0) {
        r = ndt_set_error(NDT_B200_E_UNSUPPORTED, "recursive anti-aliasing with aperture_radius %g: the depth-of-field "
                          "jitter draws from drand48 (ndt.c:528-542) and is not reproducible", scn->cam.aperture_radius);
        goto done;
    }
    const int has_view = cam_type != NDT_CAM_NORMAL || stereo_mode != NDT_MONO;
    /* VR / PANO in a stereo mode: the eye is rotated per column (ndt.c:519-525) */
    const int view_eyes = cam_type != NDT_CAM_NORMAL && stereo_mode != NDT_MONO;
    if (view_eyes && !(host && host->vectNd_rotate2)) {
        r = ndt_set_error(NDT_B200_E_ARG, "VR / PANO camera in a stereo mode needs host->vectNd_rotate2 (vectNd.c:271)");
        goto done;
    }

    /* 1. kd items in object_kdlist_add order (ndt.c:1903-1907) */
    for (int i = 0; i < scn->num_objects; ++i)
        if ((r = items_add(st, scn->object_ptrs[i]))) goto done;
    if (st->n_items != kd->obj_num) {
        r = ndt_set_error(NDT_B200_E_ARG, "scene walk found %d objects, kd-tree holds %d", st->n_items, kd->obj_num);
        goto done;
    }
    st->map = malloc((size_t)(st->n_items ? st->n_items : 1) * sizeof *st->map);
    if (!st->map) { r = ndt_set_error(NDT_B200_E_NOMEM, "out of memory"); goto done; }
    for (int i = 0; i < st->n_items; ++i) { st->map[i].ptr = st->item[i]; st->map[i].id = i; }
    qsort(st->map, (size_t)st->n_items, sizeof *st->map, cmp_ptr);

    /* 2. objects: top level first, nested ones after.  The lazily fitted bounding spheres first, all at once */
    {
        ndtabi_object **tmp = malloc((size_t)(st->n_items ? st->n_items : 1) * sizeof *tmp);
        if (!tmp) { r = ndt_set_error(NDT_B200_E_NOMEM, "out of memory"); goto done; }
        for (int i = 0; i < st->n_items; ++i) tmp[i] = (ndtabi_object *)st->item[i];
        r = force_bounds_many(st, tmp, st->n_items);
        free(tmp);
        if (r) goto done;
    }
    if ((r = reserve_slots(st, st->n_items))) goto done;
    for (int i = 0; i < st->n_items; ++i)
        if ((r = emit_object(st, (ndtabi_object *)st->item[i], i, i))) goto done;
    for (int i = 0; i < st->n_items; ++i) {
        if (st->obj[i].type != NDT_T_HCUBE) continue;
        const ndtabi_object *hc = st->item[i];
        int begin = st->n_obj, cnt = hc->n_obj;
        if ((r = reserve_slots(st, cnt))) goto done;
        st->obj[i].child_begin = begin;
        st->obj[i].child_count = cnt;
        {
            ndtabi_object **tmp = malloc((size_t)(cnt ? cnt : 1) * sizeof *tmp);
            if (!tmp) { r = ndt_set_error(NDT_B200_E_NOMEM, "out of memory"); goto done; }
            for (int c = 0; c < cnt; ++c) tmp[c] = hc->obj[c];
            r = force_bounds_many(st, tmp, cnt);
            free(tmp);
            if (r) goto done;
        }
        for (int c = 0; c < cnt; ++c) {
            if ((r = emit_object(st, hc->obj[c], begin + c, i))) goto done;
            if (st->obj[begin + c].type == NDT_T_HCUBE) {
                r = ndt_set_error(NDT_B200_E_UNSUPPORTED, "hcube nested in hcube");
                goto done;
            }
        }
    }

    /* 3. kd-tree */
    {
        int root = -1;
        if ((r = walk_node(st, kd->root, 0, &root))) goto done;
        if (root != 0 && !(root == -1 && st->n_node == 0)) {
            r = ndt_set_error(NDT_B200_E_ARG, "kd-tree root not first"); goto done;
        }
    }

    /* 4. lights */
    int n_l = scn->num_lights;
    ndt_flat_light *fl = calloc((size_t)(n_l ? n_l : 1), sizeof *fl);
    if (!fl) { r = ndt_set_error(NDT_B200_E_NOMEM, "out of memory"); goto done; }
    for (int i = 0; i < n_l; ++i) {
        const ndtabi_light *l = scn->lights[i];
        fl[i].rgb[0] = l->rgb[0]; fl[i].rgb[1] = l->rgb[1]; fl[i].rgb[2] = l->rgb[2];
        fl[i].max_rgb = REF_MAX(l->rgb[0], REF_MAX(l->rgb[1], l->rgb[2]));
        fl[i].angle = l->angle;
        double *g = dpool_take(&st->geom, 4 * (size_t)np);
        if (!g) { free(fl); r = ndt_set_error(NDT_B200_E_NOMEM, "out of memory"); goto done; }
        fl[i].vec_off = (uint32_t)(g - st->geom.p);
        double *pos = g, *dir = g + np, *rev = g + 2 * np, *near_off = g + 3 * np;
        switch (l->type) {
        case NDTABI_LIGHT_AMBIENT: fl[i].type = NDT_L_AMBIENT; break;
        case NDTABI_LIGHT_POINT:   fl[i].type = NDT_L_POINT; break;
        case NDTABI_LIGHT_SPOT:    fl[i].type = NDT_L_SPOT; break;
        case NDTABI_LIGHT_DIRECTIONAL: fl[i].type = NDT_L_DIRECTIONAL; break;
        default:
            free(fl);
            r = ndt_set_error(NDT_B200_E_UNSUPPORTED, "light %d is an area light (type %d): its drand48 sampling "
                              "(ndt.c:116-147) is not reproducible and not on the device path", i, l->type);
            goto done;
        }
        if (fl[i].type == NDT_L_POINT || fl[i].type == NDT_L_SPOT)
            if (load_vec(st, &l->pos, pos)) { free(fl); r = ndt_set_error(NDT_B200_E_ARG, "light %d: bad pos", i); goto done; }
        if (fl[i].type == NDT_L_DIRECTIONAL || fl[i].type == NDT_L_SPOT) {
            if (load_vec(st, &l->dir, dir)) { free(fl); r = ndt_set_error(NDT_B200_E_ARG, "light %d: bad dir", i); goto done; }
            v_scale(dir, -1, rev, np);       /* ndt.c:156-158 */
            v_unitize(rev, np);
            v_copy_n(near_off, dir, n);      /* ndt.c:234-236 */
            v_unitize(near_off, np);
            v_scale(near_off, -EPS, near_off, np);
        }
    }

    /* 5. assemble the blob */
    size_t off = align16(sizeof(ndt_flat_header));
    ndt_flat_header H;
    memset(&H, 0, sizeof H);
    H.magic = NDT_FLAT_MAGIC; H.version = NDT_FLAT_VERSION;
    H.n = n; H.npad = np; H.width = width; H.height = height;
    H.max_optic_depth = max_optic_depth; H.specular = specular ? 1 : 0;
    H.n_items = st->n_items; H.n_objects = st->n_obj;
    H.n_nodes = st->n_node; H.n_leaf_refs = (int32_t)st->leaf.n;
    H.n_inf = kd->inf_obj_num; H.n_lights = n_l;
    H.max_leaf = st->max_leaf; H.tree_depth = st->depth;
    for (int k = 0; k < 4; ++k) H.bg[k] = scn->bg[k];
    for (int k = 0; k < 3; ++k) H.ambient[k] = scn->ambient.rgb[k];
    H.off_camera = off;   off = align16(off + 4 * (size_t)np * 8);
    H.off_aabb = off;     off = align16(off + 2 * (size_t)np * 8);
    H.off_objects = off;  off = align16(off + (size_t)st->n_obj * sizeof(ndt_flat_object));
    H.off_bspheres = off; off = align16(off + (size_t)st->n_obj * (np + 2) * 8);
    H.off_geom = off;     H.n_geom = st->geom.n; off = align16(off + st->geom.n * 8);
    H.off_nodes = off;    off = align16(off + (size_t)st->n_node * sizeof(ndt_flat_node));
    H.off_leaf_refs = off; off = align16(off + st->leaf.n * 4);
    H.off_inf = off;      off = align16(off + (size_t)kd->inf_obj_num * 4);
    H.off_lights = off;   off = align16(off + (size_t)n_l * sizeof(ndt_flat_light));
    H.cam_type = cam_type; H.stereo_mode = stereo_mode; H.view_eyes = view_eyes; H.aa_pad = aa_pad ? 1 : 0;
    H.cam_dist = scn->cam.focal_distance;
    if (has_view) {
        H.off_view = off;
        off = align16(off + ((size_t)5 * np + (size_t)width * 4 + (size_t)height * 6 +
                             (view_eyes ? (size_t)width * 2 * np : 0)) * 8);
    }
    H.total_bytes = off;

    char *blob = NULL;
    if (posix_memalign((void **)&blob, 256, off)) { free(fl); r = ndt_set_error(NDT_B200_E_NOMEM, "out of memory"); goto done; }
    memset(blob, 0, off);

    /* camera: camera.c:557-575 with the ray-invariant part hoisted */
    {
        double *cam = (double *)(blob + H.off_camera);
        double *pos = cam, *orig = cam + np, *dx = cam + 2 * np, *dy = cam + 3 * np;
        if (load_vec(st, &scn->cam.pos, pos) || load_vec(st, &scn->cam.imgOrig, orig) ||
            load_vec(st, &scn->cam.dirX, dx) || load_vec(st, &scn->cam.dirY, dy)) {
            free(blob); free(fl);
            r = ndt_set_error(NDT_B200_E_ARG, "camera vectors not aimed (camera_aim must run first, ndt.c:1925)");
            goto done;
        }
        if (stereo_mode != NDT_HIDEF_3D) v_scale(dx, out_width / (double)out_height, dx, np);   /* ndt.c:925-929 */
        else v_scale(dx, out_width / (double)1080, dx, np);
        double sd = v_dist(orig, pos, np);             /* camera.c:567 */
        H.use_focal = sd > EPS;
        H.focal_scale = H.use_focal ? scn->cam.focal_distance / sd : 0.0;
    }
    if (has_view) {
        /* render_pixel (ndt.c:578-653) per column / row, camera_target_point's trigonometry
         * (camera.c:507-556) and the per-column eye of VR stereo (ndt.c:519-525), evaluated here
         * with the host's libm so that the device reproduces the reference's primary rays exactly */
        double *ext = (double *)(blob + H.off_view);
        double *cols = ext + 5 * (size_t)np, *rows = cols + (size_t)width * 4, *eyes = rows + (size_t)height * 6;
        const ndtabi_camera *cam = &scn->cam;
        if (load_vec(st, &cam->leftEye, ext) || load_vec(st, &cam->rightEye, ext + np) ||
            load_vec(st, &cam->localX, ext + 2 * np) || load_vec(st, &cam->localY, ext + 3 * np) ||
            load_vec(st, &cam->localZ, ext + 4 * np)) {
            free(blob); free(fl);
            r = ndt_set_error(NDT_B200_E_ARG, "camera eyes / local axes not aimed (camera_aim must run first)");
            goto done;
        }
        const double x_scale = stereo_mode == NDT_SIDE_SIDE_3D ? 0.5 : 1.0;
        const double y_scale = stereo_mode == NDT_OVER_UNDER_3D ? 0.5 : 1.0;
        const double dist = cam->focal_distance;
        for (int i = 0; i < width; ++i) {
            double ip = i;
            int eye = 0;
            if (stereo_mode == NDT_SIDE_SIDE_3D) {
                if (i < width / 2) { ip = ip / x_scale; eye = 1; }
                else { ip = (ip - width / 2) / x_scale; eye = 2; }
            }
            const double x = ip / (double)width - 0.5;
            const double azi = x * cam->hFov;
            double *c = cols + (size_t)i * 4;
            c[0] = x; c[1] = sin(azi); c[2] = cos(azi); c[3] = eye;
            if (view_eyes) {
                for (int e = 0; e < 2; ++e) {
                    double tmp[NDT_MAX_DIM] __attribute__((aligned(16))) = {0};
                    ndtabi_vec vv;
                    double *dst = eyes + ((size_t)i * 2 + e) * np;
                    memcpy(tmp, ext + (size_t)e * np, (size_t)np * 8);
                    tmpvec_wrap(&vv, tmp, n);
                    host->vectNd_rotate2(&vv, (void *)&cam->pos, (void *)&cam->localX, (void *)&cam->localZ, azi, &vv);
                    memcpy(dst, tmp, (size_t)np * 8);
                }
            }
        }
        const double y_size = 2.0 * tan(cam->vFov / 2.0) * dist;      /* camera.c:540 */
        for (int j = 0; j < height; ++j) {
            double jp = j;
            int eye = 0, blank = 0;
            if (stereo_mode == NDT_OVER_UNDER_3D) {
                if (j < height / 2) { jp = jp / y_scale; eye = 1; }
                else { jp = (jp - height / 2) / y_scale; eye = 2; }
            }
            double y;
            if (stereo_mode == NDT_HIDEF_3D) {
                if (j < 1080) eye = 1;
                else if (j > (1080 + 45)) { jp = j - (1080 + 45); eye = 2; }
                else blank = 1;
                y = -(jp / 1080.0 - 0.5);
            } else {
                y = -(jp / (double)height - 0.5);
            }
            const double alt = y * cam->vFov;
            double *rr = rows + (size_t)j * 6;
            rr[0] = y; rr[1] = sin(alt); rr[2] = cos(alt); rr[3] = y * y_size; rr[4] = blank; rr[5] = eye;
        }
    }
    {
        double *bb = (double *)(blob + H.off_aabb);
        for (int i = 0; i < n; ++i) { bb[i] = kd->bb_lower.v[i]; bb[np + i] = kd->bb_upper.v[i]; }
    }
    memcpy(blob + H.off_objects, st->obj, (size_t)st->n_obj * sizeof(ndt_flat_object));
    memcpy(blob + H.off_bspheres, st->bs.p, (size_t)st->n_obj * (np + 2) * 8);
    if (st->geom.n) memcpy(blob + H.off_geom, st->geom.p, st->geom.n * 8);
    if (st->n_node) memcpy(blob + H.off_nodes, st->node, (size_t)st->n_node * sizeof(ndt_flat_node));
    if (st->leaf.n) memcpy(blob + H.off_leaf_refs, st->leaf.p, st->leaf.n * 4);
    {
        int32_t *inf = (int32_t *)(blob + H.off_inf);
        for (int i = 0; i < kd->inf_obj_num; ++i) {
            int id = id_of(st, kd->inf_obj_ptrs[i]);
            if (id < 0) {
                free(blob); free(fl);
                r = ndt_set_error(NDT_B200_E_ARG, "infinite object %d of the kd-tree is not in the scene", i);
                goto done;
            }
            inf[i] = id;
        }
    }
    if (n_l) memcpy(blob + H.off_lights, fl, (size_t)n_l * sizeof(ndt_flat_light));
    free(fl);
    memcpy(blob, &H, sizeof H);
    *out = (ndt_flat_scene *)blob;

done:
    free(st->item); free(st->obj); free(st->bs.p); free(st->geom.p);
    free(st->node); free(st->leaf.p); free(st->map);
    return r;
}

void ndt_b200_free_flat(ndt_flat_scene *fs) { free(fs); }

/* doubles of an object's geometry block (layouts in ndt_flat.h; the device's copy is geom_block_doubles, warp.cuh) */
static int flat_geom_doubles(int type, int m, int np)
{
    switch (type) {
    case NDT_T_SPHERE: return np + 1;
    case NDT_T_HPLANE: return 2 * np;
    case NDT_T_HDISK: return 2 * np + 1;
    case NDT_T_ORTHOTOPE: return np + m * np + 3 * m;
    case NDT_T_FACET: return 6 * np + 7;
    case NDT_T_HFACET: return 6 * np + 5;
    case NDT_T_CYLINDER: return 2 * np + 4;
    case NDT_T_HCYLINDER: return np + m * np + 3 * m + 1;
    }
    return 0;
}

/* `bytes` is what the caller can vouch for: a blob read from disk or received from another rank passes its
 * length; the in-memory entry points (ndt_b200_upload) pass header.total_bytes, i.e. THEIR caller guarantees
 * that many readable bytes (include/ndt_b200.h) */
int ndt_b200_flat_validate(const void *blob, size_t bytes)
{
    const ndt_flat_header *h = blob;
    if (!blob || bytes < sizeof *h) return ndt_set_error(NDT_B200_E_ARG, "flat scene: truncated header");
    if (h->magic != NDT_FLAT_MAGIC || h->version != NDT_FLAT_VERSION)
        return ndt_set_error(NDT_B200_E_ARG, "flat scene: bad magic/version");
    if (h->total_bytes != bytes) return ndt_set_error(NDT_B200_E_ARG, "flat scene: %zu bytes given, header says %llu", bytes, (unsigned long long)h->total_bytes);
    if (h->n < 3 || h->n > NDT_MAX_DIM - 2 || h->npad != h->n + (h->n & 1))
        return ndt_set_error(NDT_B200_E_ARG, "flat scene: bad dimensions");
    if (h->n_items < 0 || h->n_objects < h->n_items || h->n_nodes < 0 || h->n_leaf_refs < 0 ||
        h->n_inf < 0 || h->n_lights < 0 || h->width <= 0 || h->height <= 0)
        return ndt_set_error(NDT_B200_E_ARG, "flat scene: bad counts");
    const size_t np = (size_t)h->npad;
#define IN(off, len) ((off) % 8 == 0 && (off) <= bytes && (len) <= bytes - (off))
    if (!IN(h->off_camera, 4 * np * 8) || !IN(h->off_aabb, 2 * np * 8) ||
        !IN(h->off_objects, (size_t)h->n_objects * sizeof(ndt_flat_object)) ||
        !IN(h->off_bspheres, (size_t)h->n_objects * (np + 2) * 8) ||
        !IN(h->off_geom, h->n_geom * 8) ||
        !IN(h->off_nodes, (size_t)h->n_nodes * sizeof(ndt_flat_node)) ||
        !IN(h->off_leaf_refs, (size_t)h->n_leaf_refs * 4) ||
        !IN(h->off_inf, (size_t)h->n_inf * 4) ||
        !IN(h->off_lights, (size_t)h->n_lights * sizeof(ndt_flat_light)) ||
        (h->off_view && !IN(h->off_view, ((size_t)5 * h->npad + (size_t)h->width * 4 + (size_t)h->height * 6 +
                                          (h->view_eyes ? (size_t)h->width * 2 * h->npad : 0)) * 8)))
        return ndt_set_error(NDT_B200_E_ARG, "flat scene: array outside the blob");
    if (h->cam_type < 0 || h->cam_type > 2 || h->stereo_mode < 0 || h->stereo_mode > NDT_HIDEF_3D ||
        ((h->cam_type != NDT_CAM_NORMAL || h->stereo_mode != NDT_MONO) && !h->off_view))
        return ndt_set_error(NDT_B200_E_ARG, "flat scene: camera type / stereo mode without view tables");
#undef IN
    const ndt_flat_object *ob = NDT_FLAT_PTR(blob, const ndt_flat_object, h->off_objects);
    for (int i = 0; i < h->n_objects; ++i) {
        if (ob[i].type < 0 || ob[i].type >= NDT_T_COUNT || ob[i].geom_off > h->n_geom || (ob[i].geom_off & 1u) ||
            ob[i].report_id < 0 || ob[i].report_id >= h->n_items || ob[i].n_axes < 0 || ob[i].n_axes > h->n)
            return ndt_set_error(NDT_B200_E_ARG, "flat scene: object %d malformed", i);
        if (ob[i].type == NDT_T_HCUBE &&
            (ob[i].child_begin < h->n_items || ob[i].child_count < 0 ||
             ob[i].child_begin + ob[i].child_count > h->n_objects))
            return ndt_set_error(NDT_B200_E_ARG, "flat scene: hcube %d children out of range", i);
    }
    /* every geometry block the kernels follow an offset into lies inside geom[] */
    for (int i = 0; i < h->n_objects; ++i) {
        const uint64_t nd_ = (uint64_t)flat_geom_doubles(ob[i].type, ob[i].n_axes, (int)np);
        if ((uint64_t)ob[i].geom_off + nd_ > h->n_geom)
            return ndt_set_error(NDT_B200_E_ARG, "flat scene: geometry of object %d outside geom[]", i);
    }
    const ndt_flat_light *lt = NDT_FLAT_PTR(blob, const ndt_flat_light, h->off_lights);
    for (int i = 0; i < h->n_lights; ++i) {
        if (lt[i].type < NDT_L_AMBIENT || lt[i].type > NDT_L_SPOT)
            return ndt_set_error(NDT_B200_E_ARG, "flat scene: light %d has type %d", i, lt[i].type);
        /* pos, dir, rev_unit, near_off: 4 npad doubles (ndt_flat_light::vec_off) */
        if ((uint64_t)lt[i].vec_off + 4 * np > h->n_geom)
            return ndt_set_error(NDT_B200_E_ARG, "flat scene: vectors of light %d outside geom[]", i);
    }
    /* nodes are stored in pre-order: a child's index is larger than its parent's, which rules out cycles;
     * the walk below recomputes what the traversal stack and the leaf staging are sized by */
    const ndt_flat_node *nd = NDT_FLAT_PTR(blob, const ndt_flat_node, h->off_nodes);
    int max_leaf = 0;
    for (int i = 0; i < h->n_nodes; ++i) {
        if (nd[i].dim >= h->n || nd[i].left >= h->n_nodes || nd[i].right >= h->n_nodes ||
            nd[i].left < -1 || nd[i].right < -1 || nd[i].leaf_count < 0 || nd[i].leaf_begin < 0 ||
            nd[i].leaf_begin + nd[i].leaf_count > h->n_leaf_refs ||
            (nd[i].left >= 0 && nd[i].left <= i) || (nd[i].right >= 0 && nd[i].right <= i))
            return ndt_set_error(NDT_B200_E_ARG, "flat scene: kd node %d malformed", i);
        if (nd[i].leaf_count > max_leaf) max_leaf = nd[i].leaf_count;
    }
    if (h->n_nodes > 0) {
        /* depth by one forward pass (children come after their parent) */
        int *depth = calloc((size_t)h->n_nodes, sizeof *depth);
        if (!depth) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
        int deepest = 0;
        for (int i = 0; i < h->n_nodes; ++i) {
            if (depth[i] > deepest) deepest = depth[i];
            if (nd[i].dim < 0) continue;
            if (nd[i].left >= 0 && depth[nd[i].left] < depth[i] + 1) depth[nd[i].left] = depth[i] + 1;
            if (nd[i].right >= 0 && depth[nd[i].right] < depth[i] + 1) depth[nd[i].right] = depth[i] + 1;
        }
        free(depth);
        if (deepest > h->tree_depth || max_leaf > h->max_leaf)
            return ndt_set_error(NDT_B200_E_ARG, "flat scene: header says depth %d / largest leaf %d, the nodes say %d / %d",
                                 h->tree_depth, h->max_leaf, deepest, max_leaf);
    }
    const int32_t *lr = NDT_FLAT_PTR(blob, const int32_t, h->off_leaf_refs);
    for (int i = 0; i < h->n_leaf_refs; ++i)
        if (lr[i] < 0 || lr[i] >= h->n_items) return ndt_set_error(NDT_B200_E_ARG, "flat scene: leaf ref %d out of range", i);
    const int32_t *inf = NDT_FLAT_PTR(blob, const int32_t, h->off_inf);
    for (int i = 0; i < h->n_inf; ++i)
        if (inf[i] < 0 || inf[i] >= h->n_items) return ndt_set_error(NDT_B200_E_ARG, "flat scene: inf id %d out of range", i);
    return 0;
}
