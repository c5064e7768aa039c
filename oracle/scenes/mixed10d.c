/*
 * mixed10d.c -- TEST INFRASTRUCTURE: an ndt scene plugin written for this repo
 * against the reference's public scene API (README.md:71-125 of the
 * reference: scene_frames / scene_setup / scene_cleanup).  It is the C twin
 * of BASELINE config 5 ("YAML scene mixing hplane, hcylinder, hdisk and hfacet
 * in 10-D with multiple lights and shadow rays"); the reference ships no such
 * YAML and libyaml is absent here, so the same object_add_* calls are made
 * from C.  Works for any dims >= 4; objects live in the first 3 axes with
 * small offsets in the higher ones.
 */
#include <math.h>
#include <stdio.h>
#include "vectNd.h"
#include "scene.h"

int scene_frames(int dims, char *cfg) { (void)dims; (void)cfg; return 24; }
int scene_cleanup(void) { return 0; }

static void setv(vectNd *v, int dims, double x, double y, double z, double w, double rest)
{
    vectNd_reset(v);
    vectNd_set(v, 0, x); vectNd_set(v, 1, y); vectNd_set(v, 2, z);
    if (dims > 3) vectNd_set(v, 3, w);
    for (int i = 4; i < dims; ++i) vectNd_set(v, i, rest * (1 + (i & 1)));
}

static void paint(object *o, double r, double g, double b, double refl)
{
    o->red = r; o->green = g; o->blue = b;
    o->red_r = o->green_r = o->blue_r = refl;
}

int scene_setup(scene *scn, int dims, int frame, int frames, char *cfg)
{
    (void)cfg;
    double t = frames > 0 ? frame / (double)frames : 0.0;
    object *o = NULL;
    light *l = NULL;
    vectNd p, q;
    scene_init(scn, "mixed10d", dims);
    vectNd_calloc(&p, dims);
    vectNd_calloc(&q, dims);

    /* ground: hplane y = -6, mildly reflective */
    scene_alloc_object(scn, dims, &o, "hplane");
    setv(&p, dims, 0, -6, 0, 0, 0); object_add_pos(o, &p);
    setv(&p, dims, 0, 1, 0, 0, 0);  object_add_dir(o, &p);
    paint(o, 0.8, 0.8, 0.85, 0.35);

    /* two finite hcylinders (flag[0]==0): bottom + dims-2 orthogonal "tops" */
    for (int c = 0; c < 2; ++c) {
        scene_alloc_object(scn, dims, &o, "hcylinder");
        double bx = c ? 9.0 : -9.0;
        setv(&p, dims, bx, -6, 18 + 4 * c, 0, 0);
        object_add_pos(o, &p);
        /* axes: every coordinate except 0 and 2 (the circular cross-section) */
        for (int a = 1; a < dims; ++a) {
            if (a == 2) continue;
            vectNd_copy(&q, &p);
            double len = (a == 1) ? 12.0 : 6.0 + a;
            double cur; vectNd_get(&q, a, &cur);
            vectNd_set(&q, a, cur + len);
            object_add_pos(o, &q);
        }
        object_add_size(o, 2.5 - 0.5 * c);
        object_add_flag(o, 0);
        paint(o, c ? 0.2 : 0.9, 0.6, c ? 0.9 : 0.2, c ? 0.0 : 0.25);
        if (c) { o->transparent = 1; o->refract_index = 1.5; o->red_r = o->green_r = o->blue_r = 0.1; }
    }

    /* two hdisks facing the camera-ish */
    for (int d = 0; d < 2; ++d) {
        scene_alloc_object(scn, dims, &o, "hdisk");
        setv(&p, dims, d ? 4.0 : -3.0, d ? 3.0 : -1.0, 26 - 6 * d, 0.5 * d, 0.0);
        object_add_pos(o, &p);
        setv(&p, dims, d ? -0.3 : 0.2, 0.4, -1.0, 0.1, 0.0);
        object_add_dir(o, &p);
        object_add_size(o, 4.0 + d);
        paint(o, d ? 0.95 : 0.3, 0.5, d ? 0.2 : 0.8, d ? 0.5 : 0.0);
    }

    /* four hfacets forming a tilted fan; flag 0 = geometric normal */
    for (int f = 0; f < 4; ++f) {
        scene_alloc_object(scn, dims, &o, "hfacet");
        double a0 = 0.5 * f + 2 * M_PI * t, a1 = a0 + 0.5;
        setv(&p, dims, 0, 6, 14, 0, 0.05 * f); object_add_pos(o, &p);
        setv(&p, dims, 7 * cos(a0), -2 + f, 14 + 7 * sin(a0), 0.3, 0.0); object_add_pos(o, &p);
        setv(&p, dims, 7 * cos(a1), -2 + f, 14 + 7 * sin(a1), -0.3, 0.0); object_add_pos(o, &p);
        setv(&p, dims, 0, 1, 0, 0, 0);
        object_add_dir(o, &p); object_add_dir(o, &p); object_add_dir(o, &p);
        object_add_flag(o, 0);
        paint(o, 0.3 + 0.2 * f, 0.9 - 0.2 * f, 0.4, f == 2 ? 0.3 : 0.0);
    }

    /* a sphere and an orthotope so every bounded family casts shadows here */
    scene_alloc_object(scn, dims, &o, "sphere");
    setv(&p, dims, 0, -2, 20, 0, 0); object_add_pos(o, &p);
    object_add_size(o, 3.0);
    paint(o, 0.9, 0.2, 0.2, 0.4);

    camera_init(&scn->cam);
    vectNd vp, vt, up;
    vectNd_calloc(&vp, dims); vectNd_calloc(&vt, dims); vectNd_calloc(&up, dims);
    setv(&vp, dims, 30 * sin(2 * M_PI * t), 14, -22 + 6 * cos(2 * M_PI * t), 2, 0.25);
    setv(&vt, dims, 0, 0, 18, 0, 0);
    setv(&up, dims, 0, 10, 0, 0, 0);
    camera_set_aim(&scn->cam, &vp, &vt, &up, 0.0);

    vectNd_calloc(&scn->ambient.pos, dims);
    scn->ambient.red = scn->ambient.green = scn->ambient.blue = 0.2;

    const double lp[3][4] = { {0, 25, 5, 0}, {-20, 12, 30, 2}, {18, 9, -4, -3} };
    const double li[3] = { 260, 180, 140 };
    for (int k = 0; k < 3; ++k) {
        scene_alloc_light(scn, &l);
        l->type = LIGHT_POINT;
        vectNd_calloc(&l->pos, dims);
        setv(&l->pos, dims, lp[k][0], lp[k][1], lp[k][2], lp[k][3], 0.1 * k);
        l->red = li[k]; l->green = li[k] * 0.95; l->blue = li[k] * 0.9;
    }
    scene_alloc_light(scn, &l);
    l->type = LIGHT_DIRECTIONAL;
    vectNd_calloc(&l->dir, dims);
    setv(&l->dir, dims, -0.4, -1.0, 0.3, 0.05, 0.0);
    l->red = l->green = l->blue = 0.3;

    scene_alloc_light(scn, &l);
    l->type = LIGHT_SPOT;
    vectNd_calloc(&l->pos, dims);
    vectNd_calloc(&l->dir, dims);
    setv(&l->pos, dims, 5, 30, 20, 0, 0);
    setv(&l->dir, dims, 0, -1, 0, 0, 0);
    l->angle = 25.0;
    l->red = 300; l->green = 280; l->blue = 200;

    vectNd_free(&p); vectNd_free(&q);
    return 1;
}
