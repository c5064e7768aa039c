/*
 * render_image_b200.c -- the binding INTEGRATION.md describes, as real code.
 *
 * It defines render_image() with the reference's exact signature (ndt.c:900)
 * and forwards to ndt_b200_render_image().  Linked into an executable together
 * with the UNMODIFIED reference (oracle/_ref/libndt_ref.so, whose main() is
 * exported as ndt_ref_main), this definition pre-empts the library's own
 * render_image -- the call at ndt.c:1933 goes through the PLT -- so the stock
 * ndt command line (getopt, scene plugins, kd build, camera aim, frame loop)
 * runs unchanged and only the frame is rendered by the GPU:
 *
 *     integration/ndt_b200_demo -d 4 -f 0 -r 640x360 -o oracle/_ref/objects
 *
 * kd_tree_build (kd-tree.c:421, called at ndt.c:1908) is pre-empted the same way and forwards to
 * ndt_b200_kd_tree_build: the exhaustive plane search of the reference's builder runs on the GPU and
 * returns the reference's own kd_tree_t, node for node (13 s -> 17 ms per frame of BASELINE config 2).
 * NDT_B200_HOST_KD=1 keeps the reference's builder.
 *
 * Image codecs stay on the host; the reference build used here has none
 * (png/jpeg headers are absent), so the frame is written as binary PPM next to
 * the name ndt chose.  Test infrastructure only: it needs oracle/_ref.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <dlfcn.h>
#include "ndt_b200.h"
#include "ndt_abi.h"

extern char kdtree[];                    /* kd_tree_t kdtree, ndt.c:68 */
extern int specular_enabled;             /* ndt.c:41 */
extern int recursive_aa;                 /* ndt.c:44 (-w / -a) */
int object_get_bounds(void *obj);        /* object.c:582 */
int vectNd_rotate2(void *v, void *center, void *v1, void *v2, double angle, void *res);   /* vectNd.c:271 */
int ndt_ref_main(int argc, char **argv); /* ndt.c:1390 compiled with -Dmain=ndt_ref_main */

/* kd_tree_build(kd_tree_t*, kd_item_list_t*), kd-tree.c:421: same arguments, same tree */
int kd_tree_build(void *tree, void *items)
{
    const char *e = getenv("NDT_B200_HOST_KD");
    if (e && *e && *e != '0') {
        int (*ref)(void *, void *) = (int (*)(void *, void *))dlsym(RTLD_NEXT, "kd_tree_build");
        if (ref) return ref(tree, items);
    }
    int rc = ndt_b200_kd_tree_build(tree, items);
    if (rc < 0) {
        fprintf(stderr, "ndt_b200: %s\n", ndt_b200_last_error());
        exit(1);
    }
    return rc;
}

static unsigned char d2c(double d)       /* image.h:36-39 */
{
    double c = d < 1.0 ? d : 1.0;
    c = c > 0.0 ? c : 0.0;
    return (unsigned char)(sqrt(c) * 255);
}

int render_image(void *scn, char *name, char *depth_name, int width, int height, int samples,
                 int mode, int threads, int aa_diff, int aa_depth, int max_optic_depth,
                 void *img_copy, void *depth_copy)
{
    ndt_b200_host_api host = { object_get_bounds, vectNd_rotate2 };
    ndtabi_image local;
    memset(&local, 0, sizeof local);
    ndtabi_image *img = img_copy ? (ndtabi_image *)img_copy : &local;
    int rc = (recursive_aa ? ndt_b200_render_image_aa : ndt_b200_render_image)(
        scn, kdtree, &host, name, depth_name, width, height, samples, mode,
        threads, aa_diff, aa_depth, max_optic_depth, specular_enabled, img, depth_copy);
    if (rc < 0) {
        fprintf(stderr, "ndt_b200: %s\n", ndt_b200_last_error());
        exit(1);                         /* the host application decides; there is no CPU fallback */
    }
    if (name) {
        char path[4096];
        snprintf(path, sizeof path, "%s", name);
        char *dot = strrchr(path, '.');
        if (dot && strlen(dot) <= 8) *dot = '\0';
        strncat(path, ".ppm", sizeof path - strlen(path) - 1);
        FILE *f = fopen(path, "wb");
        if (f) {
            fprintf(f, "P6\n%d %d\n255\n", width, height);
            unsigned char *row = malloc((size_t)width * 3);
            for (int y = 0; row && y < height; ++y) {
                if (img->pixel_width == 4) {         /* the anti-aliased frame is already 8-bit RGBA */
                    const unsigned char *px = img->pixels + (size_t)y * width * 4;
                    for (int x = 0; x < width; ++x) { row[3 * x] = px[4 * x]; row[3 * x + 1] = px[4 * x + 1]; row[3 * x + 2] = px[4 * x + 2]; }
                } else {
                    const double *px = (const double *)img->pixels + (size_t)y * width * 4;
                    for (int x = 0; x < width; ++x) { row[3 * x] = d2c(px[4 * x]); row[3 * x + 1] = d2c(px[4 * x + 1]); row[3 * x + 2] = d2c(px[4 * x + 2]); }
                }
                fwrite(row, 3, (size_t)width, f);
            }
            free(row);
            fclose(f);
            printf("\tndt_b200: wrote %s\n", path);
        }
    }
    if (img == &local) free(local.pixels);
    return 1;
}

int main(int argc, char **argv) { return ndt_ref_main(argc, argv); }
