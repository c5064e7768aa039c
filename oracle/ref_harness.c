/*
 * ref_harness.c -- TEST INFRASTRUCTURE (oracle side), not product code.
 *
 * Drives the UNMODIFIED reference (compiled from /root/reference into
 * oracle/_ref/libndt_ref.so) through the same sequence main() uses for one
 * frame (ndt.c:1791-1937):
 *     scene_setup -> object_get_bounds + object_kdlist_add per object ->
 *     kd_tree_build -> scene_validate_objects -> camera_aim -> render_image
 * and exposes what the reference itself never exports:
 *   - the fp64 RGBA framebuffer (captured by ref_image_shim.c),
 *   - hit / object-id / distance buffers for PRIMARY rays, produced by calling
 *     the reference's own camera_target_point (camera.c:504) and trace_kd
 *     (object.c:683) per pixel exactly the way get_pixel_color does
 *     (ndt.c:519, 545-550),
 *   - a trace_kd call counter (only in the --wrap build, see Makefile).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "vectNd.h"
#include "image.h"
#include "object.h"
#include "scene.h"
#include "kd-tree.h"

/* symbols of the reference we call (all non-static in ndt.c) */
extern kd_tree_t kdtree;                       /* ndt.c:68 */
extern int specular_enabled;                   /* ndt.c:41 */
typedef enum { REF_MONO = 0 } ref_stereo_mode; /* ndt.c:46-48, MONO is first */
int render_image(scene *scn, char *name, char *depth_name, int width, int height,
                 int samples, int mode, int threads, int aa_diff, int aa_depth,
                 int max_optic_depth, image_t *img_copy, image_t *depth_copy);

int ref_shim_captured(const unsigned char **px, int *w, int *h, int *pw);
void ref_shim_drop(void);

typedef int (*scene_setup_fn)(scene *, int, int, int, char *);
typedef int (*scene_frames_fn)(int, char *);
typedef int (*scene_cleanup_fn)(void);

static scene g_scn;
static kd_item_list_t g_items;
static int g_frame_open = 0;
static int g_dirx_scaled = 0;
static void *g_scene_dl = NULL;
static scene_setup_fn g_setup = NULL;
static scene_frames_fn g_frames = NULL;
static scene_cleanup_fn g_cleanup = NULL;
static object **g_flat = NULL;  /* kd item order: id -> object* */
static int g_nflat = 0;
static void save_dirx(void);

/* ---- ray counter (used by the --wrap=trace_kd build only) -------------- */
#define STRIPES 64
static struct { volatile long n; char pad[56]; } g_cnt[STRIPES];
int __real_trace_kd(vectNd *, vectNd *, kd_tree_t *, vectNd *, vectNd *, object **, double);
int __wrap_trace_kd(vectNd *p, vectNd *l, kd_tree_t *kd, vectNd *h, vectNd *n, object **o, double lim)
{
    unsigned long id = (unsigned long)pthread_self();
    __atomic_fetch_add(&g_cnt[(id >> 12) % STRIPES].n, 1, __ATOMIC_RELAXED);
    return __real_trace_kd(p, l, kd, h, n, o, lim);
}
long refh_ray_count(int reset)
{
    long t = 0;
    for (int i = 0; i < STRIPES; ++i) {
        t += g_cnt[i].n;
        if (reset)
            g_cnt[i].n = 0;
    }
    return t;
}

static int g_quiet_fd = -1;
static void hush(int on)
{
    /* the reference prints progress to stdout; keep test logs readable */
    static int saved = -1;
    if (on) {
        fflush(stdout);
        saved = dup(1);
        if (g_quiet_fd < 0)
            g_quiet_fd = open("/dev/null", 1);
        dup2(g_quiet_fd, 1);
    } else if (saved >= 0) {
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
        saved = -1;
    }
}

int refh_init(const char *objects_dir)
{
    static int done = 0;
    if (done)
        return 0;
    hush(1);
    int r = register_objects((char *)objects_dir);
    hush(0);
    done = 1;
    return r;
}

int refh_open_scene(const char *scene_so)
{
    if (g_scene_dl) {
        if (g_cleanup)
            g_cleanup();
        dlclose(g_scene_dl);
        g_scene_dl = NULL;
    }
    g_setup = NULL; g_frames = NULL; g_cleanup = NULL;
    if (scene_so == NULL || scene_so[0] == '\0') {
        g_setup = scene_setup;  /* built-in scene, scene.c:429 */
        return 0;
    }
    g_scene_dl = dlopen(scene_so, RTLD_NOW);
    if (!g_scene_dl) {
        fprintf(stderr, "refh_open_scene: %s\n", dlerror());
        return -1;
    }
    g_setup = (scene_setup_fn)dlsym(g_scene_dl, "scene_setup");
    g_frames = (scene_frames_fn)dlsym(g_scene_dl, "scene_frames");
    g_cleanup = (scene_cleanup_fn)dlsym(g_scene_dl, "scene_cleanup");
    return g_setup ? 0 : -2;
}

int refh_scene_frames(int dims, const char *cfg)
{
    return g_frames ? g_frames(dims, (char *)cfg) : -1;
}

static void flat_add(object *o)
{
    char tn[OBJ_TYPE_MAX_LEN] = "";
    o->type_name(tn, sizeof tn);
    if (!strcmp(tn, "cluster")) { /* object_kdlist_add recursion, object.c:636-643 */
        for (int i = 0; i < o->n_obj; ++i)
            flat_add(o->obj[i]);
        return;
    }
    g_flat = realloc(g_flat, (g_nflat + 1) * sizeof *g_flat);
    g_flat[g_nflat++] = o;
}

/* run scene_setup for a frame that will not be rendered (stateful scenes,
 * ndt.c:1816-1825) */
int refh_skip_frame(int dims, int frame, int frames, const char *cfg)
{
    hush(1);
    g_setup(&g_scn, dims, frame, frames, (char *)cfg);
    scene_free(&g_scn);
    hush(0);
    return 0;
}

/* `ndt -y` (ndt.c:1798-1808): scene_setup for the frame, then the reference's own scene_write_yaml
 * (scene.c:1000) -- before kd build and camera_aim, so nothing "prepared" is dumped.  to_buffer != 0 goes
 * through scene_write_yaml_buffer (scene.c:1045, the MPI scene transport) and then writes that text. */
int refh_write_yaml(int dims, int frame, int frames, const char *cfg, const char *fname, int to_buffer)
{
    int r;
    if (g_frame_open)
        return -1;
    hush(1);
    g_setup(&g_scn, dims, frame, frames, (char *)cfg);
    if (to_buffer) {
        unsigned char *buf = NULL;
        size_t len = 0;
        r = scene_write_yaml_buffer(&g_scn, &buf, &len);
        FILE *fp = fopen(fname, "wb");
        if (!fp || fwrite(buf, 1, len, fp) != len) r = -2;
        if (fp) fclose(fp);
        free(buf);
    } else
        r = scene_write_yaml(&g_scn, (char *)fname);
    scene_free(&g_scn);
    hush(0);
    return r;
}

static int g_skip_kd_build = 0;
/* the same without kd_tree_build: the global kd-tree is initialised and left empty, to be built by
 * ndt_b200_kd_tree_build_bounded (scenes the reference's exhaustive builder cannot finish, SURVEY note 8) */
int refh_begin_frame_nokd(int dims, int frame, int frames, const char *cfg)
{
    g_skip_kd_build = 1;
    int r = refh_begin_frame(dims, frame, frames, cfg, NULL);
    g_skip_kd_build = 0;
    return r;
}

/* trace() (object.c:692) over ALL kd items, no mask, no tree: the brute-force answer the kd result is
 * checked against.  Returns trace()'s return value; *obj_id = kd item index of the reported object or -1 */
int refh_trace_brute(const double *o, const double *v, double dist_limit, double *hit_out, int *obj_id)
{
    if (!g_frame_open)
        return -1;
    int dim = g_scn.dimensions;
    vectNd pos, look, hp, hn;
    vectNd_calloc(&pos, dim); vectNd_calloc(&look, dim); vectNd_calloc(&hp, dim); vectNd_calloc(&hn, dim);
    for (int i = 0; i < dim; ++i) { vectNd_set(&pos, i, o[i]); vectNd_set(&look, i, v[i]); }
    object *ptr = NULL;
    int r = trace(&pos, &look, g_flat, NULL, g_nflat, NULL, &hp, &hn, &ptr, NULL, dist_limit);
    *obj_id = -1;
    if (ptr)
        for (int i = 0; i < g_nflat; ++i)
            if (g_flat[i] == ptr) { *obj_id = i; break; }
    if (hit_out)
        for (int i = 0; i < dim; ++i) hit_out[i] = hp.v[i];
    vectNd_free(&pos); vectNd_free(&look); vectNd_free(&hp); vectNd_free(&hn);
    return r;
}

int refh_begin_frame(int dims, int frame, int frames, const char *cfg, double *kd_seconds)
{
    if (g_frame_open)
        return -1;
    hush(1);
    g_setup(&g_scn, dims, frame, frames, (char *)cfg);

    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    kd_tree_init(&kdtree, g_scn.dimensions);
    kd_item_list_init(&g_items);
    g_nflat = 0;
    for (int i = 0; i < g_scn.num_objects; ++i) {
        object *o = g_scn.object_ptrs[i];
        object_get_bounds(o);
        object_kdlist_add(&g_items, o, i);
        flat_add(o);
    }
    if (!g_skip_kd_build)
        kd_tree_build(&kdtree, &g_items);
    clock_gettime(CLOCK_MONOTONIC, &b);
    if (kd_seconds)
        *kd_seconds = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);

    scene_validate_objects(&g_scn);
    camera_aim(&g_scn.cam);
    /* The plugins prepare themselves lazily on their first intersect() behind a double-checked lock whose
     * flag is a bit-field next to `transparent` (object.h:24, e.g. orthotope.c:23-54,152): with many render
     * threads the first rays race on it and the reference occasionally dereferences a NULL `prepped`
     * (seen as a SIGSEGV of the test process on the 16-core GPU box).  One ray per object from this thread
     * runs every prepare() up front; the prepared values do not depend on the ray. */
    {
        vectNd o, v, r, nrm;
        vectNd_calloc(&o, g_scn.dimensions);
        vectNd_calloc(&v, g_scn.dimensions);
        vectNd_calloc(&r, g_scn.dimensions);
        vectNd_calloc(&nrm, g_scn.dimensions);
        vectNd_set(&v, 0, 1.0);
        for (int i = 0; i < g_nflat; ++i) {
            object *hit = NULL;
            g_flat[i]->intersect(g_flat[i], &o, &v, &r, &nrm, &hit);
        }
        vectNd_free(&o); vectNd_free(&v); vectNd_free(&r); vectNd_free(&nrm);
    }
    hush(0);
    g_frame_open = 1;
    g_dirx_scaled = 0;
    save_dirx();
    return 0;
}

/* render_image rescales cam.dirX in place on every call (ndt.c:926) and nothing
 * else of the scene; restoring the aimed dirX lets a frame be rendered again
 * (timing repeats) without rebuilding the scene and the kd-tree */
static double g_dirx_saved[64];
static void save_dirx(void)
{
    int k = g_scn.cam.dirX.n + (g_scn.cam.dirX.n & 1);
    for (int i = 0; i < k && i < 64; ++i) g_dirx_saved[i] = g_scn.cam.dirX.v[i];
}
int refh_reaim(void)
{
    if (!g_frame_open)
        return -1;
    int k = g_scn.cam.dirX.n + (g_scn.cam.dirX.n & 1);
    for (int i = 0; i < k && i < 64; ++i) g_scn.cam.dirX.v[i] = g_dirx_saved[i];
    g_dirx_scaled = 0;
    return 0;
}

void *refh_scene(void) { return &g_scn; }
void *refh_kdtree(void) { return &kdtree; }
int refh_num_items(void) { return g_nflat; }
void *refh_object_get_bounds_ptr(void) { return (void *)object_get_bounds; }
void refh_set_specular(int on) { specular_enabled = on; }

/* render_image with the global recursive_aa set (-w / -a, ndt.c:1455-1466): returns the 8-bit
 * actual_img (ndt.c:940-942) the reference saves last */
extern int recursive_aa;
int refh_render_aa(int w, int h, int threads, int max_optic_depth, int aa_diff, int aa_depth,
                   unsigned char *rgba_u8, double *seconds)
{
    if (!g_frame_open)
        return -1;
    if (g_dirx_scaled)
        return -2;
    struct timespec a, b;
    hush(1);
    clock_gettime(CLOCK_MONOTONIC, &a);
    recursive_aa = 1;
    render_image(&g_scn, "oracle", NULL, w, h, 1, REF_MONO, threads, aa_diff, aa_depth, max_optic_depth, NULL, NULL);
    recursive_aa = 0;
    clock_gettime(CLOCK_MONOTONIC, &b);
    hush(0);
    g_dirx_scaled = 1;
    if (seconds)
        *seconds = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
    const unsigned char *px; int cw, ch, pw;
    if (!ref_shim_captured(&px, &cw, &ch, &pw) || cw != w || ch != h || pw != 4)
        return -3;
    if (rgba_u8)
        memcpy(rgba_u8, px, (size_t)w * h * 4);
    ref_shim_drop();
    return 0;
}

void *refh_rotate2_ptr(void) { return (void *)vectNd_rotate2; }

/* what main() does for -V / -P before camera_aim (ndt.c:1915-1925): camera type and fields of view */
int refh_set_camera(int type, double h_fov, double v_fov)
{
    if (!g_frame_open)
        return -1;
    refh_reaim();
    g_scn.cam.type = type;
    if (type != CAMERA_NORMAL) {
        g_scn.cam.hFov = h_fov;
        g_scn.cam.vFov = v_fov;
    }
    hush(1);
    camera_aim(&g_scn.cam);
    hush(0);
    g_dirx_scaled = 0;
    save_dirx();
    return 0;
}

int refh_render_ex(int w, int h, int threads, int max_optic_depth, int stereo, double *rgba, double *seconds);

/* render_image (ndt.c:900) with samples=1, MONO; copies the fp64 RGBA frame */
int refh_render(int w, int h, int threads, int max_optic_depth, double *rgba, double *seconds)
{
    return refh_render_ex(w, h, threads, max_optic_depth, REF_MONO, rgba, seconds);
}

/* the same for any stereo_mode of ndt.c:46-48 */
int refh_render_ex(int w, int h, int threads, int max_optic_depth, int stereo, double *rgba, double *seconds)
{
    if (!g_frame_open)
        return -1;
    if (g_dirx_scaled) /* render_image rescales cam.dirX on every call (ndt.c:926) */
        return -2;
    struct timespec a, b;
    hush(1);
    clock_gettime(CLOCK_MONOTONIC, &a);
    render_image(&g_scn, "oracle", NULL, w, h, 1, (ref_stereo_mode)stereo, threads, 20, 4, max_optic_depth, NULL, NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    hush(0);
    g_dirx_scaled = 1;
    if (seconds)
        *seconds = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
    const unsigned char *px; int cw, ch, pw;
    if (!ref_shim_captured(&px, &cw, &ch, &pw) || cw != w || ch != h || pw != 32)
        return -3;
    if (rgba)
        memcpy(rgba, px, (size_t)w * h * 32);
    ref_shim_drop();
    return 0;
}

/* primary-ray hit / id / distance buffers, reference code only */
int refh_primary(int w, int h, unsigned char *hit, int *obj_id, double *dist)
{
    if (!g_frame_open)
        return -1;
    int dim = g_scn.dimensions;
    if (!g_dirx_scaled) { /* same mutation render_image applies, ndt.c:925-926 */
        vectNd_scale(&g_scn.cam.dirX, w / (double)h, &g_scn.cam.dirX);
        g_dirx_scaled = 1;
    }
    vectNd pixel, look, hp, hn;
    vectNd_alloc(&pixel, dim);
    vectNd_alloc(&look, dim);
    for (int j = 0; j < h; ++j) {
        for (int i = 0; i < w; ++i) {
            double x = (double)i / (double)w - 0.5;      /* ndt.c:632 */
            double y = -((double)j / (double)h - 0.5);   /* ndt.c:633 */
            camera_target_point(&g_scn.cam, x, y, g_scn.cam.focal_distance, &pixel);
            vectNd_sub(&pixel, &g_scn.cam.pos, &look);
            vectNd_unitize(&look);
            vectNd_calloc(&hp, dim);
            vectNd_calloc(&hn, dim);
            object *op = NULL;
            trace_kd(&g_scn.cam.pos, &look, &kdtree, &hp, &hn, &op, -1.0);
            double d = -1.0;
            int id = -1;
            if (op != NULL) {
                vectNd_dist(&hp, &g_scn.cam.pos, &d);
                for (int k = 0; k < g_nflat; ++k)
                    if (g_flat[k] == op) { id = k; break; }
            }
            /* get_ray_color shades only when obj && dist > EPSILON (ndt.c:376) */
            size_t p = (size_t)j * w + i;
            hit[p] = (op != NULL && d > EPSILON) ? 1 : 0;
            obj_id[p] = id;
            if (dist)
                dist[p] = d;
            vectNd_free(&hp);
            vectNd_free(&hn);
        }
    }
    vectNd_free(&pixel);
    vectNd_free(&look);
    return 0;
}

/* single-ray probe used by the per-primitive known-answer tests */
int refh_trace_ray(const double *o, const double *v, double dist_limit,
                   double *hit_out, double *normal_out, int *obj_id)
{
    int dim = g_scn.dimensions;
    vectNd vo, vv, hp, hn;
    vectNd_calloc(&vo, dim); vectNd_calloc(&vv, dim);
    vectNd_calloc(&hp, dim); vectNd_calloc(&hn, dim);
    for (int i = 0; i < dim; ++i) { vo.v[i] = o[i]; vv.v[i] = v[i]; }
    object *op = NULL;
    int r = trace_kd(&vo, &vv, &kdtree, &hp, &hn, &op, dist_limit);
    *obj_id = -1;
    if (op)
        for (int k = 0; k < g_nflat; ++k)
            if (g_flat[k] == op) { *obj_id = k; break; }
    for (int i = 0; i < dim; ++i) { hit_out[i] = hp.v[i]; normal_out[i] = hn.v[i]; }
    vectNd_free(&vo); vectNd_free(&vv); vectNd_free(&hp); vectNd_free(&hn);
    return r;
}

/* a second kd-tree over the SAME item list, built by an external builder with
 * kd_tree_build's signature (used to check ndt_b200_kd_tree_build node for node) */
static kd_tree_t g_alt;
static int g_alt_open = 0;
void *refh_items(void) { return &g_items; }
void *refh_alt_tree_begin(void)
{
    if (g_alt_open) { hush(1); kd_tree_free(&g_alt); hush(0); }
    kd_tree_init(&g_alt, g_scn.dimensions);
    g_alt_open = 1;
    return &g_alt;
}
void refh_alt_tree_end(void)
{
    if (g_alt_open) { hush(1); kd_tree_free(&g_alt); hush(0); g_alt_open = 0; }
}
/* the reference's own builder on the alternate tree, for timing */
double refh_alt_tree_build_reference(void)
{
    struct timespec a, b;
    hush(1);
    clock_gettime(CLOCK_MONOTONIC, &a);
    kd_tree_build(&g_alt, &g_items);
    clock_gettime(CLOCK_MONOTONIC, &b);
    hush(0);
    return (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
}

int refh_end_frame(void)
{
    refh_alt_tree_end();
    if (!g_frame_open)
        return -1;
    hush(1);
    kd_item_list_free(&g_items, 1);
    kd_tree_free(&kdtree);
    scene_free(&g_scn);
    hush(0);
    g_frame_open = 0;
    return 0;
}
