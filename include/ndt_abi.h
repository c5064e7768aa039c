/*
 * ndt_abi.h -- binary layout of the ndt host structures the drop-in reads.
 *
 * libndt_b200 is linked into (or loaded by) an ndt host whose scene/object
 * plugins poke these structures directly, so the layout is an ABI.  The
 * declarations below are OUR mirror of that ABI for x86-64 SysV/gcc, written
 * from the offsets measured in SURVEY.md section 8b; every offset the flattener
 * relies on is pinned with a static assertion, and tests/test_abi.py
 * recompiles the assertions against the reference's own headers when they
 * are available.  Field names carry an `a_` prefix-free but distinct naming
 * so that this header can be included next to the reference's headers.
 *
 * reference: vectNd.h:42-51, bounding.h:12-19, object.h:23-74, scene.h:36-62,
 *            camera.h:32-75, kd-tree.h:16-18,52-74, image.h:81-89
 */
#ifndef NDT_ABI_H
#define NDT_ABI_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* vectNd.h:42-51 : inline storage for n<=4, heap otherwise; v==space when inline */
typedef struct ndtabi_vec {
    double inl[4];
    double *v;
    int n;
} __attribute__((aligned(16))) ndtabi_vec;

/* bounding.h:12-19 */
typedef struct ndtabi_bsphere {
    ndtabi_vec center;
    double radius;          /* 0 = not computed yet, <0 = infinite object */
    unsigned int prepared : 1;
    double radius_sqr;
} ndtabi_bsphere;

struct ndtabi_object;
typedef int (*ndtabi_type_name_fn)(char *name, int size);
typedef int (*ndtabi_intersect_fn)(struct ndtabi_object *, ndtabi_vec *o, ndtabi_vec *v,
                                   ndtabi_vec *res, ndtabi_vec *normal, struct ndtabi_object **);
typedef int (*ndtabi_color_fn)(struct ndtabi_object *, ndtabi_vec *at, double *, double *, double *);

/* object.h:23-74 */
typedef struct ndtabi_object {
    unsigned int transparent : 1;
    unsigned int prepared : 1;
    int dimensions;
    double rgb[3];
    double refl[3];
    double refract_index;
    char name[32];
    ndtabi_vec *pos;  int n_pos,  cap_pos;
    ndtabi_vec *dir;  int n_dir,  cap_dir;
    double *size;     int n_size, cap_size;
    int *flag;        int n_flag, cap_flag;
    struct ndtabi_object **obj; int n_obj, cap_obj;
    ndtabi_bsphere bounds;
    void *prepped;
    void *dl_handle;
    ndtabi_type_name_fn type_name;
    void *params;
    void *cleanup;
    void *bounding_points;
    ndtabi_intersect_fn intersect;
    ndtabi_color_fn get_color;
    ndtabi_color_fn get_reflect;
    void *get_trans;
    void *refract_ray;
} ndtabi_object;

/* scene.h:17-23 enum order */
enum { NDTABI_LIGHT_AMBIENT = 0, NDTABI_LIGHT_POINT, NDTABI_LIGHT_DIRECTIONAL,
       NDTABI_LIGHT_SPOT, NDTABI_LIGHT_DISK, NDTABI_LIGHT_RECT };

/* scene.h:36-49 */
typedef struct ndtabi_light {
    ndtabi_vec pos, target, dir, u, v;
    double radius;
    int type;
    double rgb[3];
    double angle;
    ndtabi_vec u1, v1;
    unsigned int prepared : 1;
    char name[32];
} ndtabi_light;

/* camera.h:17-20 enum order */
enum { NDTABI_CAMERA_NORMAL = 0, NDTABI_CAMERA_VR, NDTABI_CAMERA_PANO };

/* camera.h:32-75 */
typedef struct ndtabi_camera {
    int type;
    ndtabi_vec viewPoint, viewTarget, up;
    double rotation, eye_offset;
    double aperture_radius, focal_distance;
    double zoom;
    unsigned int flip_x : 1, flip_y : 1, flatten : 1;
    double hFov, vFov;
    unsigned int prepared : 1;
    double leveling;
    ndtabi_vec pos, leftEye, rightEye;
    ndtabi_vec dirX, dirY, imgOrig;
    ndtabi_vec localX, localY, localZ;
} ndtabi_camera;

/* scene.h:51-62 */
typedef struct ndtabi_scene {
    int dimensions;
    ndtabi_camera cam;
    int num_objects;
    int num_lights;
    ndtabi_object **object_ptrs;
    ndtabi_light **lights;
    ndtabi_light ambient;
    double bg[4];
    char name[64];
} ndtabi_scene;

/* kd-tree.h:52-59 : leaves have dim<0 and num>0; inner nodes num==0 */
typedef struct ndtabi_kd_node {
    int dim;
    double boundary;
    int num;
    int *obj_ids;
    void **objs;
    struct ndtabi_kd_node *left, *right;
} ndtabi_kd_node;

/* kd-tree.h:16-18, 66-74 */
typedef struct ndtabi_kd_tree {
    ndtabi_vec bb_lower, bb_upper;
    void **obj_ptrs;
    void **inf_obj_ptrs;
    int *ids;
    int obj_num;       /* all items, finite and infinite (kd-tree.c:470) */
    int inf_obj_num;
    ndtabi_kd_node *root;
} ndtabi_kd_tree;

/* image.h:81-89 */
typedef struct ndtabi_image {
    int width, height;
    int pixel_width;     /* 4 = u8 RGBA, 32 = fp64 RGBA */
    int allocated;
    int type;
    int edge_style;
    unsigned char *pixels;
} ndtabi_image;

#define NDTABI_ASSERT(c) _Static_assert(c, #c)
#ifndef __cplusplus
NDTABI_ASSERT(sizeof(ndtabi_vec) == 48 && offsetof(ndtabi_vec, v) == 32 && offsetof(ndtabi_vec, n) == 40);
NDTABI_ASSERT(sizeof(ndtabi_bsphere) == 80 && offsetof(ndtabi_bsphere, radius) == 48 &&
              offsetof(ndtabi_bsphere, radius_sqr) == 64);
NDTABI_ASSERT(sizeof(ndtabi_object) == 352);
NDTABI_ASSERT(offsetof(ndtabi_object, dimensions) == 4 && offsetof(ndtabi_object, rgb) == 8 &&
              offsetof(ndtabi_object, refl) == 32 && offsetof(ndtabi_object, refract_index) == 56);
NDTABI_ASSERT(offsetof(ndtabi_object, pos) == 96 && offsetof(ndtabi_object, n_pos) == 104 &&
              offsetof(ndtabi_object, dir) == 112 && offsetof(ndtabi_object, n_dir) == 120 &&
              offsetof(ndtabi_object, size) == 128 && offsetof(ndtabi_object, n_size) == 136 &&
              offsetof(ndtabi_object, flag) == 144 && offsetof(ndtabi_object, n_flag) == 152 &&
              offsetof(ndtabi_object, obj) == 160 && offsetof(ndtabi_object, n_obj) == 168);
NDTABI_ASSERT(offsetof(ndtabi_object, bounds) == 176 && offsetof(ndtabi_object, prepped) == 256 &&
              offsetof(ndtabi_object, type_name) == 272 && offsetof(ndtabi_object, intersect) == 304 &&
              offsetof(ndtabi_object, get_color) == 312 && offsetof(ndtabi_object, get_reflect) == 320 &&
              offsetof(ndtabi_object, refract_ray) == 336);
NDTABI_ASSERT(sizeof(ndtabi_light) == 432 && offsetof(ndtabi_light, dir) == 96 &&
              offsetof(ndtabi_light, radius) == 240 && offsetof(ndtabi_light, type) == 248 &&
              offsetof(ndtabi_light, rgb) == 256 && offsetof(ndtabi_light, angle) == 280 &&
              offsetof(ndtabi_light, u1) == 288);
NDTABI_ASSERT(sizeof(ndtabi_camera) == 672 && offsetof(ndtabi_camera, viewPoint) == 16 &&
              offsetof(ndtabi_camera, focal_distance) == 184 && offsetof(ndtabi_camera, zoom) == 192 &&
              offsetof(ndtabi_camera, hFov) == 208 && offsetof(ndtabi_camera, leveling) == 232 &&
              offsetof(ndtabi_camera, pos) == 240 && offsetof(ndtabi_camera, dirX) == 384 &&
              offsetof(ndtabi_camera, dirY) == 432 && offsetof(ndtabi_camera, imgOrig) == 480 &&
              offsetof(ndtabi_camera, localZ) == 624);
NDTABI_ASSERT(sizeof(ndtabi_scene) == 1248 && offsetof(ndtabi_scene, cam) == 16 &&
              offsetof(ndtabi_scene, num_objects) == 688 && offsetof(ndtabi_scene, num_lights) == 692 &&
              offsetof(ndtabi_scene, object_ptrs) == 696 && offsetof(ndtabi_scene, lights) == 704 &&
              offsetof(ndtabi_scene, ambient) == 720 && offsetof(ndtabi_scene, bg) == 1152 &&
              offsetof(ndtabi_scene, name) == 1184);
NDTABI_ASSERT(sizeof(ndtabi_kd_node) == 56 && offsetof(ndtabi_kd_node, boundary) == 8 &&
              offsetof(ndtabi_kd_node, num) == 16 && offsetof(ndtabi_kd_node, obj_ids) == 24 &&
              offsetof(ndtabi_kd_node, left) == 40 && offsetof(ndtabi_kd_node, right) == 48);
NDTABI_ASSERT(sizeof(ndtabi_kd_tree) == 144 && offsetof(ndtabi_kd_tree, bb_upper) == 48 &&
              offsetof(ndtabi_kd_tree, inf_obj_ptrs) == 104 && offsetof(ndtabi_kd_tree, obj_num) == 120 &&
              offsetof(ndtabi_kd_tree, inf_obj_num) == 124 && offsetof(ndtabi_kd_tree, root) == 128);
NDTABI_ASSERT(sizeof(ndtabi_image) == 32 && offsetof(ndtabi_image, pixels) == 24);
#endif

#ifdef __cplusplus
}
#endif
#endif /* NDT_ABI_H */
