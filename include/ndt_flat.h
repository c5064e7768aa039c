/*
 * ndt_flat.h -- the flattened scene handed to the device (one blob).
 *
 * One contiguous, position-independent allocation: a header followed by dense
 * arrays addressed by byte offsets from the start of the blob, so the same
 * bytes can be memcpy'd to HBM, written to disk or sent to another rank.
 * Everything numeric is fp64 / int32.  Every N-vector is stored `npad` wide
 * (npad = N rounded up to even) because the reference's SSE2 vector ops work
 * on lane PAIRS and its dot product sums even and odd lanes separately
 * (vectNd.h:215-227); the pad lane holds what the reference holds there (0.0).
 *
 * Arrays (structure-of-arrays at the subsystem level):
 *   objects[]  fixed-size records: type, material, geometry offset, nesting
 *   bspheres[] (npad+2) doubles per object: centre, radius, radius^2
 *              -- kept apart from the geometry because the bounding-sphere
 *              pre-test (bounding.c:34-85) is by far the most frequent fetch
 *   geom[]     fp64 pool, per-type packed blocks (layouts below)
 *   nodes[]    kd-tree nodes in pre-order (kd-tree.h:52-59 flattened)
 *   leaf_refs[] object ids of all leaves, in the reference's in-leaf order
 *   inf_ids[]  ids of infinite objects in kd_tree_t.inf_obj_ptrs order
 *   lights[]   fixed-size records + vectors in geom[]
 *
 * Object ids are the reference's kd item ids (kd-tree.c:448): the position of
 * the object in the object_kdlist_add walk (object.c:633-681, clusters
 * expanded in place).  Objects synthesised inside another object (hcube
 * faces, hcube.c:33-152) follow after the n_items top-level ones.
 */
#ifndef NDT_FLAT_H
#define NDT_FLAT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDT_FLAT_MAGIC   0x3146444eu /* "NDF1" */
#define NDT_FLAT_VERSION 5u   /* 4: geometry blocks 16-byte aligned; 5: view tables (VR / PANO cameras, stereo modes) */
#define NDT_MAX_DIM      16

/* reference tolerances: vectNd.h:24-29, object.h:15-18 */
#define NDT_EPS  (1e-4)
#define NDT_EPS2 ((NDT_EPS) * (NDT_EPS))

enum ndt_obj_type {
    NDT_T_SPHERE = 0,   /* objects/sphere.c    */
    NDT_T_HPLANE,       /* objects/hplane.c    */
    NDT_T_HDISK,        /* objects/hdisk.c     */
    NDT_T_ORTHOTOPE,    /* objects/orthotope.c */
    NDT_T_HCUBE,        /* objects/hcube.c     */
    NDT_T_FACET,        /* objects/facet.c     */
    NDT_T_HFACET,       /* objects/hfacet.c    */
    NDT_T_CYLINDER,     /* objects/cylinder.c  */
    NDT_T_HCYLINDER,    /* objects/hcylinder.c */
    NDT_T_COUNT
};

enum ndt_obj_flags {
    NDT_OF_TRANSPARENT = 1,  /* object.h:24 */
    NDT_OF_NO_END_TEST = 2,  /* cylinder.c:87 / hcylinder.c:107: infinite, skip the end test */
    NDT_OF_USE_NORMALS = 4   /* hfacet.c:282: interpolate the three vertex normals */
};

enum ndt_light_type { /* scene.h:17-23 */
    NDT_L_AMBIENT = 0, NDT_L_POINT = 1, NDT_L_DIRECTIONAL = 2, NDT_L_SPOT = 3
};

/*
 * geom[] block layouts (all vectors npad doubles; A = n_axes).  Every block
 * starts at an EVEN index of geom[] (16-byte aligned in the blob and in HBM) so
 * that a block can be staged into shared memory with one TMA bulk copy:
 *  SPHERE    c, r^2                                         (sphere.c:18-32)
 *  HPLANE    p, n                                           (hplane.c:39-75)
 *  HDISK     p, n, r                                        (hdisk.c:15-34,61-85)
 *  ORTHOTOPE p0, basis[A], len[A], BdB[A], BdP[A]           (orthotope.c:23-54)
 *  HCUBE     (nothing; children [child_begin, child_begin+child_count))
 *  FACET     p[3], basis[2], normal, AdA[2], BdA[2], angle[3]   (facet.c:42-83,166-269)
 *  HFACET    v0, uedge0, eperp, normals[3],
 *            x2, y2, x3, y3, ones_pad                       (hfacet.c:43-92,147-188)
 *  CYLINDER  p0, axis, length, AdA, BdA, r                  (cylinder.c:22-41)
 *  HCYLINDER p0, axes[A], len[A], AdA[A], BdA[A], r         (hcylinder.c:23-54)
 */
typedef struct ndt_flat_object {
    int32_t type;
    int32_t flags;
    int32_t report_id;    /* id reported as the hit object (hcube.c:244-247) */
    int32_t n_axes;
    int32_t child_begin;
    int32_t child_count;
    uint32_t geom_off;    /* index of the first double of this object's block in geom[] */
    uint32_t reserved;
    double rgb[3];        /* object.h:26 */
    double refl[3];       /* object.h:27 */
    double refract_index; /* object.h:28 */
    double bs_radius;     /* copy of bspheres[].radius: >0 pre-test, <=0 none (object.c:618) */
} ndt_flat_object;        /* 96 bytes */

typedef struct ndt_flat_node {
    int32_t dim;          /* <0: leaf (kd-tree.c:369) */
    int32_t left, right;  /* node index or -1 */
    int32_t leaf_begin;   /* into leaf_refs[] */
    int32_t leaf_count;   /* kd_node_t.num */
    int32_t reserved;
    double boundary;
} ndt_flat_node;          /* 32 bytes */

typedef struct ndt_flat_light {
    int32_t type;
    int32_t reserved;
    double rgb[3];
    double angle;         /* SPOT cone, degrees (ndt.c:204) */
    double max_rgb;       /* MAX(r, MAX(g, b)) as image.h:33 evaluates it (ndt.c:302) */
    uint32_t vec_off;     /* geom[]: pos, dir, rev_unit, near_off (npad each) */
    uint32_t reserved2;
} ndt_flat_light;         /* 56 bytes */

typedef struct ndt_flat_header {
    uint32_t magic, version;
    uint64_t total_bytes;
    int32_t n, npad;
    int32_t width, height;          /* the frame the camera basis was scaled for (ndt.c:926) */
    int32_t max_optic_depth;        /* ndt.c:1413 */
    int32_t specular;               /* ndt.c:41 */
    int32_t n_items;                /* kd items = id space of hit buffers */
    int32_t n_objects;              /* n_items + nested */
    int32_t n_nodes, n_leaf_refs, n_inf, n_lights;
    int32_t max_leaf;               /* largest leaf_count */
    int32_t tree_depth;
    int32_t use_focal;              /* screen_dist > EPSILON (camera.c:568) */
    int32_t reserved;
    double bg[4];                   /* scene.h:60 */
    double ambient[3];              /* scene.h:59 (scn->ambient.red/green/blue) */
    double focal_scale;             /* focal_distance / screen_dist (camera.c:573) */
    uint64_t off_camera;            /* doubles: pos, imgOrig, dirX (scaled), dirY */
    uint64_t off_aabb;              /* doubles: lower[npad], upper[npad] (kd_tree_t.bb) */
    uint64_t off_objects;
    uint64_t off_bspheres;
    uint64_t off_geom;
    uint64_t n_geom;                /* doubles in geom[] */
    uint64_t off_nodes;
    uint64_t off_leaf_refs;
    uint64_t off_inf;
    uint64_t off_lights;
    /* -- version 5: cameras other than CAMERA_NORMAL and the stereo modes of render_pixel
     *    (camera.c:504-556, ndt.c:578-653).  off_view == 0 means CAMERA_NORMAL + MONO: x and y
     *    come from the pixel index (ndt.c:632-633).  Otherwise off_view points at doubles
     *      ext[5*npad]        leftEye, rightEye, localX, localY, localZ (camera.h:60-75)
     *      cols[width][4]     x, sin(x*hFov), cos(x*hFov), eye (0 centre 1 left 2 right)
     *      rows[height][6]    y, sin(y*vFov), cos(y*vFov), y*y_size (PANO, camera.c:540),
     *                         blank (HIDEF_3D rows 1080..1125, ndt.c:619-626), eye
     *      eyes[width][2*npad] only if view_eyes: the eye rotated about cam.pos by the column's
     *                         azimuth (VR / PANO stereo, ndt.c:519-525), left then right
     *    All trigonometry is evaluated by the HOST's libm while flattening (W + H values), so the
     *    device's primary rays are bit-identical to the reference's. */
    int32_t cam_type;               /* camera.h:16-20: 0 CAMERA_NORMAL, 1 CAMERA_VR, 2 CAMERA_PANO */
    int32_t stereo_mode;            /* ndt.c:46-48 */
    int32_t view_eyes;              /* eyes[] present */
    int32_t aa_pad;                 /* 1: width/height are the (W+1) x (H+1) sample grid of the recursive
                                       anti-aliasing pass (ndt.c:921-924), dirX scaled by W/H */
    uint64_t off_view;
    double cam_dist;                /* cam.focal_distance as passed to camera_target_point (ndt.c:515) */
} ndt_flat_header;

enum ndt_stereo_mode { NDT_MONO = 0, NDT_SIDE_SIDE_3D, NDT_OVER_UNDER_3D, NDT_ANAGLYPH_3D, NDT_HIDEF_3D };
enum ndt_cam_type { NDT_CAM_NORMAL = 0, NDT_CAM_VR = 1, NDT_CAM_PANO = 2 };

typedef struct ndt_flat_scene {     /* the blob starts with its header */
    ndt_flat_header h;
} ndt_flat_scene;

#define NDT_FLAT_PTR(fs, type, off) ((type *)((char *)(fs) + (off)))

#ifdef __cplusplus
}
#endif
#endif /* NDT_FLAT_H */
