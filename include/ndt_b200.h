/*
 * ndt_b200.h -- C ABI of libndt_b200.so, the B200 drop-in for ndt's render path.
 *
 * The reference renders a frame with
 *     render_image(scene*, name, depth_name, w, h, samples, stereo_mode, threads,
 *                  aa_diff, aa_depth, max_optic_depth, img_copy, depth_copy)
 * (ndt.c:900), called once per frame from main() (ndt.c:1933) after the kd-tree
 * build (ndt.c:1899-1908) and camera_aim (ndt.c:1925).  Everything below that
 * call -- the pthread row loop (ndt.c:803-849), render_pixel / get_pixel_color
 * / get_ray_color / apply_lights (ndt.c:71-653), trace_kd and the kd traversal
 * (object.c:683-747, kd-tree.c:482-625), the per-type intersect functions
 * (objects/<type>.c) and pixel_d2c (image.h:36-39) -- is what this library
 * replaces.  Everything above it (CLI, scene and object plugins, kd build,
 * camera aim) keeps running unmodified on the host.
 *
 * Plain pointers and sizes only: no torch / CUDA types in any signature
 * (stream handles travel as void*).  All functions return 0 on success or a
 * negative ndt_b200_status; ndt_b200_last_error() gives the message for the
 * calling thread.  The library never calls exit() (the reference's convention,
 * object.c:233-236, is not acceptable for a library).  There is NO CPU
 * fallback: without a usable CUDA device every device entry point fails with
 * NDT_B200_E_CUDA.
 */
#ifndef NDT_B200_H
#define NDT_B200_H
#include <stddef.h>
#include <stdint.h>
#include "ndt_flat.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ndt_b200_status {
    NDT_B200_OK = 0,
    NDT_B200_E_ARG = -1,          /* bad argument / inconsistent scene */
    NDT_B200_E_UNSUPPORTED = -2,  /* object type, light type or camera outside the device path */
    NDT_B200_E_NOMEM = -3,
    NDT_B200_E_CUDA = -4,         /* CUDA runtime failure (message has the cudaError string) */
    NDT_B200_E_OVERFLOW = -5,     /* ray pool / traversal stack exhausted: render smaller tiles */
    NDT_B200_E_STATE = -6         /* call order (e.g. render before upload) */
} ndt_b200_status;

typedef struct ndt_b200_ctx ndt_b200_ctx;

/* Host services the flattener needs from the ndt host it is embedded in.
 * object_get_bounds: object.c:582 (Nelder-Mead bounding sphere, bounding.c:177);
 * the flattener calls it for every object whose lazily computed bounds have
 * not been produced yet (object.c:608-615), so that both sides of a parity
 * check use the very same centre/radius. */
typedef struct ndt_b200_host_api {
    int (*object_get_bounds)(void *object);
    /* vectNd_rotate2 (vectNd.c:271): only needed for VR / PANO cameras in a stereo mode, where
     * get_pixel_color rotates the eye about the camera by each column's azimuth (ndt.c:519-525).
     * May be NULL otherwise. */
    int (*vectNd_rotate2)(void *v, void *center, void *v1, void *v2, double angle, void *res);
} ndt_b200_host_api;

/* What one render call did.  "rays" are nearest-hit queries, i.e. trace_kd
 * calls (object.c:683): primary, reflection/refraction, shadow. */
typedef struct ndt_b200_stats {
    uint64_t rays_primary;
    uint64_t rays_bounce;
    uint64_t rays_shadow;
    uint64_t rays_unique;     /* sum of the three */
    uint64_t rays_ref;        /* what the reference executes for the same frame: every
                                 pixel's ray tree times its sample-loop count (ndt.c:488) */
    uint64_t samples;         /* sum over pixels of sample-loop iterations */
    uint64_t flops;           /* algorithmic fp64 operations (counting build only, else 0) */
    uint64_t launches;        /* kernels launched */
    uint32_t generations;     /* bounce generations processed */
    uint32_t reserved;
    double device_ms;         /* CUDA-event time of the kernels of this call */
} ndt_b200_stats;

/* ---- host side: struct-ABI adapter ------------------------------------- */

/* Replaces the entry of render_image (ndt.c:900-929): reads the host's scene
 * and the global kd-tree (ndt.c:68) through the struct ABI in ndt_abi.h,
 * forces all lazily prepared state (sphere.c:18-32, hcube.c:155-170,
 * object.c:608-615), applies the dirX *= w/h scaling (ndt.c:926) to its own
 * copy and emits the flat scene.  `scene` is a `scene*`, `kdtree` a
 * `kd_tree_t*` of the reference.  Unknown object types, custom
 * get_color/get_reflect hooks and area lights are refused with
 * NDT_B200_E_UNSUPPORTED. */
int ndt_b200_flatten(const void *scene, const void *kdtree, int width, int height,
                     int max_optic_depth, int specular,
                     const ndt_b200_host_api *host, ndt_flat_scene **out);
/* Same for any stereo_mode of ndt.c:46-48 (MONO, SIDE_SIDE_3D, OVER_UNDER_3D, ANAGLYPH_3D,
 * HIDEF_3D) and any camera type (NORMAL, VR, PANO: camera.c:504-556); ndt_b200_flatten is
 * stereo_mode = MONO.  width x height is the OUTPUT frame (HIDEF_3D: height 2205). */
int ndt_b200_flatten_view(const void *scene, const void *kdtree, int width, int height,
                          int max_optic_depth, int specular, int stereo_mode,
                          const ndt_b200_host_api *host, ndt_flat_scene **out);
/* The scene for recursive anti-aliasing (-w / -a; render_image with the global recursive_aa set,
 * ndt.c:921-926): header width/height are the (width+1) x (height+1) grid of corner samples.
 * Render it with ndt_b200_render_aa. */
int ndt_b200_flatten_aa(const void *scene, const void *kdtree, int width, int height,
                        int max_optic_depth, int specular,
                        const ndt_b200_host_api *host, ndt_flat_scene **out);
void ndt_b200_free_flat(ndt_flat_scene *fs);
/* structural check of a blob received from disk or another rank */
int ndt_b200_flat_validate(const void *blob, size_t bytes);

/* ---- device side --------------------------------------------------------- */

int ndt_b200_init(int device, ndt_b200_ctx **ctx);
void ndt_b200_destroy(ndt_b200_ctx *ctx);

/* copy the flat scene to HBM (one cudaMemcpyAsync of the whole blob) */
int ndt_b200_upload(ndt_b200_ctx *ctx, const ndt_flat_scene *fs);

/* options: 0 = default */
#define NDT_B200_OPT_COUNT_FLOPS 1u   /* run the instrumented kernels and fill stats.flops */
#define NDT_B200_OPT_FUSED 2u         /* one fused kernel per bounce generation instead of the trace/shade wavefront (A/B measurements) */
int ndt_b200_set_options(ndt_b200_ctx *ctx, uint32_t options);

/* Sizing of the per-render ray pools.  A pass over n0 primary slots gets a record pool of
 * n0 * (1 + bounce_factor) + slack_records entries (defaults 6.0 and 65536) and works a generation off in
 * batches of at most rays_per_batch rays (default 2^23; rounded up to a multiple of 32; 0 = default).
 * ndt_b200_render_tile recovers from an exhausted pool by halving the tile (a single row that still does
 * not fit is retried once with a larger pool), so small values are safe, only slower -- the tests use them
 * to exercise exactly that recovery and the batching of the device-side generation loop. */
int ndt_b200_set_pool(ndt_b200_ctx *ctx, double bounce_factor, int slack_records, int rays_per_batch);

/* Render the tile [x0,x0+tw) x [y0,y0+th) of the uploaded frame into HOST
 * buffers laid out tile-row-major (tw*th elements): fp64 RGBA as render_line
 * stores it (ndt.c:752, image.h:22-26), 8-bit RGBA through pixel_d2c
 * (image.h:36-39), hit flag and object id of the primary ray (what
 * get_ray_color tests at ndt.c:376) and 1/distance (ndt.c:363-373).  Any
 * output pointer may be NULL.  Synchronous; device->host copies included. */
int ndt_b200_render_tile(ndt_b200_ctx *ctx, int x0, int y0, int tw, int th,
                         double *rgba_f64, uint8_t *rgba_u8, uint8_t *hit,
                         int32_t *obj_id, double *inv_depth, ndt_b200_stats *stats);

/* Same, but the outputs are DEVICE pointers and the work is only enqueued on the
 * context's stream (ndt_b200_stream): one small kernel and ONE CUDA-graph launch per
 * frame -- the loop over bounce generations runs on the device -- so the call returns
 * without waiting for the GPU.  What goes wrong inside the frame (ray pool exhausted:
 * NDT_B200_E_OVERFLOW) is therefore reported by ndt_b200_sync(), which also makes the
 * statistics available to ndt_b200_last_stats(); this entry point does not retry --
 * render a smaller tile, or use ndt_b200_render_tile, which does. */
int ndt_b200_launch_tile(ndt_b200_ctx *ctx, int x0, int y0, int tw, int th,
                         void *d_rgba_f64, void *d_rgba_u8, void *d_hit,
                         void *d_obj_id, void *d_inv_depth);
int ndt_b200_sync(ndt_b200_ctx *ctx);
int ndt_b200_last_stats(ndt_b200_ctx *ctx, ndt_b200_stats *stats);
/* the stream all of a context's work is enqueued on (cudaStream_t as void*) */
void *ndt_b200_stream(ndt_b200_ctx *ctx);

/* Recursive (Whitted) anti-aliasing, the second half of render_image when recursive_aa is set
 * (ndt.c:655-733, 1039-1088): the uploaded scene must come from ndt_b200_flatten_aa.  Renders the
 * (W+1) x (H+1) corner samples, then refines every pixel whose four corners differ by more than
 * aa_diff/255 (image_avg_dbl_pixels4, image.c:1175) level by level -- each level's new samples are
 * one more wavefront over an explicit sample list -- down to steps of 1/2^aa_depth (ndt.c:663).
 * Output: the W x H frame as the reference stores it, 8-bit RGBA through pixel_d2c (actual_img is an
 * 8-bit image, ndt.c:940-942), plus optionally the fp64 colours before quantisation.  HOST buffers. */
int ndt_b200_render_aa(ndt_b200_ctx *ctx, int aa_diff, int aa_depth,
                       uint8_t *rgba_u8, double *rgba_f64, uint64_t *pixels_resampled, ndt_b200_stats *stats);

/* Drop-in for render_image (ndt.c:900), same parameters and return value
 * (1).  `kdtree` is the extra argument: the reference reads its global
 * (ndt.c:68), a library cannot.  Supports what the device path supports:
 * samples == 1, every stereo mode and camera type, recursive_aa off; anything
 * else returns a negative status instead of rendering.  img_copy receives the fp64
 * frame when name and img_copy are non-NULL, depth_copy the depth map when depth_name
 * and depth_copy are (ndt.c:1024-1031: image_copy, which initialises the destination
 * without freeing it); writing the image FILES is the binding's job -- the codecs stay
 * on the host.  cam.dirX is rescaled in place like ndt.c:925-926 once
 * the frame has been rendered.  Environment: NDT_B200_DEVICES=<n>|all renders the frame
 * on that many GPUs of the box (ndt_b200_mgpu_render_frame); default 1. */
int ndt_b200_render_image(void *scene, const void *kdtree, const ndt_b200_host_api *host,
                          char *name, char *depth_name, int width, int height,
                          int samples, int stereo_mode, int threads, int aa_diff,
                          int aa_depth, int max_optic_depth, int specular,
                          void *img_copy, void *depth_copy);

/* The same call when the reference's global recursive_aa (ndt.c:44, set by -w / -a) is non-zero:
 * ndt_b200_flatten_aa + ndt_b200_render_aa; img_copy receives the 8-bit actual_img (ndt.c:1124-1127).
 * MONO only. */
int ndt_b200_render_image_aa(void *scene, const void *kdtree, const ndt_b200_host_api *host,
                             char *name, char *depth_name, int width, int height,
                             int samples, int stereo_mode, int threads, int aa_diff,
                             int aa_depth, int max_optic_depth, int specular,
                             void *img_copy, void *depth_copy);

/* ---- all GPUs of one box (mgpu.cu) ------------------------------------------
 * Replaces the MPI layer of the reference for one node: rows of a frame (ndt.c:812-820) and frames of an
 * animation (ndt.c:1771-1787) are spread over the GPUs, results land in the caller's host buffers -- what
 * mpi_collect_image (ndt.c:1277-1309) does for rank 0.  One host thread per context, two contexts per GPU,
 * work items pulled from one atomic counter (dynamic queue); no collective on the data path.
 * n_devices <= 0: every visible GPU; devices == NULL: 0 .. n_devices-1. */
typedef struct ndt_b200_mgpu ndt_b200_mgpu;
int ndt_b200_mgpu_init(int n_devices, const int *devices, ndt_b200_mgpu **out);
void ndt_b200_mgpu_destroy(ndt_b200_mgpu *m);
int ndt_b200_mgpu_devices(const ndt_b200_mgpu *m);
/* ONE frame as row bands of band_rows rows (<= 0: about four bands per context): every context uploads the scene
 * and pulls bands until none is left; each band is copied from its GPU into the rows it covers of the HOST
 * buffers (layout and meaning as ndt_b200_render_tile for the whole frame; any may be NULL).  Byte-identical to
 * the one-GPU render of the same frame: a pixel does not depend on which tile it is in. */
int ndt_b200_mgpu_render_frame(ndt_b200_mgpu *m, const ndt_flat_scene *fs, int band_rows,
                               double *rgba_f64, uint8_t *rgba_u8, uint8_t *hit, int32_t *obj_id,
                               double *inv_depth, ndt_b200_stats *stats);
/* An ANIMATION, frame by frame: the scene is copied, the call returns as soon as the frame is queued (it blocks
 * while 2 x contexts frames are already waiting), whichever context is free uploads and renders it into the HOST
 * buffers (either may be NULL), which must stay valid until ndt_b200_mgpu_wait.  This is how a stateful scene
 * (scenes/balls.c: scene_setup must run for every frame in order, ndt.c:1816-1825) keeps all GPUs busy: the
 * host produces flat scenes sequentially, the frames render concurrently. */
int ndt_b200_mgpu_submit(ndt_b200_mgpu *m, const ndt_flat_scene *fs, uint8_t *rgba_u8, double *rgba_f64);
/* all submitted frames are in their buffers; statistics summed over them; the first failure, if any */
int ndt_b200_mgpu_wait(ndt_b200_mgpu *m, ndt_b200_stats *stats);

/* Page-locked host memory (cudaMallocHost) for render_tile outputs and flat
 * scenes; plain malloc'ed buffers work too, only slower to copy. */
void *ndt_b200_host_alloc(size_t bytes);
void ndt_b200_host_free(void *p);

/* Drop-in for kd_tree_build(kd_tree_t*, kd_item_list_t*) (kd-tree.c:421-477),
 * the serial pre-pass in front of the render path (ndt.c:1908; 10-13 s per
 * frame for BASELINE config 2).  Same arguments, same result: the reference's
 * own kd_tree_t in host memory, node for node and bit for bit (the exhaustive
 * split search of kd-tree.c:315-345 runs on the GPU, one thread per candidate
 * plane).  `kd_tree` must have been initialised with kd_tree_init.  Returns
 * what kd_tree_build returns (1: the root stayed a leaf, 0: it was split) or
 * a negative status. */
int ndt_b200_kd_tree_build(void *kd_tree, void *kd_item_list);

/* The same exhaustive plane search with a bounded recursion, for scenes the reference's builder does
 * not finish (scenes/random.c beyond a few hundred objects: every plane is straddled by many objects,
 * both sides keep them, kd-tree.c:381-403 recurses on ever larger lists).  Stops at max_depth, at
 * leaf_size items, and where a split would grow the reference count by more than max_growth (> 1).
 * Same kd_tree_t as output; not the tree the reference would build, so hit / id parity is checked
 * against the reference's tree-less trace() (object.c:692) instead of its kd result. */
int ndt_b200_kd_tree_build_bounded(void *kd_tree, void *kd_item_list, int max_depth, int leaf_size, double max_growth);

/* trace_kd (object.c:683) for n_rays explicit rays: origins/dirs are n_rays x N
 * doubles (row-major, HOST), dist_limits may be NULL (= -1.0, "check all
 * objects", ndt.c:172-183).  Outputs per ray: trace_kd's return value, the id
 * of the object it reported (-1 = NULL), the accepted distance, the hit point
 * and the normal exactly as the plugin returned it (not normalised).  This is
 * the probe the per-primitive known-answer tests use. */
int ndt_b200_trace_rays(ndt_b200_ctx *ctx, int n_rays, const double *origins, const double *dirs,
                        const double *dist_limits, int32_t *found, int32_t *obj_id,
                        double *t, double *hit, double *normal);

/* one FP64 pipe probe: returns sustained GFLOP/s of an all-SM chain of
 * dependent-free DFMA (fused=1) or DMUL+DADD pairs (fused=0); the roofline
 * denominators of bench.py */
int ndt_b200_fp64_peak(ndt_b200_ctx *ctx, int fused, double *gflops);

/* the sample loop of get_pixel_color (ndt.c:488-568) for n traced colours rgba_in[n][4]: the averaged colour
 * the loop converges to (rgba_out[n][4]) and the number of identical samples it takes (samples[n]).  k_finish
 * runs exactly this per pixel; exposed as the probe of the known-answer test of its division sequence. */
int ndt_b200_replay_samples(ndt_b200_ctx *ctx, int n, const double *rgba_in, double *rgba_out, int32_t *samples);

const char *ndt_b200_last_error(void);
const char *ndt_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NDT_B200_H */
