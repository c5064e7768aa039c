/*
 * kernels.cu -- the host half of the wavefront and the device half of the C ABI (ndt_b200.h).
 *
 * One render = one pass over the tile's rays, generation by generation, with the loop ON THE DEVICE:
 *   k_begin            resets WaveState (gen.cuh): generation 0 = the tile's pixels (8x4 blocks per warp)
 *   WHILE (CUDA graph conditional node; condition set by k_next_gen)
 *     k_pre<NP,0>      every ray of the batch: the infinite objects and the root box; rays that do not
 *                      enter it are answered, the others listed                       -> HitRec, walker list
 *     k_trace<NP,0>    the listed rays walk the kd-tree: nearest hit                  -> HitRec
 *     k_shade<NP,A>    hit point, normal, side tests; one shadow query per light that needs one
 *     k_pre<NP,1>, k_trace<NP,1>   the same for the shadow queries                    -> HitRec per (ray, light)
 *     k_light<NP>, k_libm          per query: the vector half of the lit term, then acos / cos / pow
 *     k_shade<NP,B>    the light loop in the reference's order, RayRec, reflection / refraction rays
 *                      appended to the next generation (two ballots + one atomic per warp)
 *     k_next_gen       next batch / next generation / stop
 *   k_pre_resolve, WHILE { k_resolve, k_resolve_next }   fold generation g+1 into g, deepest first
 *   k_finish           generation 0 -> pixels: sample-loop replay (ndt.c:488), fp64 RGBA as 2x16-byte
 *                      stores, u8 RGBA as one 4-byte store per pixel (pixel_d2c, image.h:36-39)
 * The host enqueues k_begin + ONE graph launch per frame and never waits inside a frame; errors (pool
 * exhausted, traversal stack) are flags in WaveState that ndt_b200_sync reports.  Without graph support
 * (or with NDT_B200_NO_GRAPH=1) the same kernels run from a host loop that reads `cont` back per batch.
 * The fused k_generation path (NDT_B200_OPT_FUSED, the counting build) keeps its host loop.
 * fp64 arithmetic is never contracted (-fmad=false) -- see core.cuh.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "ndt_abi.h"
#include "ndt_internal.h"
#include "gen.cuh"

/* per-NP launch tables (np_inst.cu); -DNDT_ONLY_NP=8 links a single dimension (fast experiment builds) */
#ifdef NDT_ONLY_NP
#define NP_DECL(n)
#define NP_REF(n) ((n) == NDT_ONLY_NP ? &NP_ONLY_SYM : NULL)
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define NP_ONLY_SYM CAT(ndt_np_ops_, NDT_ONLY_NP)
extern const NpOps NP_ONLY_SYM;
#else
extern const NpOps ndt_np_ops_4, ndt_np_ops_6, ndt_np_ops_8, ndt_np_ops_10, ndt_np_ops_12;
#define NP_REF(n) (&ndt_np_ops_##n)
#endif
const NpOps *ndt_np_ops(int np)
{
    switch (np) {
    case 4: return NP_REF(4);
    case 6: return NP_REF(6);
    case 8: return NP_REF(8);
    case 10: return NP_REF(10);
    case 12: return NP_REF(12);
    }
    return NULL;
}

__global__ void k_resolve(RayRec *rec, int start, int count, int specular)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    RayRec r;
    rec_load(r, rec + start + i);
    if (r.child_refl == CHILD_NONE && r.child_refr == CHILD_NONE) return;
    const RayRec *c1 = r.child_refl >= 0 ? rec + r.child_refl : nullptr;
    const RayRec *c2 = r.child_refr >= 0 ? rec + r.child_refr : nullptr;
    resolve_rec(r, c1, c2, specular);
    rec_store(rec + start + i, r);
}

/* generation 0 -> pixels */
__global__ void k_finish(RayRec *rec, int tw, int th, int bpr, int specular,
                         double *out_f64, uint8_t *out_u8, unsigned long long *stats)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long rays_ref = 0, samples = 0, hits = 0, traced = 0;
    if (p < tw * th) {
        const int tx = p % tw, ty = p / tw;
        /* bpr == 0: a sample list, record r belongs to sample r */
        const int slot = bpr ? ((ty >> 2) * bpr + (tx >> 3)) * 32 + ((ty & 3) << 3) + (tx & 7) : p;
        RayRec r;
        rec_load(r, rec + slot);
        const RayRec *c1 = r.child_refl >= 0 ? rec + r.child_refl : nullptr;
        const RayRec *c2 = r.child_refr >= 0 ? rec + r.child_refr : nullptr;
        resolve_rec(r, c1, c2, specular);
        double l[4] = { r.clr[0], r.clr[1], r.clr[2], r.alpha }, o[4] = { 0.0, 0.0, 0.0, 0.0 };
        /* a pixel render_pixel leaves black without calling get_pixel_color has no sample loop */
        const int ns = (r.flags & REC_UNTRACED) ? 0 : replay_samples(l, o);
        if (out_f64) {
            double2 *d = reinterpret_cast<double2 *>(out_f64 + 4 * (size_t)p);
            d[0] = make_double2(o[0], o[1]);
            d[1] = make_double2(o[2], o[3]);
        }
        if (out_u8) {
            uchar4 c = make_uchar4(d2c(o[0]), d2c(o[1]), d2c(o[2]), d2c(o[3]));
            reinterpret_cast<uchar4 *>(out_u8)[p] = c;
        }
        rays_ref = (unsigned long long)r.nrays * (unsigned long long)ns;
        samples = (unsigned long long)ns;
        hits = r.flags & 1u;
        traced = (r.flags & REC_UNTRACED) ? 0 : 1;
    }
    for (int d = 16; d > 0; d >>= 1) {
        rays_ref += __shfl_down_sync(0xffffffffu, rays_ref, d);
        samples += __shfl_down_sync(0xffffffffu, samples, d);
        hits += __shfl_down_sync(0xffffffffu, hits, d);
        traced += __shfl_down_sync(0xffffffffu, traced, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (rays_ref) atomicAdd(&stats[2], rays_ref);
        if (samples) atomicAdd(&stats[3], samples);
        if (hits) atomicAdd(&stats[4], hits);
        if (traced) atomicAdd(&stats[5], traced);
    }
}

constexpr unsigned FULL_MASK = 0xffffffffu;
/* the scalar half: acos, cos, pow (ndt.c:261-268, 300) per lit pair */
__global__ void __launch_bounds__(256) k_libm(const WaveArgs a, int specular, int aux_word)
{
    const int lane = threadIdx.x & 31;
    WaveState *st = a.st;
    int nq = *(volatile const int *)&st->stail;
    if (nq > a.scap) nq = a.scap;
    const size_t sstride = (size_t)a.gen_cap * a.nl_eff;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&st->nextM, 128);
        base = __shfl_sync(FULL_MASK, base, 0);
        if (base >= nq) break;
        for (int k = 0; k < 4; ++k) {
            const int j = base + 32 * k + lane;
            if (j >= nq) continue;
            /* where the answer of query j lives: the high word of its last 16-byte word (ray_store_s) */
            const int aux = __double2hiint(a.srays[(size_t)aux_word * a.scap + j].y);
            const double2 g0 = a.shits[aux], g1 = a.shits[sstride + aux];
            const int lit = __double2loint(g1.y), qok = __double2hiint(g1.y);
            if (!lit) continue;
            double light_scale, rvn;
            light_libm(g0.x, qok, g0.y, g1.x, specular, light_scale, rvn);
            a.shits[aux] = make_double2(light_scale, rvn);
        }
    }
}

/* ---- bookkeeping kernels of the device-side generation loop (one thread each) ---- */
__global__ void k_begin(WaveState *st, const WaveBegin b, int gen_cap, unsigned long long *stats)
{
    if (threadIdx.x || blockIdx.x) return;
    st->n0 = b.n0; st->x0 = b.x0; st->y0 = b.y0; st->tw = b.tw; st->th = b.th; st->bpr = b.bpr;
    st->eye = b.eye; st->first = b.first;
    st->samples_xy = b.samples_xy;
    st->out_hit = b.out_hit; st->out_id = b.out_id; st->out_depth = b.out_depth;
    st->out_f64 = b.out_f64; st->out_u8 = b.out_u8;
    st->gen = 0; st->start = 0; st->count = b.n0 < gen_cap ? b.n0 : gen_cap;
    st->gen_start = 0; st->gen_count = b.n0;
    st->ngen = 0; st->iters = 0; st->cont = 1; st->resolve_g = 0; st->fail = 0;
    st->nextP0 = 0; st->nextP1 = 0; st->wfull0 = 0; st->wpart0 = 0; st->wfull1 = 0; st->wpart1 = 0;
    st->tail = b.n0; st->next0 = 0; st->stail = 0; st->next1 = 0; st->nextA = 0; st->nextB = 0; st->nextR = 0; st->nextF = 0; st->nextL = 0; st->nextM = 0;
    st->pool_overflow = 0; st->kd_fault = 0;
    if (b.first) for (int k = 0; k < 8; ++k) stats[k] = 0ull;
}

/* after the four launches of a batch: the next batch of the generation, or the next generation, or the end.
 * `h` is the condition of the graph's WHILE node (use_h = 0: launched from the host loop). */
__global__ void k_next_gen(WaveState *st, int cap, int gen_cap, cudaGraphConditionalHandle h, int use_h)
{
    if (threadIdx.x || blockIdx.x) return;
    int cont = 0;
    st->iters += 1;
    const int bend = st->start + st->count, gend = st->gen_start + st->gen_count;
    if (st->pool_overflow || st->kd_fault) {
        cont = 0;
    } else if (bend < gend) {
        st->start = bend;
        st->count = (gend - bend) < gen_cap ? (gend - bend) : gen_cap;
        cont = 1;
    } else {
        const int g = st->ngen;
        st->gstart[g] = st->gen_start;
        st->gcount[g] = st->gen_count;
        st->ngen = g + 1;
        int tail = st->tail;
        if (tail > cap) tail = cap;          /* cannot happen without pool_overflow */
        const int ncount = tail - gend;
        if (ncount > 0) {
            if (g + 1 >= WAVE_MAX_GEN) {
                st->pool_overflow = 3;
            } else {
                st->gen += 1;
                st->gen_start = gend; st->gen_count = ncount;
                st->start = gend;
                st->count = ncount < gen_cap ? ncount : gen_cap;
                cont = 1;
            }
        }
    }
    st->next0 = 0; st->stail = 0; st->next1 = 0; st->nextA = 0; st->nextB = 0; st->nextL = 0; st->nextM = 0;
    st->nextP0 = 0; st->nextP1 = 0; st->wfull0 = 0; st->wpart0 = 0; st->wfull1 = 0; st->wpart1 = 0;
    st->cont = cont;
    if (!cont) st->fail = st->pool_overflow | (st->kd_fault << 8);
    if (use_h) cudaGraphSetConditional(h, cont ? 1u : 0u);
}

/* between the two loops: the backward fold starts at the deepest generation */
__global__ void k_pre_resolve(WaveState *st, cudaGraphConditionalHandle h, int use_h)
{
    if (threadIdx.x || blockIdx.x) return;
    const int g = st->fail ? 0 : st->ngen - 1;
    st->resolve_g = g;
    st->nextR = 0; st->nextF = 0;
    if (use_h) cudaGraphSetConditional(h, g >= 1 ? 1u : 0u);
}
__global__ void k_resolve_next(WaveState *st, cudaGraphConditionalHandle h, int use_h)
{
    if (threadIdx.x || blockIdx.x) return;
    const int g = st->resolve_g - 1;
    st->resolve_g = g;
    st->nextR = 0;
    if (use_h) cudaGraphSetConditional(h, g >= 1 ? 1u : 0u);
}

/* the same two kernels for the device-side loop: what to fold comes from WaveState, the grids are fixed */
/* the children's colours and ray counts: words 0, 1 and 4 of their records */
__device__ __forceinline__ void child_load(RayRec &c, const double2 *rec, size_t cap, int slot)
{
    double2 *d = reinterpret_cast<double2 *>(&c);
    d[0] = rec[slot];
    d[1] = rec[cap + slot];
    d[4] = rec[4 * cap + slot];
}

__global__ void k_resolve_dev(double2 *rec, int cap, WaveState *st, int specular)
{
    if (st->fail) return;
    const int g = st->resolve_g;
    const int start = st->gstart[g], count = st->gcount[g];
    const int lane = threadIdx.x & 31;
    while (true) {      /* 128 records per warp and draw */
        int base = 0;
        if (lane == 0) base = atomicAdd(&st->nextR, 128);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        for (int k = 0; k < 4; ++k) {
            const int i = base + 32 * k + lane;
            if (i >= count) continue;
            RayRec r, c1, c2;
            rec_load_s(r, rec, (size_t)cap, (size_t)(start + i));
            if (r.child_refl == CHILD_NONE && r.child_refr == CHILD_NONE) continue;
            if (r.child_refl >= 0) child_load(c1, rec, (size_t)cap, r.child_refl);
            if (r.child_refr >= 0) child_load(c2, rec, (size_t)cap, r.child_refr);
            resolve_rec(r, r.child_refl >= 0 ? &c1 : nullptr, r.child_refr >= 0 ? &c2 : nullptr, specular);
            rec_store_s(rec, (size_t)cap, (size_t)(start + i), r);
        }
    }
}

__global__ void __launch_bounds__(256, 4) k_finish_dev(double2 *rec, int cap, WaveState *st, int specular, unsigned long long *stats)
{
    if (st->fail) return;
    const int tw = st->tw, th = st->th, bpr = st->bpr;
    double *out_f64 = st->out_f64;
    uint8_t *out_u8 = st->out_u8;
    unsigned long long rays_ref = 0, samples = 0, hits = 0, traced = 0;
    const int n = tw * th, lane = threadIdx.x & 31;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&st->nextF, 128);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        for (int k = 0; k < 4; ++k) {
            const int p = base + 32 * k + lane;
            if (p >= n) continue;
            const int tx = p % tw, ty = p / tw;
            /* bpr == 0: a sample list, record r belongs to sample r */
            const int slot = bpr ? ((ty >> 2) * bpr + (tx >> 3)) * 32 + ((ty & 3) << 3) + (tx & 7) : p;
            RayRec r, c1, c2;
            rec_load_s(r, rec, (size_t)cap, (size_t)slot);
            if (r.child_refl >= 0) child_load(c1, rec, (size_t)cap, r.child_refl);
            if (r.child_refr >= 0) child_load(c2, rec, (size_t)cap, r.child_refr);
            resolve_rec(r, r.child_refl >= 0 ? &c1 : nullptr, r.child_refr >= 0 ? &c2 : nullptr, specular);
            double l[4] = { r.clr[0], r.clr[1], r.clr[2], r.alpha }, o[4] = { 0.0, 0.0, 0.0, 0.0 };
            /* a pixel render_pixel leaves black without calling get_pixel_color has no sample loop */
            const int ns = (r.flags & REC_UNTRACED) ? 0 : replay_samples(l, o);
            if (out_f64) {
                double2 *d = reinterpret_cast<double2 *>(out_f64 + 4 * (size_t)p);
                d[0] = make_double2(o[0], o[1]);
                d[1] = make_double2(o[2], o[3]);
            }
            if (out_u8) {
                uchar4 c = make_uchar4(d2c(o[0]), d2c(o[1]), d2c(o[2]), d2c(o[3]));
                reinterpret_cast<uchar4 *>(out_u8)[p] = c;
            }
            rays_ref += (unsigned long long)r.nrays * (unsigned long long)ns;
            samples += (unsigned long long)ns;
            hits += r.flags & 1u;
            traced += (r.flags & REC_UNTRACED) ? 0 : 1;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        rays_ref += __shfl_down_sync(0xffffffffu, rays_ref, d);
        samples += __shfl_down_sync(0xffffffffu, samples, d);
        hits += __shfl_down_sync(0xffffffffu, hits, d);
        traced += __shfl_down_sync(0xffffffffu, traced, d);
    }
    if (lane == 0) {
        if (rays_ref) atomicAdd(&stats[2], rays_ref);
        if (samples) atomicAdd(&stats[3], samples);
        if (hits) atomicAdd(&stats[4], hits);
        if (traced) atomicAdd(&stats[5], traced);
    }
}

/* FP64 pipe probe: 8 independent chains per thread */
template <bool FUSED> __global__ void k_fp64_probe(double *sink, int iters)
{
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        if (FUSED) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        } else {
            a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c);
            a2 = __dadd_rn(__dmul_rn(a2, m), c); a3 = __dadd_rn(__dmul_rn(a3, m), c);
            a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
            a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;
}

/* ---------------------------------------------------------------------------
 * host runtime
 * ------------------------------------------------------------------------- */

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return ndt_set_error(NDT_B200_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

/* what ndt_b200_sync reads back of a pass: WaveState up to (not including) gstart[] */
#define WAVE_HEAD_BYTES (sizeof(WaveState) - 2 * WAVE_MAX_GEN * sizeof(int))

struct WaveGraph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int valid;
    /* what the nodes were built from: rebuilt when any of it changes */
    Scene sc; WaveArgs a; int np, n_sh, specular, trace_grid, shade_grid;
};

struct ndt_b200_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    /* scene */
    char *d_blob; size_t blob_cap;
    char *d_leafrec; size_t leafrec_cap;   /* LeafRec<npad>[n_leaf_refs] */
    char *d_boxrec; size_t boxrec_cap;     /* BoxRec<npad>[n_leaf_refs] */
    char *d_nrec; size_t nrec_cap;         /* LeafRec<npad>[n_objects - n_items]: the objects nested in hcubes (warp_nested) */
    char *d_nbox; size_t nbox_cap;         /* BoxRec<npad>[n_objects - n_items] */
    ndt_flat_header hdr;
    Scene sc;
    int have_scene;
    /* pools */
    RayRec *d_rec; size_t rec_cap;       /* records */
    void *d_rays; size_t rays_bytes;
    uint32_t *d_mb; size_t mb_bytes;
    HitRec *d_hits; size_t hits_cap;     /* one per record slot */
    char *d_srays; size_t srays_bytes;   /* shadow queries of one batch */
    HitRec *d_shits; size_t shits_cap;
    /* recursive anti-aliasing scratch (ndt_b200_render_aa): grown on demand, kept for the next frame */
    struct { void *p; size_t bytes; } aa_img, aa_fin, aa_u8, aa_samp, aa_xy[2], aa_cells[64];
    int *d_aa_cnt; unsigned long long *d_aa_res;
    double *d_hgeo; size_t hgeo_bytes;   /* hit point + normal per ray of a batch */
    uint32_t *d_qmask; size_t qmask_bytes;
    int *d_wl0, *d_wl1; size_t wl0_bytes, wl1_bytes;          /* walker lists of k_pre (gen.cuh) */
    double2 *d_wt0, *d_wt1; size_t wt0_bytes, wt1_bytes;
    int trace_grid[8][8];                /* cached k_trace occupancy per (stage mode = Scene::any_boxed: 0, 1, 3; NP/2) */
    char *d_ana; size_t ana_bytes;       /* ANAGLYPH_3D: the two eyes' fp64 frames */
    int *d_ctr;                          /* fused path: [0] tail [1] next [2..3] overflow */
    WaveState *d_state;                  /* device-side generation loop (gen.cuh) */
    unsigned long long *d_stats;         /* 8 counters */
    int *h_ctr; unsigned long long *h_stats; /* pinned mirrors */
    char *h_snap;                        /* pinned: WAVE_HEAD_BYTES per pass of the pending render (2) */
    int n_snap;                          /* passes of the pending wavefront render; 0: none / the fused path */
    int n_sh_pending;
    WaveGraph wg;
    int use_graph;                       /* 0: host loop (NDT_B200_NO_GRAPH=1, or the graph could not be built) */
    /* outputs for the host-buffer entry point */
    char *d_out; size_t out_cap;
    double bounce_factor;                /* record pool = n0 * (1 + bounce_factor) + pool_slack */
    int pool_slack;
    int gen_cap_max;                     /* rays per batch of the generation loop */
    uint32_t options;
    ndt_b200_stats last;
    int grid_blocks[16][8];               /* cached occupancy per (CNT + 2 * stage mode, NP/2) */
    int light_type[256];                 /* host copy of lights[].type (which lights can ask a shadow query) */
};

static int grid_for(ndt_b200_ctx *c, int np, bool cnt)
{
    int &g = c->grid_blocks[(cnt ? 1 : 0) + 2 * (c->sc.any_boxed & 7)][np / 2];
    if (g == 0) {
        const NpOps *ops = ndt_np_ops(np);
        g = (ops ? ops->blocks_per_sm(cnt, c->sc.any_boxed) : 1) * c->sm_count;
    }
    return g;
}

static int grow(ndt_b200_ctx *c, void **p, size_t *cap, size_t want_bytes)
{
    if (*cap >= want_bytes) return 0;
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(*p); *p = NULL; *cap = 0;
    CK(cudaMalloc(p, want_bytes));
    *cap = want_bytes;
    return 0;
}

static int trace_grid_for(ndt_b200_ctx *c, int np)
{
    int &g = c->trace_grid[c->sc.any_boxed & 7][np / 2];
    if (g == 0) g = ndt_np_ops(np)->trace_blocks_per_sm(c->sc.any_boxed) * c->sm_count;
    return g;
}

static size_t rayin_bytes(int np) { return (size_t)(2 * np + 2) * sizeof(double); }

extern "C" int ndt_b200_init(int device, ndt_b200_ctx **out)
{
    if (!out) return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_init: NULL ctx pointer");
    *out = NULL;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ndt_set_error(NDT_B200_E_CUDA, "no CUDA device (%s); libndt_b200 has no CPU path",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return ndt_set_error(NDT_B200_E_ARG, "device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    ndt_b200_ctx *c = (ndt_b200_ctx *)calloc(1, sizeof *c);
    if (!c) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaMalloc(&c->d_ctr, 8 * sizeof(int)));
    CK(cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long)));
    CK(cudaMallocHost(&c->h_ctr, 8 * sizeof(int)));
    CK(cudaMallocHost(&c->h_stats, 8 * sizeof(unsigned long long)));
    CK(cudaMalloc(&c->d_state, sizeof(WaveState)));
    CK(cudaMallocHost(&c->h_snap, 2 * WAVE_HEAD_BYTES));
    c->bounce_factor = 6.0;
    c->pool_slack = 65536;
    c->gen_cap_max = 1 << 23;
    {
        const char *e = getenv("NDT_B200_NO_GRAPH");
        c->use_graph = !(e && *e && *e != '0');
    }
    *out = c;
    return 0;
}

extern "C" void ndt_b200_destroy(ndt_b200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_blob); cudaFree(c->d_leafrec); cudaFree(c->d_boxrec); cudaFree(c->d_nrec); cudaFree(c->d_nbox); cudaFree(c->d_rec); cudaFree(c->d_rays); cudaFree(c->d_mb);
    cudaFree(c->d_ana); cudaFree(c->d_hits); cudaFree(c->d_srays); cudaFree(c->d_shits); cudaFree(c->d_hgeo); cudaFree(c->d_qmask);
    cudaFree(c->d_wl0); cudaFree(c->d_wl1); cudaFree(c->d_wt0); cudaFree(c->d_wt1);
    cudaFree(c->aa_img.p); cudaFree(c->aa_fin.p); cudaFree(c->aa_u8.p); cudaFree(c->aa_samp.p);
    cudaFree(c->aa_xy[0].p); cudaFree(c->aa_xy[1].p); cudaFree(c->d_aa_cnt); cudaFree(c->d_aa_res);
    for (int l = 0; l < 64; ++l) cudaFree(c->aa_cells[l].p);
    cudaFree(c->d_ctr); cudaFree(c->d_stats); cudaFree(c->d_out); cudaFree(c->d_state);
    cudaFreeHost(c->h_ctr); cudaFreeHost(c->h_stats); cudaFreeHost(c->h_snap);
    if (c->wg.exec) cudaGraphExecDestroy(c->wg.exec);
    if (c->wg.graph) cudaGraphDestroy(c->wg.graph);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    free(c);
}

extern "C" void *ndt_b200_stream(ndt_b200_ctx *c) { return c ? (void *)c->stream : NULL; }

extern "C" int ndt_b200_set_options(ndt_b200_ctx *c, uint32_t options)
{
    if (!c) return ndt_set_error(NDT_B200_E_ARG, "NULL ctx");
    c->options = options;
    return 0;
}

extern "C" int ndt_b200_set_pool(ndt_b200_ctx *c, double bounce_factor, int slack_records, int rays_per_batch)
{
    if (!c) return ndt_set_error(NDT_B200_E_ARG, "NULL ctx");
    if (!(bounce_factor >= 0.0) || slack_records < 0 || rays_per_batch < 0)
        return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_set_pool: negative argument");
    c->bounce_factor = bounce_factor > 0.0 ? bounce_factor : 6.0;
    c->pool_slack = slack_records;
    c->gen_cap_max = rays_per_batch > 0 ? ((rays_per_batch + 31) & ~31) : (1 << 23);
    return 0;
}

extern "C" int ndt_b200_upload(ndt_b200_ctx *c, const ndt_flat_scene *fs)
{
    if (!c || !fs) return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_upload: NULL argument");
    const ndt_flat_header *h = &fs->h;
    int r = ndt_b200_flat_validate(fs, (size_t)h->total_bytes);
    if (r) return r;
    /* every reject check comes before the context is touched: a failed upload leaves NO scene behind
     * (a later launch then fails with NDT_B200_E_STATE instead of rendering a half-replaced one) */
    c->have_scene = 0;
    if (h->npad < 4 || h->npad > 12 || !ndt_np_ops(h->npad))
        return ndt_set_error(NDT_B200_E_UNSUPPORTED, "%d dimensions: kernels are instantiated for 3..12", h->n);
    if (h->n_lights > 256) return ndt_set_error(NDT_B200_E_UNSUPPORTED, "%d lights (limit 256)", h->n_lights);
    if (h->tree_depth + 2 > KD_STACK)
        return ndt_set_error(NDT_B200_E_UNSUPPORTED, "kd-tree depth %d exceeds the traversal stack (%d)", h->tree_depth, KD_STACK);
    CK(cudaSetDevice(c->device));
    if (c->blob_cap < h->total_bytes) {
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_blob);
        c->d_blob = NULL; c->blob_cap = 0;
        size_t cap = (size_t)h->total_bytes + (size_t)h->total_bytes / 4 + 4096;
        CK(cudaMalloc(&c->d_blob, cap));
        c->blob_cap = cap;
    }
    CK(cudaMemcpyAsync(c->d_blob, fs, (size_t)h->total_bytes, cudaMemcpyHostToDevice, c->stream));
    c->hdr = *h;
    Scene &s = c->sc;
    char *b = c->d_blob;
    s.cam = (const double *)(b + h->off_camera);
    s.aabb = (const double *)(b + h->off_aabb);
    s.bs = (const double *)(b + h->off_bspheres);
    s.geom = (const double *)(b + h->off_geom);
    s.obj = (const ndt_flat_object *)(b + h->off_objects);
    s.nodes = (const ndt_flat_node *)(b + h->off_nodes);
    s.leaf = (const int32_t *)(b + h->off_leaf_refs);
    s.inf = (const int32_t *)(b + h->off_inf);
    s.lights = (const ndt_flat_light *)(b + h->off_lights);
    s.n = h->n; s.n_items = h->n_items; s.n_objects = h->n_objects; s.n_nodes = h->n_nodes;
    s.n_inf = h->n_inf; s.n_lights = h->n_lights;
    s.max_optic_depth = h->max_optic_depth; s.specular = h->specular; s.use_focal = h->use_focal;
    s.width = h->width; s.height = h->height;
    for (int k = 0; k < 4; ++k) s.bg[k] = h->bg[k];
    for (int k = 0; k < 3; ++k) s.ambient[k] = h->ambient[k];
    s.focal_scale = h->focal_scale;
    s.view = h->off_view ? (const double *)(b + h->off_view) : NULL;
    s.cam_type = h->cam_type; s.stereo_mode = h->stereo_mode; s.view_eyes = h->view_eyes;
    s.eye_override = 0; s.cam_dist = h->cam_dist;
    {   /* k_pre inlines trace() over the infinite objects when they are hplanes / cylinders / hcylinders (what the stock plugins make infinite) */
        const ndt_flat_object *ho = (const ndt_flat_object *)((const char *)fs + h->off_objects);
        const int32_t *hinf = (const int32_t *)((const char *)fs + h->off_inf);
        s.inf_hplanes = 1;
        for (int i = 0; i < h->n_inf; ++i) {
            const int t = ho[hinf[i]].type;
            if (t == NDT_T_HPLANE) continue;
            if (t == NDT_T_CYLINDER || t == NDT_T_HCYLINDER) { if (s.inf_hplanes) s.inf_hplanes = 2; }
            else s.inf_hplanes = 0;
        }
    }
    s.any_boxed = 0;
    {   /* k_pack_leaf gives orthotopes with a bounding sphere a box (warp.cuh: box_hit) */
        const ndt_flat_object *ho = (const ndt_flat_object *)((const char *)fs + h->off_objects);
        /* ... and, in boxed scenes, every other primitive the box of its bounding sphere.  Scenes without
         * orthotopes only pay for the second record stream when their leaves are large enough for the culls to
         * matter */
        const bool force = getenv("NDT_B200_FORCE_BOXES") != NULL;      /* tests: the culls on scenes too small to need them */
        for (int i = 0; i < h->n_items && !s.any_boxed; ++i)
            if (ho[i].bs_radius > 0 && (ho[i].type == NDT_T_ORTHOTOPE || h->max_leaf >= 48 || force)) s.any_boxed = 1;
        /* the slab test runs in fp32 with a fixed margin (warp.cuh: box_hit): only for scenes whose
         * coordinates keep its rounding error far below that margin */
        if (s.any_boxed) {
            const double *bb = (const double *)((const char *)fs + h->off_aabb);
            const double *cm = (const double *)((const char *)fs + h->off_camera);
            const double *gg = (const double *)((const char *)fs + h->off_geom);
            const ndt_flat_light *hl2 = (const ndt_flat_light *)((const char *)fs + h->off_lights);
            double ext = 0.0;
            for (int i = 0; i < 2 * h->npad; ++i) if (fabs(bb[i]) > ext) ext = fabs(bb[i]);
            for (int i = 0; i < h->npad; ++i) if (fabs(cm[i]) > ext) ext = fabs(cm[i]);
            for (int l = 0; l < h->n_lights; ++l)
                for (int i = 0; i < h->npad; ++i) if (fabs(gg[hl2[l].vec_off + i]) > ext) ext = fabs(gg[hl2[l].vec_off + i]);
            if (!(ext < 2e4)) s.any_boxed = 0;
        }
    }
    {
        const ndt_flat_light *hl = (const ndt_flat_light *)((const char *)fs + h->off_lights);
        for (int i = 0; i < h->n_lights; ++i) c->light_type[i] = hl[i].type;
    }
    /* the leaf-ordered record stream the warps stage through shared memory (warp.cuh) */
    {
        const size_t recb = (size_t)h->npad * 8 + 48;       /* sizeof(LeafRec<npad>), warp.cuh */
        const size_t need = (size_t)(h->n_leaf_refs > 0 ? h->n_leaf_refs : 1) * recb;
        if (c->leafrec_cap < need) {
            CK(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_leafrec); c->d_leafrec = NULL; c->leafrec_cap = 0;
            CK(cudaMalloc(&c->d_leafrec, need + need / 4));
            c->leafrec_cap = need + need / 4;
        }
        /* the face lists nested in hcubes get the same two streams and a second staging area per warp
         * (warp.cuh: warp_nested) when the scene is boxed and every nested object is an orthotope (hcube.c:33-152
         * creates nothing else); otherwise they stay with the scalar loop */
        s.nrec = s.nbox = NULL;
        const int n_nested = h->n_objects - h->n_items;
        if (s.any_boxed && n_nested > 0 && !getenv("NDT_B200_NO_NESTED_STAGE")) {
            const ndt_flat_object *ho = (const ndt_flat_object *)((const char *)fs + h->off_objects);
            bool all_orthotopes = true;
            for (int i = h->n_items; i < h->n_objects && all_orthotopes; ++i) all_orthotopes = ho[i].type == NDT_T_ORTHOTOPE;
            if (all_orthotopes) {
                if ((r = grow(c, (void **)&c->d_nrec, &c->nrec_cap, (size_t)n_nested * recb))) return r;
                if ((r = grow(c, (void **)&c->d_nbox, &c->nbox_cap, (size_t)n_nested * (size_t)h->npad * 8))) return r;
                s.nrec = c->d_nrec; s.nbox = c->d_nbox;
                s.any_boxed |= 2;
            }
        }
        /* scenes with nested lists or very large leaves run the k_trace instantiation that carries the nested
         * call and the sparse-warp broad phase (warp.cuh); the others keep the smaller kernel */
        if (s.any_boxed && ((s.any_boxed & 2) || h->max_leaf >= 2048)) s.any_boxed |= 4;
        if (h->n_leaf_refs > 0) {
            if ((r = grow(c, (void **)&c->d_boxrec, &c->boxrec_cap, (size_t)h->n_leaf_refs * (size_t)h->npad * 8))) return r;
            ndt_np_ops(h->npad)->pack_leaf(c->stream, s, h->n_leaf_refs, c->d_leafrec, c->d_boxrec, -1);
            CK(cudaGetLastError());
        }
        if (s.any_boxed & 2) {
            ndt_np_ops(h->npad)->pack_leaf(c->stream, s, n_nested, c->d_nrec, c->d_nbox, h->n_items);
            CK(cudaGetLastError());
        }
    }
    c->have_scene = 1;
    return 0;
}

/* grow a pool; running out of device memory is reported like an exhausted pool (NDT_B200_E_OVERFLOW), so
 * that the callers that can split their tile do so instead of failing */
static int grow_pool(ndt_b200_ctx *c, void **p, size_t *cap, size_t want_bytes)
{
    if (*cap >= want_bytes) return 0;
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(*p); *p = NULL; *cap = 0;
    cudaError_t e = cudaMalloc(p, want_bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return ndt_set_error(NDT_B200_E_OVERFLOW, "no device memory for a %zu-byte pool; render a smaller tile", want_bytes);
    }
    if (e != cudaSuccess) return ndt_set_error(NDT_B200_E_CUDA, "cudaMalloc(%zu): %s", want_bytes, cudaGetErrorString(e));
    *cap = want_bytes;
    return 0;
}

/* logical capacity of the record pool for a pass of n0 primary slots */
static size_t pool_records(const ndt_b200_ctx *c, int n0)
{
    size_t want = (size_t)n0 + (size_t)((double)n0 * c->bounce_factor) + (size_t)c->pool_slack;
    if (want > 0x7ffffff0u) want = 0x7ffffff0u;
    return want;
}

static int ensure_pools(ndt_b200_ctx *c, int n0, int np, int grid_threads)
{
    int r;
    size_t cap = c->rec_cap;             /* in records */
    const size_t want = pool_records(c, n0);
    {
        size_t bytes = cap * sizeof(RayRec);
        if ((r = grow_pool(c, (void **)&c->d_rec, &bytes, want * sizeof(RayRec)))) { c->rec_cap = 0; return r; }
        c->rec_cap = bytes / sizeof(RayRec);
    }
    /* word-major with stride = pool capacity (gen.cuh): a full capacity's worth of entries */
    if ((r = grow_pool(c, &c->d_rays, &c->rays_bytes, c->rec_cap * rayin_bytes(np)))) return r;
    size_t words = ((size_t)c->hdr.n_items + 31) / 32;
    if (words == 0) words = 1;
    if ((r = grow_pool(c, (void **)&c->d_mb, &c->mb_bytes, words * (size_t)grid_threads * sizeof(uint32_t)))) return r;
    return 0;
}

/* ANAGLYPH_3D (ndt.c:634-646): the pixel is rendered once per eye and the two colours are
 * mixed into red (left) and blue (right) */
__global__ void k_anaglyph(const double *left, const double *right, int n, double *out_f64, uint8_t *out_u8)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double *l = left + 4 * (size_t)p, *r = right + 4 * (size_t)p;
    const double cr = 0.299 * l[0] + 0.587 * l[1] + 0.114 * l[2];
    const double cb = 0.299 * r[0] + 0.587 * r[1] + 0.114 * r[2];
    if (out_f64) {
        double2 *d = reinterpret_cast<double2 *>(out_f64 + 4 * (size_t)p);
        d[0] = make_double2(cr, 0.0);
        d[1] = make_double2(cb, 1.0);
    }
    if (out_u8) reinterpret_cast<uchar4 *>(out_u8)[p] = make_uchar4(d2c(cr), d2c(0.0), d2c(cb), d2c(1.0));
}

static int launch_pass(ndt_b200_ctx *c, int x0, int y0, int tw, int th,
                       void *d_rgba_f64, void *d_rgba_u8, void *d_hit,
                       void *d_obj_id, void *d_inv_depth, bool first, bool last,
                       const double *d_samples = NULL, int n_samples = 0);

extern "C" int ndt_b200_launch_tile(ndt_b200_ctx *c, int x0, int y0, int tw, int th,
                                    void *d_rgba_f64, void *d_rgba_u8, void *d_hit,
                                    void *d_obj_id, void *d_inv_depth)
{
    if (!c) return ndt_set_error(NDT_B200_E_ARG, "NULL ctx");
    if (!c->have_scene) return ndt_set_error(NDT_B200_E_STATE, "ndt_b200_launch_tile before ndt_b200_upload");
    if (c->hdr.stereo_mode != NDT_ANAGLYPH_3D) {
        c->sc.eye_override = 0;
        return launch_pass(c, x0, y0, tw, th, d_rgba_f64, d_rgba_u8, d_hit, d_obj_id, d_inv_depth, true, true);
    }
    const size_t px = (size_t)(tw > 0 ? tw : 0) * (size_t)(th > 0 ? th : 0);
    int r = grow(c, (void **)&c->d_ana, &c->ana_bytes, 2 * px * 32 + 64);
    if (r) return r;
    double *dl = (double *)c->d_ana, *dr = dl + 4 * px;
    c->sc.eye_override = 1;         /* depth and the hit / id buffers come from the left eye (ndt.c:637) */
    r = launch_pass(c, x0, y0, tw, th, dl, NULL, d_hit, d_obj_id, d_inv_depth, true, false);
    if (r) { c->sc.eye_override = 0; return r; }
    const ndt_b200_stats left = c->last;         /* fused path only; the wavefront's passes are summed by ndt_b200_sync */
    c->sc.eye_override = 2;
    r = launch_pass(c, x0, y0, tw, th, dr, NULL, NULL, NULL, NULL, false, true);
    c->sc.eye_override = 0;
    if (r) return r;
    if (c->n_snap == 0) {
        c->last.rays_bounce += left.rays_bounce;
        c->last.launches += left.launches + 1;
        if (left.generations > c->last.generations) c->last.generations = left.generations;
    }
    k_anaglyph<<<(unsigned)((px + 255) / 256), 256, 0, c->stream>>>(dl, dr, (int)px, (double *)d_rgba_f64, (uint8_t *)d_rgba_u8);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, c->stream));
    return 0;
}

/* grids of the per-query kernels: enough CTAs to fill the GPU, the warps draw their work */
static int light_grid(const ndt_b200_ctx *c) { return c->sm_count * 6; }
static int libm_grid(const ndt_b200_ctx *c) { return c->sm_count * 8; }

/* ---- the CUDA graph of one pass (see the header of this file) ------------------------------------ */
static void wave_graph_drop(ndt_b200_ctx *c)
{
    if (c->wg.exec) cudaGraphExecDestroy(c->wg.exec);
    if (c->wg.graph) cudaGraphDestroy(c->wg.graph);
    memset(&c->wg, 0, sizeof c->wg);
}

static cudaError_t add_kernel(cudaGraph_t g, cudaGraphNode_t *node, cudaGraphNode_t *dep, const void *fn,
                              int blocks, int threads, size_t smem, void **args)
{
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof kp);
    kp.func = (void *)fn;
    kp.gridDim = dim3((unsigned)blocks); kp.blockDim = dim3((unsigned)threads);
    kp.sharedMemBytes = (unsigned)smem;
    kp.kernelParams = args;
    return cudaGraphAddKernelNode(node, g, dep, dep ? 1 : 0, &kp);
}

static int wave_graph_build(ndt_b200_ctx *c, int np, const WaveArgs &a, int n_sh, int trace_grid, int shade_grid)
{
    const NpOps *ops = ndt_np_ops(np);
    WaveGraph &w = c->wg;
    wave_graph_drop(c);
    memcpy(&w.sc, &c->sc, sizeof w.sc); memcpy(&w.a, &a, sizeof w.a); w.np = np; w.n_sh = n_sh; w.specular = c->hdr.specular;
    w.trace_grid = trace_grid; w.shade_grid = shade_grid;
#define GK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        fprintf(stderr, "ndt_b200: %s: %s -- falling back to the host-side generation loop\n", #call, cudaGetErrorString(e_)); \
        cudaGetLastError(); wave_graph_drop(c); c->use_graph = 0; return 1; } } while (0)
    GK(cudaGraphCreate(&w.graph, 0));
    cudaGraphConditionalHandle h1, h2;
    GK(cudaGraphConditionalHandleCreate(&h1, w.graph, 1, cudaGraphCondAssignDefault));     /* generation 0 always runs */
    GK(cudaGraphConditionalHandleCreate(&h2, w.graph, 0, cudaGraphCondAssignDefault));
    /* loop 1: the generations */
    cudaGraphNodeParams p1 = { cudaGraphNodeTypeConditional };
    p1.type = cudaGraphNodeTypeConditional;
    p1.conditional.handle = h1; p1.conditional.type = cudaGraphCondTypeWhile; p1.conditional.size = 1;
    cudaGraphNode_t loop1, loop2, pre, fin, n_prev, n_cur;
    GK(cudaGraphAddNode(&loop1, w.graph, NULL, 0, &p1));
    cudaGraph_t b1 = p1.conditional.phGraph_out[0];
    Scene sc = c->sc;
    WaveArgs wa = a;
    const size_t smem = ops->trace_smem_bytes(sc.any_boxed);
    void *targs[] = { &sc, &wa };
    const int pre_grid = ops->pre_grid(c->sm_count);
    GK(add_kernel(b1, &n_prev, NULL, ops->pre_fn(0), pre_grid, BLOCK, 0, targs));
    GK(add_kernel(b1, &n_cur, &n_prev, ops->trace_fn(0, sc.any_boxed), trace_grid, BLOCK, smem, targs)); n_prev = n_cur;
    GK(add_kernel(b1, &n_cur, &n_prev, ops->shade_fn(0), shade_grid, BLOCK, 0, targs)); n_prev = n_cur;
    if (n_sh > 0) {
        GK(add_kernel(b1, &n_cur, &n_prev, ops->pre_fn(1), pre_grid, BLOCK, 0, targs)); n_prev = n_cur;
        GK(add_kernel(b1, &n_cur, &n_prev, ops->trace_fn(1, sc.any_boxed), trace_grid, BLOCK, smem, targs)); n_prev = n_cur;
        GK(add_kernel(b1, &n_cur, &n_prev, ops->light_fn(), light_grid(c), BLOCK, 0, targs)); n_prev = n_cur;
        int spec = c->hdr.specular, aux_word = np;
        void *largs[] = { &wa, &spec, &aux_word };
        GK(add_kernel(b1, &n_cur, &n_prev, (const void *)k_libm, libm_grid(c), 256, 0, largs)); n_prev = n_cur;
    }
    GK(add_kernel(b1, &n_cur, &n_prev, ops->shade_fn(1), shade_grid, BLOCK, 0, targs)); n_prev = n_cur;
    WaveState *st = c->d_state;
    int cap = a.cap, gen_cap = a.gen_cap, use_h = 1;
    {
        void *args[] = { &st, &cap, &gen_cap, &h1, &use_h };
        GK(add_kernel(b1, &n_cur, &n_prev, (const void *)k_next_gen, 1, 32, 0, args));
    }
    /* the fold */
    {
        void *args[] = { &st, &h2, &use_h };
        GK(add_kernel(w.graph, &pre, &loop1, (const void *)k_pre_resolve, 1, 32, 0, args));
    }
    cudaGraphNodeParams p2 = { cudaGraphNodeTypeConditional };
    p2.type = cudaGraphNodeTypeConditional;
    p2.conditional.handle = h2; p2.conditional.type = cudaGraphCondTypeWhile; p2.conditional.size = 1;
    GK(cudaGraphAddNode(&loop2, w.graph, &pre, 1, &p2));
    cudaGraph_t b2 = p2.conditional.phGraph_out[0];
    double2 *rec = a.rec;
    int specular = c->hdr.specular;
    unsigned long long *stats = c->d_stats;
    {
        void *args[] = { &rec, &cap, &st, &specular };
        GK(add_kernel(b2, &n_prev, NULL, (const void *)k_resolve_dev, c->sm_count * 4, 256, 0, args));
        void *args2[] = { &st, &h2, &use_h };
        GK(add_kernel(b2, &n_cur, &n_prev, (const void *)k_resolve_next, 1, 32, 0, args2));
    }
    {
        void *args[] = { &rec, &cap, &st, &specular, &stats };
        GK(add_kernel(w.graph, &fin, &loop2, (const void *)k_finish_dev, c->sm_count * 8, 256, 0, args));
    }
    GK(cudaGraphInstantiate(&w.exec, w.graph, 0));
#undef GK
    w.valid = 1;
    return 0;
}

/* One pass over a tile of the frame, or -- d_samples != NULL -- over an explicit list of n_samples
 * pixel-space positions (their colours go to d_rgba_f64[n_samples][4]).  Wavefront path: nothing here
 * waits for the device; failures inside the pass surface in ndt_b200_sync. */
static int launch_pass(ndt_b200_ctx *c, int x0, int y0, int tw, int th,
                       void *d_rgba_f64, void *d_rgba_u8, void *d_hit,
                       void *d_obj_id, void *d_inv_depth, bool first, bool last,
                       const double *d_samples, int n_samples)
{
    if (d_samples) { x0 = 0; y0 = 0; tw = n_samples; th = 1; }
    if (!c->have_scene) return ndt_set_error(NDT_B200_E_STATE, "ndt_b200_launch_tile before ndt_b200_upload");
    const ndt_flat_header &h = c->hdr;
    if (!d_samples && (tw <= 0 || th <= 0 || x0 < 0 || y0 < 0 || x0 + tw > h.width || y0 + th > h.height))
        return ndt_set_error(NDT_B200_E_ARG, "tile %dx%d+%d+%d outside the %dx%d frame", tw, th, x0, y0, h.width, h.height);
    if (d_samples && n_samples <= 0) return ndt_set_error(NDT_B200_E_ARG, "empty sample list");
    CK(cudaSetDevice(c->device));
    const int np = h.npad;
    const bool cnt = (c->options & NDT_B200_OPT_COUNT_FLOPS) != 0;
    const int bpr = d_samples ? 0 : (tw + 7) / 8, bprows = (th + 3) / 4;
    const long long n0ll = d_samples ? (long long)n_samples : (long long)bpr * bprows * 32;
    if (n0ll > 0x3fffffff) return ndt_set_error(NDT_B200_E_ARG, "tile too large; render in smaller tiles");
    const int n0 = (int)n0ll;
    const bool wave = d_samples || (!cnt && !(c->options & NDT_B200_OPT_FUSED));
    const int full_grid = wave ? trace_grid_for(c, np) : grid_for(c, np, cnt);
    int r = ensure_pools(c, n0, np, full_grid * BLOCK);
    if (r) return r;
    const uint32_t mb_words = (uint32_t)((h.n_items + 31) / 32) ? (uint32_t)((h.n_items + 31) / 32) : 1u;
    uint32_t mb_shift = 0; while ((mb_words >> mb_shift) >= 64) ++mb_shift;
    cudaStream_t st = c->stream;
    const NpOps *ops = ndt_np_ops(np);

    if (wave) {
        /* lights that can ask for a shadow query */
        int n_sh = 0;
        for (int i = 0; i < h.n_lights; ++i) n_sh += c->light_type[i] != NDT_L_AMBIENT;
        /* rays per batch: a whole generation (a ray has at most two children, and only generation 1 of a
         * frame of glass comes near 2 n0) unless that exceeds the cap */
        int gen_cap = n0 > (1 << 29) ? (1 << 30) : ((2 * n0 + 127) & ~127);
        if (gen_cap < 32768) gen_cap = 32768;
        if (gen_cap > c->gen_cap_max) gen_cap = c->gen_cap_max & ~31;
        if (gen_cap < 32) gen_cap = 32;
        const size_t cap = pool_records(c, n0);
        size_t bytes = c->hits_cap * sizeof(HitRec);
        if ((r = grow_pool(c, (void **)&c->d_hits, &bytes, c->rec_cap * sizeof(HitRec)))) { c->hits_cap = 0; return r; }
        c->hits_cap = bytes / sizeof(HitRec);
        /* every ray of a batch may ask one shadow query per non-ambient light; answers are indexed [ray * n_lights + light] */
        const size_t scap = (size_t)gen_cap * (size_t)(n_sh > 0 ? n_sh : 1);
        const size_t nans = (size_t)gen_cap * (size_t)(h.n_lights > 0 ? h.n_lights : 1);
        if (scap > 0x7ffffff0u || nans > 0x7ffffff0u)
            return ndt_set_error(NDT_B200_E_OVERFLOW, "shadow queue too large; render a smaller tile");
        if ((r = grow_pool(c, (void **)&c->d_srays, &c->srays_bytes, scap * rayin_bytes(np)))) return r;
        bytes = c->shits_cap * sizeof(HitRec);
        if ((r = grow_pool(c, (void **)&c->d_shits, &bytes, nans * sizeof(HitRec)))) { c->shits_cap = 0; return r; }
        c->shits_cap = bytes / sizeof(HitRec);
        const uint32_t mw = (uint32_t)(h.n_lights + 1 + 31) / 32;
        if ((r = grow_pool(c, (void **)&c->d_hgeo, &c->hgeo_bytes, (size_t)gen_cap * 2 * np * sizeof(double)))) return r;
        if ((r = grow_pool(c, (void **)&c->d_qmask, &c->qmask_bytes, (size_t)gen_cap * mw * sizeof(uint32_t)))) return r;
        if ((r = grow_pool(c, (void **)&c->d_wl0, &c->wl0_bytes, (size_t)gen_cap * sizeof(int)))) return r;
        if ((r = grow_pool(c, (void **)&c->d_wl1, &c->wl1_bytes, scap * sizeof(int)))) return r;
        if ((r = grow_pool(c, (void **)&c->d_wt0, &c->wt0_bytes, (size_t)gen_cap * sizeof(double2)))) return r;
        if ((r = grow_pool(c, (void **)&c->d_wt1, &c->wt1_bytes, scap * sizeof(double2)))) return r;

        WaveArgs a;
        memset(&a, 0, sizeof a);
        a.cap = (int)cap; a.gen_cap = gen_cap; a.scap = (int)scap;
        a.rec = (double2 *)c->d_rec; a.rays = (double2 *)c->d_rays; a.hits = (double2 *)c->d_hits;
        a.srays = (double2 *)c->d_srays; a.shits = (double2 *)c->d_shits;
        a.hgeo = (double2 *)c->d_hgeo; a.qmask = c->d_qmask; a.mw = mw;
        a.nl_eff = (uint32_t)(h.n_lights > 0 ? h.n_lights : 1);
        a.st = c->d_state; a.stats = c->d_stats;
        a.mb_bits = c->d_mb; a.mb_stride = (uint32_t)(full_grid * BLOCK);
        a.mb_words = mb_words; a.mb_shift = mb_shift;
        a.leafrec = c->d_leafrec;
        a.boxrec = c->d_boxrec;
        a.wl0 = c->d_wl0; a.wl1 = c->d_wl1; a.wt0 = c->d_wt0; a.wt1 = c->d_wt1;
        const int shade_grid = ops->shade_grid(c->sm_count, gen_cap);

        WaveBegin wb;
        memset(&wb, 0, sizeof wb);
        wb.n0 = n0; wb.x0 = x0; wb.y0 = y0; wb.tw = tw; wb.th = th; wb.bpr = bpr;
        wb.eye = c->sc.eye_override; wb.first = first ? 1 : 0;
        wb.samples_xy = d_samples;
        wb.out_hit = (uint8_t *)d_hit; wb.out_id = (int32_t *)d_obj_id; wb.out_depth = (double *)d_inv_depth;
        wb.out_f64 = (double *)d_rgba_f64; wb.out_u8 = (uint8_t *)d_rgba_u8;
        if (first) c->n_snap = 0;
        if (c->n_snap >= 2) return ndt_set_error(NDT_B200_E_STATE, "more than two passes pending");
        k_begin<<<1, 32, 0, st>>>(c->d_state, wb, gen_cap, c->d_stats);
        if (first) CK(cudaEventRecord(c->ev0, st));

        bool graphed = false;
        if (c->use_graph) {
            WaveGraph &w = c->wg;
            Scene key_sc;
            memcpy(&key_sc, &c->sc, sizeof key_sc);
            key_sc.eye_override = 0;        /* travels in WaveState */
            const bool same = w.valid && w.np == np && w.n_sh == n_sh && w.specular == h.specular &&
                              w.trace_grid == full_grid && w.shade_grid == shade_grid &&
                              memcmp(&w.a, &a, sizeof a) == 0 && memcmp(&w.sc, &key_sc, sizeof key_sc) == 0;
            if (!same) {
                CK(cudaStreamSynchronize(st));       /* the old exec may still be running */
                const int eo = c->sc.eye_override;
                c->sc.eye_override = 0;
                const int br = wave_graph_build(c, np, a, n_sh, full_grid, shade_grid);
                c->sc.eye_override = eo;
                if (br < 0) return br;
            }
            if (c->wg.valid) {
                CK(cudaGraphLaunch(c->wg.exec, st));
                graphed = true;
            }
        }
        if (!graphed) {
            /* the same kernels from a host loop: one read-back of `cont` per batch */
            WaveState *hs = (WaveState *)(c->h_snap + (size_t)c->n_snap * WAVE_HEAD_BYTES);
            cudaGraphConditionalHandle nohandle = 0;
            int guard = 0;
            const bool trace_gens = getenv("NDT_B200_TRACE_GENS") != NULL;     /* the size of every generation, on stderr */
            do {
                const int pre_grid = ops->pre_grid(c->sm_count);
                ops->pre(0, pre_grid, st, c->sc, a);
                ops->trace(0, full_grid, st, c->sc, a);
                ops->shade(0, shade_grid, st, c->sc, a);
                if (n_sh > 0) {
                    ops->pre(1, pre_grid, st, c->sc, a);
                    ops->trace(1, full_grid, st, c->sc, a);
                    ops->light(light_grid(c), st, c->sc, a);
                    k_libm<<<libm_grid(c), 256, 0, st>>>(a, h.specular, np);
                }
                ops->shade(1, shade_grid, st, c->sc, a);
                k_next_gen<<<1, 32, 0, st>>>(c->d_state, a.cap, a.gen_cap, nohandle, 0);
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(hs, c->d_state, WAVE_HEAD_BYTES, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if (trace_gens) fprintf(stderr, "ndt_b200: batch %d done -> next: generation %d, slots [%d, %d), cont %d\n",
                                        hs->iters, hs->gen, hs->start, hs->start + hs->count, hs->cont);
            } while (hs->cont && ++guard < (1 << 20));
            k_pre_resolve<<<1, 32, 0, st>>>(c->d_state, nohandle, 0);
            const int ngen = hs->fail ? 0 : hs->ngen;
            for (int g = ngen - 1; g >= 1; --g) {
                k_resolve_dev<<<c->sm_count * 4, 256, 0, st>>>(a.rec, a.cap, c->d_state, h.specular);
                k_resolve_next<<<1, 32, 0, st>>>(c->d_state, nohandle, 0);
            }
            k_finish_dev<<<c->sm_count * 8, 256, 0, st>>>(a.rec, a.cap, c->d_state, h.specular, c->d_stats);
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(c->h_snap + (size_t)c->n_snap * WAVE_HEAD_BYTES, c->d_state, WAVE_HEAD_BYTES,
                           cudaMemcpyDeviceToHost, st));
        ++c->n_snap;
        c->n_sh_pending = n_sh;
        if (last) {
            CK(cudaEventRecord(c->ev1, st));
            CK(cudaMemcpyAsync(c->h_stats, c->d_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        }
        return 0;
    }

    /* the fused kernel (one launch per generation, host loop): A/B measurements and the counting build */
    c->n_snap = 0;
    c->h_ctr[0] = n0; c->h_ctr[1] = 0; c->h_ctr[2] = 0; c->h_ctr[3] = 0; c->h_ctr[4] = 0; c->h_ctr[5] = 0;
    CK(cudaMemcpyAsync(c->d_ctr, c->h_ctr, 6 * sizeof(int), cudaMemcpyHostToDevice, st));
    if (first) {
        CK(cudaMemsetAsync(c->d_stats, 0, 8 * sizeof(unsigned long long), st));
        CK(cudaEventRecord(c->ev0, st));
    }
    int gstart[1024], gcount[1024], ngen = 0;
    int start = 0, count = n0;
    uint64_t launches = 0;
    GenArgs a;
    memset(&a, 0, sizeof a);
    a.n0 = n0; a.cap = (int)c->rec_cap;
    a.x0 = x0; a.y0 = y0; a.tw = tw; a.th = th; a.bpr = bpr;
    a.rec = c->d_rec; a.rays = c->d_rays;
    a.tail = c->d_ctr; a.next = c->d_ctr + 1; a.overflow = c->d_ctr + 2;
    a.stats = c->d_stats;
    a.out_hit = (uint8_t *)d_hit; a.out_id = (int32_t *)d_obj_id; a.out_depth = (double *)d_inv_depth;
    a.mb_bits = c->d_mb; a.mb_stride = (uint32_t)(full_grid * BLOCK);
    a.mb_words = mb_words; a.mb_shift = mb_shift;
    a.leafrec = c->d_leafrec;
    a.boxrec = c->d_boxrec;
    while (count > 0) {
        if (ngen >= 1024) return ndt_set_error(NDT_B200_E_OVERFLOW, "more than 1024 bounce generations");
        gstart[ngen] = start; gcount[ngen] = count;
        a.gen = ngen; a.start = start; a.count = count;
        if (ngen > 0) CK(cudaMemsetAsync(c->d_ctr + 1, 0, sizeof(int), st));
        int blocks = (count + BLOCK - 1) / BLOCK;
        if (blocks > full_grid) blocks = full_grid;
        ops->generation(cnt, blocks, st, c->sc, a);
        CK(cudaGetLastError());
        ++launches; ++ngen;
        CK(cudaMemcpyAsync(c->h_ctr, c->d_ctr, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (c->h_ctr[3] == 2) return ndt_set_error(NDT_B200_E_CUDA, "leaf staging copy timed out (mbarrier never completed)");
        if (c->h_ctr[3]) return ndt_set_error(NDT_B200_E_OVERFLOW, "kd traversal stack overflow");
        if (c->h_ctr[2]) return ndt_set_error(NDT_B200_E_OVERFLOW, "ray pool exhausted (%zu records); render a smaller tile", c->rec_cap);
        int tail = c->h_ctr[0];
        start += count;
        count = tail - start;
    }
    for (int g = ngen - 1; g >= 1; --g) {
        k_resolve<<<(gcount[g] + 255) / 256, 256, 0, st>>>(c->d_rec, gstart[g], gcount[g], h.specular);
        ++launches;
    }
    k_finish<<<(tw * th + 255) / 256, 256, 0, st>>>(c->d_rec, tw, th, bpr, h.specular,
                                                   (double *)d_rgba_f64, (uint8_t *)d_rgba_u8, c->d_stats);
    ++launches;
    CK(cudaGetLastError());
    if (last) {
        CK(cudaEventRecord(c->ev1, st));
        CK(cudaMemcpyAsync(c->h_stats, c->d_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    }

    memset(&c->last, 0, sizeof c->last);
    c->last.rays_primary = (uint64_t)tw * th;
    c->last.rays_bounce = (uint64_t)(start - n0);
    c->last.generations = (uint32_t)ngen;
    c->last.launches = launches;
    return 0;
}

extern "C" int ndt_b200_sync(ndt_b200_ctx *c)
{
    if (!c) return ndt_set_error(NDT_B200_E_ARG, "NULL ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->n_snap > 0) {
        /* the wavefront: what the device-side loop did, pass by pass */
        const int passes = c->n_snap;
        c->n_snap = 0;
        memset(&c->last, 0, sizeof c->last);
        for (int p = 0; p < passes; ++p) {
            const WaveState *hs = (const WaveState *)(c->h_snap + (size_t)p * WAVE_HEAD_BYTES);
            const int pool = hs->fail & 0xff, kd = hs->fail >> 8;
            if (kd == 2) return ndt_set_error(NDT_B200_E_CUDA, "leaf staging copy timed out (mbarrier never completed)");
            if (kd) return ndt_set_error(NDT_B200_E_OVERFLOW, "kd traversal stack overflow");
            if (pool == 3) return ndt_set_error(NDT_B200_E_OVERFLOW, "more than %d bounce generations", WAVE_MAX_GEN);
            if (pool == 2) return ndt_set_error(NDT_B200_E_OVERFLOW, "shadow queue exhausted; render a smaller tile");
            if (pool) return ndt_set_error(NDT_B200_E_OVERFLOW, "ray pool exhausted; render a smaller tile");
            if (hs->cont) return ndt_set_error(NDT_B200_E_CUDA, "the generation loop did not finish");
            c->last.rays_bounce += (uint64_t)(hs->tail - hs->n0);
            if ((uint32_t)hs->ngen > c->last.generations) c->last.generations = (uint32_t)hs->ngen;
            /* k_begin, 5 or 9 kernels per batch, k_pre_resolve, 2 per folded generation, k_finish */
            c->last.launches += 1 + (uint64_t)hs->iters * (c->n_sh_pending > 0 ? 9 : 5) + 1 +
                                2 * (uint64_t)(hs->ngen > 1 ? hs->ngen - 1 : 0) + 1;
        }
        if (passes == 2) c->last.launches += 1;      /* k_anaglyph */
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last.device_ms = ms;
    c->last.rays_primary = c->h_stats[5];     /* pixels that were traced (all but HIDEF_3D's blanking rows), per eye */
    c->last.rays_shadow = c->h_stats[0];
    c->last.flops = c->h_stats[1];
    c->last.rays_ref = c->h_stats[2];
    c->last.samples = c->h_stats[3];
    c->last.rays_unique = c->last.rays_primary + c->last.rays_bounce + c->last.rays_shadow;
    return 0;
}

extern "C" int ndt_b200_last_stats(ndt_b200_ctx *c, ndt_b200_stats *s)
{
    if (!c || !s) return ndt_set_error(NDT_B200_E_ARG, "NULL argument");
    *s = c->last;
    return 0;
}

static void stats_add(ndt_b200_stats *acc, const ndt_b200_stats *s)
{
    acc->rays_primary += s->rays_primary; acc->rays_bounce += s->rays_bounce;
    acc->rays_shadow += s->rays_shadow; acc->rays_unique += s->rays_unique;
    acc->rays_ref += s->rays_ref; acc->samples += s->samples; acc->flops += s->flops;
    acc->launches += s->launches;
    if (s->generations > acc->generations) acc->generations = s->generations;
    acc->device_ms += s->device_ms;
}

/* One tile into HOST buffers.  A tile whose ray trees do not fit the record pool (or whose pools do not fit
 * the device) is rendered as two half-height tiles; a single row that still overflows gets a larger pool for
 * that one retry.  bounce_factor itself is never left changed. */
static int render_rows(ndt_b200_ctx *c, int x0, int y0, int tw, int th,
                       double *f64, uint8_t *u8, uint8_t *hit, int32_t *id, double *dep,
                       ndt_b200_stats *acc, int depth)
{
    const size_t px = (size_t)tw * th;
    const size_t need = px * (32 + 4 + 1 + 4 + 8) + 256;
    int r = grow_pool(c, (void **)&c->d_out, &c->out_cap, need);
    if (!r) {
        char *b = c->d_out;
        double *d_f64 = (double *)b;              b += px * 32;
        double *d_dep = (double *)b;              b += px * 8;
        int32_t *d_id = (int32_t *)b;             b += px * 4;
        uint8_t *d_u8 = (uint8_t *)b;             b += px * 4;
        uint8_t *d_hit = (uint8_t *)b;
        r = ndt_b200_launch_tile(c, x0, y0, tw, th, f64 ? d_f64 : NULL, u8 ? d_u8 : NULL,
                                 hit ? d_hit : NULL, id ? d_id : NULL, dep ? d_dep : NULL);
        if (!r) {
            cudaStream_t st = c->stream;
            if (f64) CK(cudaMemcpyAsync(f64, d_f64, px * 32, cudaMemcpyDeviceToHost, st));
            if (u8)  CK(cudaMemcpyAsync(u8, d_u8, px * 4, cudaMemcpyDeviceToHost, st));
            if (hit) CK(cudaMemcpyAsync(hit, d_hit, px, cudaMemcpyDeviceToHost, st));
            if (id)  CK(cudaMemcpyAsync(id, d_id, px * 4, cudaMemcpyDeviceToHost, st));
            if (dep) CK(cudaMemcpyAsync(dep, d_dep, px * 8, cudaMemcpyDeviceToHost, st));
            r = ndt_b200_sync(c);           /* the device-side loop reports an exhausted pool here */
        }
    }
    if (r == NDT_B200_E_OVERFLOW && depth < 16) {
        cudaStreamSynchronize(c->stream);
        c->n_snap = 0;
        if (th >= 2) {       /* rows are contiguous in a tile-row-major buffer: split along y */
            int h1 = th / 2;
            size_t o1 = (size_t)tw * h1;
            r = render_rows(c, x0, y0, tw, h1, f64, u8, hit, id, dep, acc, depth + 1);
            if (r) return r;
            return render_rows(c, x0, y0 + h1, tw, th - h1, f64 ? f64 + 4 * o1 : NULL, u8 ? u8 + 4 * o1 : NULL,
                               hit ? hit + o1 : NULL, id ? id + o1 : NULL, dep ? dep + o1 : NULL, acc, depth + 1);
        }
        if (c->bounce_factor < 1e5) {
            const double keep = c->bounce_factor;
            const int keep_slack = c->pool_slack;
            c->bounce_factor = keep * 4.0 + 4.0;
            if (c->pool_slack < 4096) c->pool_slack = 4096;
            r = render_rows(c, x0, y0, tw, th, f64, u8, hit, id, dep, acc, depth + 1);
            c->bounce_factor = keep;
            c->pool_slack = keep_slack;
            return r;
        }
    }
    if (r) return r;
    stats_add(acc, &c->last);
    return 0;
}

extern "C" int ndt_b200_render_tile(ndt_b200_ctx *c, int x0, int y0, int tw, int th,
                                    double *rgba_f64, uint8_t *rgba_u8, uint8_t *hit,
                                    int32_t *obj_id, double *inv_depth, ndt_b200_stats *stats)
{
    if (!c) return ndt_set_error(NDT_B200_E_ARG, "NULL ctx");
    ndt_b200_stats acc;
    memset(&acc, 0, sizeof acc);
    int r = render_rows(c, x0, y0, tw, th, rgba_f64, rgba_u8, hit, obj_id, inv_depth, &acc, 0);
    if (r) return r;
    c->last = acc;
    if (stats) *stats = acc;
    return 0;
}

extern "C" int ndt_b200_trace_rays(ndt_b200_ctx *c, int n_rays, const double *origins, const double *dirs,
                                   const double *dist_limits, int32_t *found, int32_t *obj_id,
                                   double *t, double *hit, double *normal)
{
    if (!c || n_rays < 0 || !origins || !dirs || !found || !obj_id || !t || !hit || !normal)
        return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_trace_rays: NULL argument");
    if (!c->have_scene) return ndt_set_error(NDT_B200_E_STATE, "ndt_b200_trace_rays before ndt_b200_upload");
    if (n_rays == 0) return 0;
    CK(cudaSetDevice(c->device));
    const ndt_flat_header &h = c->hdr;
    const int np = h.npad, n = h.n;
    int blocks = (n_rays + BLOCK - 1) / BLOCK;
    if (blocks > c->sm_count * 2) blocks = c->sm_count * 2;
    int r = ensure_pools(c, 32, np, blocks * BLOCK);
    if (r) return r;
    const size_t vb = (size_t)n_rays * n * sizeof(double);
    const size_t total = 4 * vb + (size_t)n_rays * (2 * sizeof(double) + 2 * sizeof(int32_t)) + 256;
    char *d = NULL;
    CK(cudaMalloc(&d, total));
    double *d_o = (double *)d, *d_v = d_o + (size_t)n_rays * n, *d_hit = d_v + (size_t)n_rays * n,
           *d_nrm = d_hit + (size_t)n_rays * n, *d_lim = d_nrm + (size_t)n_rays * n, *d_t = d_lim + n_rays;
    int32_t *d_found = (int32_t *)(d_t + n_rays), *d_id = d_found + n_rays;
    cudaStream_t st = c->stream;
    cudaError_t e = cudaMemcpyAsync(d_o, origins, vb, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_v, dirs, vb, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && dist_limits) e = cudaMemcpyAsync(d_lim, dist_limits, n_rays * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_ctr, 0, 4 * sizeof(int), st);
    uint32_t words = (uint32_t)((h.n_items + 31) / 32); if (!words) words = 1;
    uint32_t shift = 0; while ((words >> shift) >= 64) ++shift;
    if (e == cudaSuccess) {
        ndt_np_ops(np)->trace_rays(blocks, st, c->sc, n_rays, d_o, d_v, dist_limits ? d_lim : NULL, d_found, d_id, d_t,
                                   d_hit, d_nrm, c->d_mb, (uint32_t)(blocks * BLOCK), words, shift, c->d_ctr + 2,
                                   c->d_leafrec, c->d_boxrec);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(found, d_found, n_rays * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(obj_id, d_id, n_rays * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t, d_t, n_rays * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hit, d_hit, vb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(normal, d_nrm, vb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_ctr, c->d_ctr, 4 * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return ndt_set_error(NDT_B200_E_CUDA, "ndt_b200_trace_rays: %s", cudaGetErrorString(e));
    if (c->h_ctr[3] == 2) return ndt_set_error(NDT_B200_E_CUDA, "leaf staging copy timed out (mbarrier never completed)");
    if (c->h_ctr[3]) return ndt_set_error(NDT_B200_E_OVERFLOW, "kd traversal stack overflow");
    return 0;
}

/* page-locked host memory for the output buffers of ndt_b200_render_tile: the
 * device->host copies then run at PCIe/NVLink-C2C speed instead of being
 * staged through the driver's bounce buffer */
extern "C" void *ndt_b200_host_alloc(size_t bytes)
{
    void *p = NULL;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        ndt_set_error(NDT_B200_E_NOMEM, "cudaMallocHost(%zu) failed", bytes);
        return NULL;
    }
    return p;
}
extern "C" void ndt_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

__global__ void k_replay_probe(int n, const double *in, double *out, int32_t *ns)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double l[4] = { in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3] };
    double o[4];
    ns[i] = replay_samples(l, o);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = o[k];
}

extern "C" int ndt_b200_replay_samples(ndt_b200_ctx *c, int n, const double *rgba_in, double *rgba_out, int32_t *samples)
{
    if (!c || !rgba_in || !rgba_out || !samples || n < 0) return ndt_set_error(NDT_B200_E_ARG, "bad argument");
    if (n == 0) return 0;
    CK(cudaSetDevice(c->device));
    double *d_in = NULL, *d_out = NULL;
    int32_t *d_ns = NULL;
    cudaError_t e = cudaMalloc(&d_in, (size_t)n * 32);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n * 32);
    if (e == cudaSuccess) e = cudaMalloc(&d_ns, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, rgba_in, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        k_replay_probe<<<(n + 127) / 128, 128, 0, c->stream>>>(n, d_in, d_out, d_ns);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(rgba_out, d_out, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(samples, d_ns, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_ns);
    if (e != cudaSuccess) return ndt_set_error(NDT_B200_E_CUDA, "ndt_b200_replay_samples: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int ndt_b200_fp64_peak(ndt_b200_ctx *c, int fused, double *gflops)
{
    if (!c || !gflops) return ndt_set_error(NDT_B200_E_ARG, "NULL argument");
    CK(cudaSetDevice(c->device));
    double *sink = NULL;
    CK(cudaMalloc(&sink, 64));
    const int iters = 1 << 16, blocks = c->sm_count * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(c->ev0, c->stream));
        if (fused) k_fp64_probe<true><<<blocks, threads, 0, c->stream>>>(sink, iters);
        else k_fp64_probe<false><<<blocks, threads, 0, c->stream>>>(sink, iters);
        CK(cudaEventRecord(c->ev1, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(sink);
    /* per iteration and thread: 8 chains x 2 flops (mul+add, fused or not) */
    double flops = (double)blocks * threads * (double)iters * 16.0;
    *gflops = flops / (best * 1e-3) / 1e9;
    return 0;
}

/* ---------------------------------------------------------------------------
 * Recursive (Whitted) anti-aliasing: resample_pixel / recursive_resample (ndt.c:655-733) as a
 * level-synchronous refinement.  A cell is one call of recursive_resample: a square of side `step`
 * (in pixels of the (W+1) x (H+1) sample grid) with its four corner colours.  Level by level:
 *   k_aa_level0     every pixel: image_avg_dbl_pixels4 of its corners (image.c:1175); pixels whose
 *                   variance exceeds aa_diff/255 become level-0 cells and ask for their five samples
 *   (wavefront over the sample list: launch_pass with d_samples)
 *   k_aa_subdivide  every cell: the four sub-pixel averages and variances in the reference's operand
 *                   order (ndt.c:684-702); sub-pixels over the threshold become cells of the next level
 *                   (or, at the last level, are averaged in the CHILD's operand order: ndt.c:663-666)
 *   k_aa_fold       deepest level first: res = average of the four sub-pixels -> the parent's slot
 * The recursion of the reference is depth first, but every value only depends on its own subtree, so
 * the order of evaluation is free; the order of OPERANDS inside every average is kept.
 * ------------------------------------------------------------------------- */
struct AaCell {
    double x, y, step;
    int32_t parent;          /* level 0: pixel index j*W+i; deeper: cell index in the previous level */
    int32_t slot;            /* which sub-pixel of the parent this cell refines */
    double corner[4][4];     /* p1..p4 RGBA, recursive_resample's argument order */
    double sp[4][4];         /* sp1..sp4 (ndt.c:680): sub-pixel colours */
};

__device__ __forceinline__ void aa_avg4(const double *p1, const double *p2, const double *p3, const double *p4,
                                        double *avg, double *var)          /* image.c:1175-1197 */
{
    for (int k = 0; k < 4; ++k) avg[k] = (p1[k] + p2[k] + p3[k] + p4[k]) / 4;
    if (var) {
        double v = 0;
        for (int k = 0; k < 4; ++k)
            v += fabs(avg[k] - p1[k]) + fabs(avg[k] - p2[k]) + fabs(avg[k] - p3[k]) + fabs(avg[k] - p4[k]);
        *var = v;
    }
}

/* the five samples recursive_resample renders for a cell: centre, top middle, left, right, bottom (ndt.c:668-676) */
__device__ __forceinline__ void aa_emit_samples(double *xy, double x, double y, double step)
{
    const double hs = step / 2;
    xy[0] = x + hs;   xy[1] = y + hs;
    xy[2] = x + hs;   xy[3] = y;
    xy[4] = x;        xy[5] = y + hs;
    xy[6] = x + step; xy[7] = y + hs;
    xy[8] = x + hs;   xy[9] = y + step;
}

__global__ void k_aa_level0(const double *img, int W, int H, double thr, int subdivide,
                            double *fin, AaCell *cells, int *n_cells, double *samples_xy,
                            unsigned long long *resampled)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int i = p % W, j = p / W;
    const double *p1 = img + 4 * ((size_t)(W + 1) * j + i), *p2 = p1 + 4;
    const double *p3 = img + 4 * ((size_t)(W + 1) * (j + 1) + i), *p4 = p3 + 4;
    double avg[4], var;
    aa_avg4(p1, p2, p3, p4, avg, &var);                 /* resample_pixel, ndt.c:719-723 */
    for (int k = 0; k < 4; ++k) fin[4 * (size_t)p + k] = avg[k];
    if (!(var > thr)) return;
    atomicAdd(resampled, 1ull);
    if (!subdivide) return;                             /* recursive_resample returns the same average (ndt.c:663-666) */
    const int c = atomicAdd(n_cells, 1);
    AaCell *cell = cells + c;
    cell->x = i; cell->y = j; cell->step = 1.0;
    cell->parent = p; cell->slot = 0;
    for (int k = 0; k < 4; ++k) {
        cell->corner[0][k] = p1[k]; cell->corner[1][k] = p2[k];
        cell->corner[2][k] = p3[k]; cell->corner[3][k] = p4[k];
    }
    aa_emit_samples(samples_xy + 10 * (size_t)c, i, j, 1.0);
}

__global__ void k_aa_subdivide(AaCell *cells, int n, const double *samp, double thr, int next_terminal,
                               AaCell *next, int *n_next, double *next_xy)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    AaCell *cell = cells + c;
    const double *p1 = cell->corner[0], *p2 = cell->corner[1], *p3 = cell->corner[2], *p4 = cell->corner[3];
    const double *p5 = samp + 20 * (size_t)c, *p6 = p5 + 4, *p7 = p5 + 8, *p8 = p5 + 12, *p9 = p5 + 16;
    const double x = cell->x, y = cell->y, hs = cell->step / 2;
    /* operands of the sub-pixel average (ndt.c:685,690,695,700) and of the recursive call (:687,692,697,702) */
    const double *av[4][4] = { { p1, p6, p7, p5 }, { p2, p6, p8, p5 }, { p3, p9, p7, p5 }, { p4, p9, p8, p5 } };
    const double *rc[4][4] = { { p1, p6, p7, p5 }, { p6, p2, p5, p8 }, { p7, p5, p3, p9 }, { p5, p8, p9, p4 } };
    const double ox[4] = { x, x + hs, x, x + hs }, oy[4] = { y, y, y + hs, y + hs };
    for (int k = 0; k < 4; ++k) {
        double var;
        aa_avg4(av[k][0], av[k][1], av[k][2], av[k][3], cell->sp[k], &var);
        if (!(var > thr)) continue;
        if (next_terminal) {
            /* the recursive call returns at once with the average in ITS operand order (ndt.c:663-666) */
            aa_avg4(rc[k][0], rc[k][1], rc[k][2], rc[k][3], cell->sp[k], NULL);
            continue;
        }
        const int d = atomicAdd(n_next, 1);
        AaCell *ch = next + d;
        ch->x = ox[k]; ch->y = oy[k]; ch->step = hs;
        ch->parent = c; ch->slot = k;
        for (int q = 0; q < 4; ++q)
            for (int e = 0; e < 4; ++e) ch->corner[q][e] = rc[k][q][e];
        aa_emit_samples(next_xy + 10 * (size_t)d, ox[k], oy[k], hs);
    }
}

__global__ void k_aa_fold(const AaCell *cells, int n, AaCell *parents, double *fin)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const AaCell *cell = cells + c;
    double res[4];
    aa_avg4(cell->sp[0], cell->sp[1], cell->sp[2], cell->sp[3], res, NULL);     /* ndt.c:704 */
    double *dst = parents ? parents[cell->parent].sp[cell->slot] : fin + 4 * (size_t)cell->parent;
    for (int k = 0; k < 4; ++k) dst[k] = res[k];
}

/* dbl_image_set_pixel into the 8-bit actual_img (ndt.c:777, image.c:126-146, image.h:36-39) */
__global__ void k_aa_store(const double *fin, int n, uint8_t *u8)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double *f = fin + 4 * (size_t)p;
    reinterpret_cast<uchar4 *>(u8)[p] = make_uchar4(d2c(f[0]), d2c(f[1]), d2c(f[2]), d2c(f[3]));
}

/* "simply copy img to actual_img" (ndt.c:1089-1100) when aa_depth < 0 or aa_diff >= 256 */
__global__ void k_aa_copy(const double *img, int W, int H, double *fin)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= W * H) return;
    const int i = p % W, j = p / W;
    const double *s = img + 4 * ((size_t)(W + 1) * j + i);
    for (int k = 0; k < 4; ++k) fin[4 * (size_t)p + k] = s[k];
}

static int aa_terminal(int aa_depth, double step)      /* ndt.c:663 */
{
    return aa_depth <= 0 || step < 1.0 / (2 << (aa_depth - 1));
}

/* colours of n samples; a list whose ray trees do not fit the record pool is split in two (like
 * render_rows does with tiles), a tiny one gets a larger pool */
static int aa_render_samples(ndt_b200_ctx *c, const double *d_xy, int n, double *d_samp, ndt_b200_stats *acc, int depth)
{
    c->sc.eye_override = 0;
    int r = launch_pass(c, 0, 0, 0, 0, d_samp, NULL, NULL, NULL, NULL, true, true, d_xy, n);
    if (!r) r = ndt_b200_sync(c);
    if (r == NDT_B200_E_OVERFLOW && depth < 24) {
        cudaStreamSynchronize(c->stream);
        c->n_snap = 0;
        if (n >= 64) {
            const int h1 = n / 2;
            if ((r = aa_render_samples(c, d_xy, h1, d_samp, acc, depth + 1))) return r;
            return aa_render_samples(c, d_xy + 2 * (size_t)h1, n - h1, d_samp + 4 * (size_t)h1, acc, depth + 1);
        }
        if (c->bounce_factor < 1e5) {
            const double keep = c->bounce_factor;
            const int keep_slack = c->pool_slack;
            c->bounce_factor = keep * 4.0 + 4.0;
            if (c->pool_slack < 4096) c->pool_slack = 4096;
            r = aa_render_samples(c, d_xy, n, d_samp, acc, depth + 1);
            c->bounce_factor = keep;
            c->pool_slack = keep_slack;
            return r;
        }
    }
    if (r) return r;
    stats_add(acc, &c->last);
    return 0;
}

extern "C" int ndt_b200_render_aa(ndt_b200_ctx *c, int aa_diff, int aa_depth,
                                  uint8_t *rgba_u8, double *rgba_f64, uint64_t *pixels_resampled,
                                  ndt_b200_stats *stats)
{
    if (!c) return ndt_set_error(NDT_B200_E_ARG, "NULL ctx");
    if (!c->have_scene) return ndt_set_error(NDT_B200_E_STATE, "ndt_b200_render_aa before ndt_b200_upload");
    if (!c->hdr.aa_pad)
        return ndt_set_error(NDT_B200_E_STATE, "ndt_b200_render_aa needs a scene from ndt_b200_flatten_aa (the (W+1) x (H+1) sample grid)");
    if (aa_depth > 30) return ndt_set_error(NDT_B200_E_UNSUPPORTED, "aa_depth %d: 2 << (aa_depth-1) overflows (ndt.c:663)", aa_depth);
    CK(cudaSetDevice(c->device));
    const int W = c->hdr.width - 1, H = c->hdr.height - 1;
    const size_t px = (size_t)W * H, gpx = (size_t)(W + 1) * (H + 1);
    cudaStream_t st = c->stream;
    ndt_b200_stats acc;
    memset(&acc, 0, sizeof acc);
    int r = 0;
    /* every buffer of the pass lives in the context and only ever grows: an animation rendered with -a allocates
     * during its first frames and then never again (the per-level cudaMalloc / cudaFree of round 1 cost a device
     * synchronisation each) */
    double *d_img = NULL, *d_fin = NULL;
    uint8_t *d_u8 = NULL;
    int *d_cnt = NULL;
    unsigned long long *d_res = NULL;
    AaCell *lvl_cells[64];
    int lvl_n[64], nlvl = 0;
    double *d_xy = NULL, *d_samp = NULL;
    int xy_cur = 0;
    memset(lvl_cells, 0, sizeof lvl_cells);
#define AA_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        r = ndt_set_error(NDT_B200_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); goto aa_done; } } while (0)
#define AA_GROW(slot, want) do { if ((r = grow_pool(c, &(slot).p, &(slot).bytes, (want)))) goto aa_done; } while (0)
    AA_GROW(c->aa_img, gpx * 32); d_img = (double *)c->aa_img.p;
    AA_GROW(c->aa_fin, px * 32); d_fin = (double *)c->aa_fin.p;
    AA_GROW(c->aa_u8, px * 4); d_u8 = (uint8_t *)c->aa_u8.p;
    if (!c->d_aa_cnt) AA_CK(cudaMalloc(&c->d_aa_cnt, sizeof(int)));
    if (!c->d_aa_res) AA_CK(cudaMalloc(&c->d_aa_res, sizeof(unsigned long long)));
    d_cnt = c->d_aa_cnt; d_res = c->d_aa_res;
    AA_CK(cudaMemsetAsync(d_res, 0, sizeof(unsigned long long), st));

    /* the initial image: one sample per corner (render_lines_thread with width+1, height+1) */
    {
        const double keep = c->bounce_factor;
        for (int attempt = 0; ; ++attempt) {
            r = ndt_b200_launch_tile(c, 0, 0, W + 1, H + 1, d_img, NULL, NULL, NULL, NULL);
            if (!r) r = ndt_b200_sync(c);
            if (r != NDT_B200_E_OVERFLOW || attempt >= 6) break;
            cudaStreamSynchronize(st);
            c->n_snap = 0;
            c->bounce_factor *= 2.0;        /* deep ray trees (glass, mirrors): a larger record pool for this frame */
        }
        c->bounce_factor = keep;
    }
    if (r) goto aa_done;
    stats_add(&acc, &c->last);

    if (!(aa_depth >= 0 && aa_diff < 256)) {
        k_aa_copy<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(d_img, W, H, d_fin);
    } else {
        const double thr = aa_diff / 255.0;
        double step = 1.0;
        AA_GROW(c->aa_cells[0], (px ? px : 1) * sizeof(AaCell)); lvl_cells[0] = (AaCell *)c->aa_cells[0].p;
        AA_GROW(c->aa_xy[0], (px ? px : 1) * 10 * sizeof(double)); d_xy = (double *)c->aa_xy[0].p;
        AA_CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), st));
        k_aa_level0<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(d_img, W, H, thr, !aa_terminal(aa_depth, step), d_fin,
                                                                 lvl_cells[0], d_cnt, d_xy, d_res);
        AA_CK(cudaGetLastError());
        int n = 0;
        AA_CK(cudaMemcpyAsync(&n, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
        AA_CK(cudaStreamSynchronize(st));
        while (n > 0) {
            if (nlvl >= 62) { r = ndt_set_error(NDT_B200_E_OVERFLOW, "more than 62 anti-aliasing levels"); goto aa_done; }
            lvl_n[nlvl] = n;
            if ((size_t)n * 5 > 0x3ffffff0u) { r = ndt_set_error(NDT_B200_E_OVERFLOW, "too many anti-aliasing samples in one level"); goto aa_done; }
            /* this level's samples: one wavefront over the list */
            AA_GROW(c->aa_samp, (size_t)n * 5 * 32); d_samp = (double *)c->aa_samp.p;
            if ((r = aa_render_samples(c, d_xy, n * 5, d_samp, &acc, 0))) goto aa_done;
            const int next_terminal = aa_terminal(aa_depth, step / 2);
            double *d_xy_next = NULL;
            if (!next_terminal) {
                AA_GROW(c->aa_cells[nlvl + 1], (size_t)n * 4 * sizeof(AaCell)); lvl_cells[nlvl + 1] = (AaCell *)c->aa_cells[nlvl + 1].p;
                AA_GROW(c->aa_xy[xy_cur ^ 1], (size_t)n * 4 * 10 * sizeof(double)); d_xy_next = (double *)c->aa_xy[xy_cur ^ 1].p;
            }
            AA_CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), st));
            k_aa_subdivide<<<(n + 127) / 128, 128, 0, st>>>(lvl_cells[nlvl], n, d_samp, thr, next_terminal,
                                                            lvl_cells[nlvl + 1], d_cnt, d_xy_next);
            AA_CK(cudaGetLastError());
            int nn = 0;
            AA_CK(cudaMemcpyAsync(&nn, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
            AA_CK(cudaStreamSynchronize(st));
            d_xy = d_xy_next; xy_cur ^= 1;
            ++nlvl;
            step /= 2;
            n = nn;
        }
        for (int l = nlvl - 1; l >= 0; --l)
            k_aa_fold<<<(lvl_n[l] + 127) / 128, 128, 0, st>>>(lvl_cells[l], lvl_n[l], l ? lvl_cells[l - 1] : NULL, d_fin);
    }
    k_aa_store<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(d_fin, (int)px, d_u8);
    AA_CK(cudaGetLastError());
    if (rgba_u8) AA_CK(cudaMemcpyAsync(rgba_u8, d_u8, px * 4, cudaMemcpyDeviceToHost, st));
    if (rgba_f64) AA_CK(cudaMemcpyAsync(rgba_f64, d_fin, px * 32, cudaMemcpyDeviceToHost, st));
    {
        unsigned long long res = 0;
        AA_CK(cudaMemcpyAsync(&res, d_res, sizeof res, cudaMemcpyDeviceToHost, st));
        AA_CK(cudaStreamSynchronize(st));
        if (pixels_resampled) *pixels_resampled = res;
    }
    acc.launches += 2 + (uint64_t)nlvl * 2;
    c->last = acc;
    if (stats) *stats = acc;
aa_done:
#undef AA_CK
#undef AA_GROW
    cudaStreamSynchronize(st);
    return r;
}

/* drop-in for render_image (ndt.c:900) */
static ndt_b200_ctx *g_render_image_ctx = NULL;     /* one context per process, like the reference's global kdtree */

static ndt_b200_mgpu *g_render_image_mgpu = NULL;

/* NDT_B200_DEVICES=<n>|all: how many GPUs of the box render_image uses (default 1) */
static int render_image_devices(void)
{
    const char *e = getenv("NDT_B200_DEVICES");
    if (!e || !*e) return 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) return 1;
    if (!strcmp(e, "all")) return ndev > 0 ? ndev : 1;
    int n = atoi(e);
    if (n > ndev) n = ndev;
    return n > 0 ? n : 1;
}

/* render_image rescales the camera in place (ndt.c:925-926) and callers rely on it (main() re-aims the camera
 * every frame); done once the frame has been rendered, so that a failed call leaves the scene as it was */
static void rescale_dirx(void *scene, double s)
{
    ndtabi_scene *scn = (ndtabi_scene *)scene;
    const int n = scn->cam.dirX.n, k = (n + 1) / 2;
    for (int i = 0; i < 2 * k; ++i) scn->cam.dirX.v[i] = scn->cam.dirX.v[i] * s;
}

/* image_copy (image.c:255-269) into the caller's image_t: image_init / dbl_image_init memset the struct without
 * freeing anything (image.c:52-80), so neither do we -- an image_t the caller never initialised works here as it
 * does with the reference; `pixels` is malloc'ed like image_set_size does, the caller's image_free releases it */
static void give_image(ndtabi_image *img, int width, int height, int pixel_width, void *pixels)
{
    memset(img, 0, sizeof *img);
    img->width = width; img->height = height; img->pixel_width = pixel_width;
    img->allocated = (int)((size_t)width * height * pixel_width);
    img->pixels = (unsigned char *)pixels;
}

/* render_image with the global recursive_aa set (ndt.c:44): the result is the 8-bit actual_img */
extern "C" int ndt_b200_render_image_aa(void *scene, const void *kdtree, const ndt_b200_host_api *host,
                                        char *name, char *depth_name, int width, int height,
                                        int samples, int stereo_mode, int threads, int aa_diff,
                                        int aa_depth, int max_optic_depth, int specular,
                                        void *img_copy, void *depth_copy)
{
    (void)threads; (void)depth_name; (void)depth_copy;
    if (samples != 1) return ndt_set_error(NDT_B200_E_UNSUPPORTED, "samples=%d: jittered sampling uses drand48 (ndt.c:505-542) and is not on the device path", samples);
    if (stereo_mode != NDT_MONO) return ndt_set_error(NDT_B200_E_UNSUPPORTED, "recursive anti-aliasing in stereo mode %d is not on the device path", stereo_mode);
    int r;
    if (!g_render_image_ctx && (r = ndt_b200_init(0, &g_render_image_ctx))) return r;
    ndt_flat_scene *fs = NULL;
    if ((r = ndt_b200_flatten_aa(scene, kdtree, width, height, max_optic_depth, specular, host, &fs))) return r;
    r = ndt_b200_upload(g_render_image_ctx, fs);
    ndt_b200_free_flat(fs);
    if (r) return r;
    const size_t px = (size_t)width * height;
    uint8_t *u8 = (uint8_t *)calloc(px, 4);
    if (!u8) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    r = ndt_b200_render_aa(g_render_image_ctx, aa_diff, aa_depth, u8, NULL, NULL, NULL);
    if (r) { free(u8); return r; }
    rescale_dirx(scene, width / (double)height);
    /* image_copy(img_copy, actual_img), ndt.c:1124-1127: 8-bit RGBA, only for a named frame */
    if (name && img_copy) { give_image((ndtabi_image *)img_copy, width, height, 4, u8); u8 = NULL; }
    free(u8);
    return 1;
}

extern "C" int ndt_b200_render_image(void *scene, const void *kdtree, const ndt_b200_host_api *host,
                                     char *name, char *depth_name, int width, int height,
                                     int samples, int stereo_mode, int threads, int aa_diff,
                                     int aa_depth, int max_optic_depth, int specular,
                                     void *img_copy, void *depth_copy)
{
    (void)threads; (void)aa_diff; (void)aa_depth;
    if (samples != 1) return ndt_set_error(NDT_B200_E_UNSUPPORTED, "samples=%d: jittered sampling uses drand48 (ndt.c:505-542) and is not on the device path", samples);
    ndt_b200_ctx *&ctx = g_render_image_ctx;
    int r;
    const bool timing = getenv("NDT_B200_TIMING") != NULL;     /* per-stage wall times of the call on stderr */
    struct timespec ts0, ts1, ts2, ts3;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    const int want_devices = render_image_devices();
    if (want_devices > 1) {
        if (!g_render_image_mgpu && (r = ndt_b200_mgpu_init(want_devices, NULL, &g_render_image_mgpu))) return r;
    } else if (!ctx && (r = ndt_b200_init(0, &ctx))) return r;
    ndt_flat_scene *fs = NULL;
    if ((r = ndt_b200_flatten_view(scene, kdtree, width, height, max_optic_depth, specular, stereo_mode, host, &fs))) return r;
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    if (want_devices <= 1) {
        r = ndt_b200_upload(ctx, fs);
        ndt_b200_free_flat(fs);
        fs = NULL;
        if (r) return r;
    }
    /* the copies the reference hands back: the frame for a named image, the depth map for a named depth image
     * (ndt.c:1024-1031); writing the files themselves is the binding's job (codecs stay on the host) */
    ndtabi_image *img = (name && img_copy) ? (ndtabi_image *)img_copy : NULL;
    ndtabi_image *dimg = (depth_name && depth_copy) ? (ndtabi_image *)depth_copy : NULL;
    double *f64 = NULL, *dep = NULL, *dep_rgba = NULL;
    const size_t px = (size_t)width * height;
    f64 = (double *)calloc(px, 32);
    if (dimg) { dep = (double *)calloc(px, 8); dep_rgba = (double *)calloc(px, 32); }
    if (!f64 || (dimg && (!dep || !dep_rgba))) {
        free(f64); free(dep); free(dep_rgba); ndt_b200_free_flat(fs);
        return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    }
    if (want_devices > 1) {
        /* rows of the frame over the GPUs of the box, gathered in f64 / dep (ndt.c:812-820, 1277-1309) */
        r = ndt_b200_mgpu_render_frame(g_render_image_mgpu, fs, 0, f64, NULL, NULL, NULL, dep, NULL);
        ndt_b200_free_flat(fs);
    } else {
        r = ndt_b200_render_tile(ctx, 0, 0, width, height, f64, NULL, NULL, NULL, dep, NULL);
    }
    if (r) { free(f64); free(dep); free(dep_rgba); return r; }
    clock_gettime(CLOCK_MONOTONIC, &ts2);
    rescale_dirx(scene, stereo_mode != NDT_HIDEF_3D ? width / (double)height : width / (double)1080);
    if (img) { give_image(img, width, height, 32, f64); f64 = NULL; }      /* fp64 RGBA, row-major */
    if (dimg) {     /* depth map: r=g=b=1/dist, a=1 (ndt.c:754-756) */
        for (size_t i = 0; i < px; ++i) { dep_rgba[4 * i] = dep_rgba[4 * i + 1] = dep_rgba[4 * i + 2] = dep[i]; dep_rgba[4 * i + 3] = 1.0; }
        give_image(dimg, width, height, 32, dep_rgba);
        dep_rgba = NULL;
    }
    free(f64); free(dep); free(dep_rgba);
    if (timing) {
        clock_gettime(CLOCK_MONOTONIC, &ts3);
#define MS(a, b) (((b).tv_sec - (a).tv_sec) * 1e3 + ((b).tv_nsec - (a).tv_nsec) * 1e-6)
        fprintf(stderr, "ndt_b200_render_image %dx%d on %d GPU(s): init + flatten %.1f ms, upload + alloc + render + read-back %.1f ms "
                "(device %.2f ms), copy-out %.1f ms\n", width, height, want_devices, MS(ts0, ts1), MS(ts1, ts2),
                want_devices <= 1 ? ctx->last.device_ms : 0.0, MS(ts2, ts3));
#undef MS
    }
    return 1;
}
