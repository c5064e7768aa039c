/*
 * gen.cuh -- the kernels that depend on the padded dimension NP (k_pack_leaf,
 * k_generation, k_trace_rays) and the per-NP launch table.  Each NP is
 * instantiated in its own translation unit (np_inst.cu built with -DNDT_NP=N)
 * so the five dimensions compile in parallel; kernels.cu looks the launchers up
 * through ndt_np_ops().
 */
#pragma once
#include <cuda_runtime.h>
#include "warp.cuh"

using namespace ndt;

#ifndef BLOCK
#define BLOCK 128
#endif
#ifndef NDT_MIN_BLOCKS
#define NDT_MIN_BLOCKS 3      /* resident CTAs per SM the register allocation is bounded for */
#endif

struct GenArgs {
    int gen;                 /* 0: rays are generated from pixels */
    int start, count;        /* this generation's slots are [start, start+count) */
    int n0;                  /* slots of generation 0 (tile padded to 8x4 blocks) */
    int cap;                 /* record pool capacity */
    int x0, y0, tw, th, bpr; /* tile, and 8-pixel blocks per tile row */
    RayRec *rec;
    void *rays;              /* RayIn<NP>[cap - n0], slot s lives at rays[s - n0] */
    int *tail;               /* next free slot */
    int *next;               /* work counter of this launch */
    int *overflow;           /* [0] ray pool, [1] kd stack */
    unsigned long long *stats; /* [0] shadow rays [1] flops [2] rays_ref [3] samples [4] hit pixels */
    uint8_t *out_hit;
    int32_t *out_id;
    double *out_depth;
    uint32_t *mb_bits;
    uint32_t mb_stride, mb_words, mb_shift;
    const void *leafrec;     /* LeafRec<NP>[n_leaf_refs], leaf order (k_pack_leaf) */
};

/* leaf_refs[] -> LeafRec stream: one thread per reference, once per uploaded scene */
template <int NP>
__global__ void k_pack_leaf(const Scene sc, int n_refs, LeafRec<NP> *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_refs) return;
    const int id = sc.leaf[i];
    const double *bs = sc.bs + (size_t)id * (NP + 2);
    const ndt_flat_object *fo = sc.obj + id;
    LeafRec<NP> r;
    NDT_UNROLL
    for (int k = 0; k < NP; ++k) r.c[k] = bs[k];
    r.r = bs[NP];
    r.r2 = bs[NP + 1];
    r.id = id;
    int nd = geom_block_doubles(fo->type, fo->n_axes, NP);
    /* ndt_b200_flat_validate guarantees geom_off even and n_axes <= n, so every block fits the staging buffer */
    r.tfa = ((uint32_t)fo->type & 0xfu) | (((uint32_t)fo->flags & 0xfu) << 4) | (((uint32_t)fo->n_axes & 0xffu) << 8) |
            ((uint32_t)((nd + 1) / 2) << 16);
    r.geom_off = fo->geom_off;
    r.report_id = fo->report_id;
    out[i] = r;
}

template <int NP, bool CNT>
__global__ void __launch_bounds__(BLOCK, NDT_MIN_BLOCKS) k_generation(const Scene sc, const GenArgs a)
{
    const int lane = threadIdx.x & 31;
    Mailbox mb;
    mb.bits = a.mb_bits; mb.stride = a.mb_stride;
    mb.slot = blockIdx.x * blockDim.x + threadIdx.x;
    mb.words = a.mb_words; mb.group_shift = a.mb_shift;
    mb.dirty = ~0ull;            /* first clear() wipes the whole column */
    Tally<CNT> tally;
    unsigned long long shadow_total = 0;
    int kd_overflow = 0;
    RayIn<NP> *rays = (RayIn<NP> *)a.rays;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpStage<NP> ws;
    if (!CNT) ws.init(smem_raw + (threadIdx.x >> 5) * warp_smem_bytes<NP>(), a.leafrec, lane);

    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(a.next, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= a.count) break;
        const int r = base + lane;
        bool active = r < a.count;
        double o[NP], v[NP], frac = 1.0;
        int depth = sc.max_optic_depth;
        int tx = 0, ty = 0;
        if (a.gen == 0) {
            const int blk = r >> 5;
            tx = (blk % a.bpr) * 8 + (lane & 7);
            ty = (blk / a.bpr) * 4 + (lane >> 3);
            active = active && tx < a.tw && ty < a.th;
            if (active) primary_ray<NP>(sc, a.x0 + tx, a.y0 + ty, o, v);
        } else if (active) {
            const RayIn<NP> *in = rays + (a.start + r - a.n0);
            NDT_UNROLL
            for (int i = 0; i < NP; ++i) { o[i] = in->o[i]; v[i] = in->v[i]; }
            frac = in->frac;
            depth = in->depth;
        }

        RayRec rec;
        Spawn<NP> sp;
        sp.want_refl = sp.want_refr = 0;
        int p_hit = 0, p_id = -1;
        double p_dist = -1.0;
        uint32_t nsh = 0;
        if (CNT) {
            /* counting build: the scalar driver, whose Tally follows the reference step by step */
            if (active) process_ray<NP, CNT>(sc, mb, o, v, frac, depth, rec, sp, p_hit, p_id, p_dist, nsh, kd_overflow, tally);
        } else {
            process_ray_warp<NP>(sc, ws, mb, active, o, v, frac, depth, rec, sp, p_hit, p_id, p_dist, nsh, kd_overflow);
        }
        if (!CNT && ws.fault) break;     /* warp-uniform (warp.cuh) */
        if (active) {
            rec.nrays = 1u + nsh;
            shadow_total += nsh;
        }

        /* hand out slots of the next generation: two ballots, one atomic per warp */
        const bool q1 = active && sp.want_refl == 1;
        const bool q2 = active && sp.want_refr == 1;
        const unsigned b1 = __ballot_sync(0xffffffffu, q1);
        const unsigned b2 = __ballot_sync(0xffffffffu, q2);
        const int total = __popc(b1) + __popc(b2);
        int wbase = 0;
        if (total > 0) {
            if (lane == 0) wbase = atomicAdd(a.tail, total);
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
        }
        const bool fits = wbase + total <= a.cap;
        if (total > 0 && !fits && lane == 0) atomicExch(a.overflow, 1);
        const unsigned lt = (1u << lane) - 1u;
        if (active) {
            if (sp.want_refl == 2) rec.child_refl = CHILD_BLACK;
            if (sp.want_refr == 2) rec.child_refr = CHILD_BLACK;
            if (q1) {
                const int s = wbase + __popc(b1 & lt);
                rec.child_refl = fits ? s : CHILD_BLACK;
                if (fits) {
                    RayIn<NP> *out = rays + (s - a.n0);
                    NDT_UNROLL
                    for (int i = 0; i < NP; ++i) { out->o[i] = sp.origin[i]; out->v[i] = sp.refl_dir[i]; }
                    out->frac = sp.refl_frac; out->depth = depth - 1; out->pad = 0;
                }
            }
            if (q2) {
                const int s = wbase + __popc(b1) + __popc(b2 & lt);
                rec.child_refr = fits ? s : CHILD_BLACK;
                if (fits) {
                    RayIn<NP> *out = rays + (s - a.n0);
                    NDT_UNROLL
                    for (int i = 0; i < NP; ++i) { out->o[i] = sp.origin[i]; out->v[i] = sp.refr_dir[i]; }
                    out->frac = sp.refr_frac; out->depth = depth - 1; out->pad = 0;
                }
            }
            a.rec[a.start + r] = rec;
            if (a.gen == 0) {
                const size_t p = (size_t)ty * a.tw + tx;
                if (a.out_hit) a.out_hit[p] = (uint8_t)p_hit;
                if (a.out_id) a.out_id[p] = p_id;
                if (a.out_depth) a.out_depth[p] = (p_id >= 0 && p_dist > EPS) ? 1.0 / p_dist : 0.0;
            }
        } else if (a.gen == 0 && r < a.count) {
            /* padding lane of a partial 8x4 block: keep the record defined */
            RayRec z;
            z.clr[0] = z.clr[1] = z.clr[2] = 0.0; z.alpha = 0.0;
            z.h[0] = z.h[1] = z.h[2] = 0.0;
            z.child_refl = z.child_refr = CHILD_NONE; z.nrays = 0; z.flags = 0;
            a.rec[a.start + r] = z;
        }
    }

    /* statistics: one atomic per warp */
    for (int d = 16; d > 0; d >>= 1) shadow_total += __shfl_down_sync(0xffffffffu, shadow_total, d);
    if (lane == 0 && shadow_total) atomicAdd(&a.stats[0], shadow_total);
    if (CNT) {
        unsigned long long f = tally.f;
        for (int d = 16; d > 0; d >>= 1) f += __shfl_down_sync(0xffffffffu, f, d);
        if (lane == 0 && f) atomicAdd(&a.stats[1], f);
    }
    if (kd_overflow) atomicMax(a.overflow + 1, 1);
    if (!CNT && ws.fault) atomicMax(a.overflow + 1, 2);
}

/* trace_kd (object.c:683) for an explicit list of rays: the probe behind
 * ndt_b200_trace_rays, used by the per-primitive known-answer tests */
template <int NP>
__global__ void __launch_bounds__(BLOCK, NDT_MIN_BLOCKS)
k_trace_rays(const Scene sc, int n_rays, const double *o_in, const double *v_in, const double *limits,
             int32_t *found, int32_t *ids, double *ts, double *hits, double *normals,
             uint32_t *mb_bits, uint32_t mb_stride, uint32_t mb_words, uint32_t mb_shift, int *overflow,
             const void *leafrec)
{
    const int lane = threadIdx.x & 31;
    Mailbox mb;
    mb.bits = mb_bits; mb.stride = mb_stride;
    mb.slot = blockIdx.x * blockDim.x + threadIdx.x;
    mb.words = mb_words; mb.group_shift = mb_shift;
    mb.dirty = ~0ull;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpStage<NP> ws;
    ws.init(smem_raw + (threadIdx.x >> 5) * warp_smem_bytes<NP>(), leafrec, lane);
    int ovf = 0;
    /* a warp takes 32 consecutive rays at a time; the loop bound is the same for all of its lanes */
    for (int r0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; r0 < n_rays; r0 += gridDim.x * blockDim.x) {
        const int r = r0 + lane;
        const bool want = r < n_rays;
        double o[NP], v[NP];
        NDT_UNROLL
        for (int i = 0; i < NP; ++i) {
            o[i] = (want && i < sc.n) ? o_in[(size_t)r * sc.n + i] : 0.0;
            v[i] = (want && i < sc.n) ? v_in[(size_t)r * sc.n + i] : 0.0;
        }
        Hit T;
        trace_kd_warp<NP>(sc, ws, mb, want, o, v, (want && limits) ? limits[r] : -1.0, T, ovf, false);
        if (want) {
            double p[NP], nr[NP];
            vzero<NP>(p); vzero<NP>(nr);
            if (T.id >= 0) materialise<NP>(sc, T.win, o, v, p, nr);
            found[r] = T.found;
            ids[r] = T.id;
            ts[r] = T.t;
            for (int i = 0; i < sc.n; ++i) { hits[(size_t)r * sc.n + i] = p[i]; normals[(size_t)r * sc.n + i] = nr[i]; }
        }
    }
    if (ovf) atomicMax(overflow + 1, 1);
    if (ws.fault) atomicMax(overflow + 1, 2);
}


/* launchers of one NP, filled in by np_inst.cu */
struct NpOps {
    int (*blocks_per_sm)(bool cnt);
    void (*generation)(bool cnt, int blocks, cudaStream_t st, const Scene &sc, const GenArgs &a);
    void (*pack_leaf)(cudaStream_t st, const Scene &sc, int n_refs, void *out);
    void (*trace_rays)(int blocks, cudaStream_t st, const Scene &sc, int n_rays, const double *o, const double *v,
                       const double *limits, int32_t *found, int32_t *ids, double *ts, double *hits, double *normals,
                       uint32_t *mb_bits, uint32_t mb_stride, uint32_t mb_words, uint32_t mb_shift, int *overflow,
                       const void *leafrec);
};
const NpOps *ndt_np_ops(int np);     /* NULL: dimension not built */
