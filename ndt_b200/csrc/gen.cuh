/*
 * gen.cuh -- the kernels that depend on the padded dimension NP (k_pack_leaf, the wavefront's k_pre / k_trace /
 * k_shade / k_light, the fused k_generation, the probe k_trace_rays) and the per-NP launch table.  Each NP is
 * instantiated in its own translation unit (np_inst.cu built with -DNDT_NP=N)
 * so the five dimensions compile in parallel; kernels.cu looks the launchers up
 * through ndt_np_ops().
 */
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include "warp.cuh"

using namespace ndt;

#ifndef BLOCK
#define BLOCK 128
#endif
#ifndef NDT_MIN_BLOCKS
#define NDT_MIN_BLOCKS 3      /* resident CTAs per SM the register allocation is bounded for */
#endif
#ifndef NDT_TRACE_MIN_BLOCKS
#define NDT_TRACE_MIN_BLOCKS 2   /* after the box cull k_trace is bound by latency on its own local memory: 255 registers (2 CTA/SM) beat 168 (3) and 128 (4): 3.48 / 3.80 / 4.49 ms on config 2 */
#endif

/* queue entries and records move as 16-byte words: a lane's entry is 144 / 80 bytes away from its
 * neighbour's, so 8-byte accesses touched 32 sectors per instruction and the load/store queue was the
 * stall of k_shade (ncu: lg 14 %) */
template <int NP>
__device__ __forceinline__ void ray_store(RayIn<NP> *out, const double *o, const double *v, double frac, int depth,
                                          int aux = 0)
{
    double2 *d = reinterpret_cast<double2 *>(out);
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) d[i / 2] = make_double2(o[i], o[i + 1]);
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) d[NP / 2 + i / 2] = make_double2(v[i], v[i + 1]);
    d[NP] = make_double2(frac, __hiloint2double(aux, depth));   /* frac | depth, aux (shadow queries: where the answer goes) */
}
template <int NP>
__device__ __forceinline__ void ray_load(const RayIn<NP> *in, double *o, double *v, double &frac, int &depth)
{
    const double2 *d = reinterpret_cast<const double2 *>(in);
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) { const double2 t = d[i / 2]; o[i] = t.x; o[i + 1] = t.y; }
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) { const double2 t = d[NP / 2 + i / 2]; v[i] = t.x; v[i + 1] = t.y; }
    const double2 t = d[NP];
    frac = t.x;
    depth = __double2loint(t.y);
}
template <int NP>
__device__ __forceinline__ int ray_aux(const RayIn<NP> *in)
{
    return __double2hiint(reinterpret_cast<const double2 *>(in)[NP].y);
}
__device__ __forceinline__ void rec_load(RayRec &r, const RayRec *src)
{
    const double2 *s = reinterpret_cast<const double2 *>(src);
    double2 *d = reinterpret_cast<double2 *>(&r);
    NDT_UNROLL
    for (int i = 0; i < (int)(sizeof(RayRec) / 16); ++i) d[i] = s[i];
}
__device__ __forceinline__ void rec_store(RayRec *dst, const RayRec &r)
{
    const double2 *s = reinterpret_cast<const double2 *>(&r);
    double2 *d = reinterpret_cast<double2 *>(dst);
    NDT_UNROLL
    for (int i = 0; i < (int)(sizeof(RayRec) / 16); ++i) d[i] = s[i];
}

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
    const void *boxrec;      /* BoxRec<NP>[n_leaf_refs] */
};

/* The box of everything an orthotope can report as a hit (see box_hit in warp.cuh): p0 + sum s_a b_a, s_a in
 * [-EPS, len_a + EPS], thickened by sqrt(2 EPS) = 0.0142; margins doubled.  Doubles, before the outward rounding. */
template <int NP>
__device__ __forceinline__ void orthotope_reach(const Scene &sc, const ndt_flat_object *fo, double *lo_out, double *hi_out)
{
    const int m = fo->n_axes;
    const double *g = sc.geom + fo->geom_off, *p0 = g, *basis = g + NP, *len = basis + (size_t)m * NP;
    for (int k = 0; k < NP; ++k) { lo_out[k] = 0.0; hi_out[k] = 0.0; }
    for (int k = 0; k < sc.n; ++k) {
        double lo = p0[k], hi = p0[k];
        for (int a = 0; a < m; ++a) {
            const double b = basis[(size_t)a * NP + k];
            const double e0 = -2 * EPS * b, e1 = (len[a] + 2 * EPS) * b;
            lo += e0 < e1 ? e0 : e1;
            hi += e0 < e1 ? e1 : e0;
        }
        const double mg = 0.03 + 1e-9 * (fabs(lo) + fabs(hi));
        lo_out[k] = lo - mg;
        hi_out[k] = hi + mg;
    }
}
/* DIRECTIONAL lights whose shadow direction is (nearly) parallel to the orthotope's flat: |P|^2 < EPS with the
 * intersection's own arithmetic (orthotope.c:170-196); bit L = light L */
template <int NP>
__device__ __forceinline__ uint32_t orthotope_parallel_lights(const Scene &sc, const ndt_flat_object *fo)
{
    const int m = fo->n_axes;
    const double *g = sc.geom + fo->geom_off, *p0 = g, *basis = g + NP, *len = basis + (size_t)m * NP;
    const double *bdb = len + m, *bdp = bdb + m;
    uint32_t mask = 0;
    for (int l = 0; l < sc.n_lights && l < 31; ++l) {
        if (sc.lights[l].type != NDT_L_DIRECTIONAL) continue;
        double v[NP], o[NP], P[NP], Q[NP];
        const double *lv = sc.geom + sc.lights[l].vec_off;
        for (int k = 0; k < NP; ++k) { v[k] = lv[2 * NP + k]; o[k] = 0.0; }
        axes_PQ<NP, LdGlobal>(o, v, p0, basis, bdb, bdp, m, P, Q);
        const double qa = vdot<NP>(P, P);
        if (!(qa >= EPS)) mask |= 1u << l;
    }
    return mask;
}
__device__ __forceinline__ bool hcube_of_orthotopes(const Scene &sc, const ndt_flat_object *fo)
{
    if (fo->child_count <= 0) return false;
    for (int ch = fo->child_begin; ch < fo->child_begin + fo->child_count; ++ch)
        if (sc.obj[ch].type != NDT_T_ORTHOTOPE) return false;
    return true;
}

/* leaf_refs[] -> LeafRec stream: one thread per reference, once per uploaded scene.  id_base >= 0: the
 * records of objects id_base .. id_base + n_refs - 1 instead (the face lists nested in hcubes, warp_nested) */
template <int NP>
__global__ void k_pack_leaf(const Scene sc, int n_refs, LeafRec<NP> *out, BoxRec<NP> *box_out, int id_base)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_refs) return;
    const int id = id_base >= 0 ? id_base + i : sc.leaf[i];
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
    r.boxed = 0;
    r.par_mask = 0x80000000u;
    r.pad[0] = r.pad[1] = 0;
    BoxRec<NP> bx;
    for (int k = 0; k < NP; ++k) { bx.lo[k] = -FLT_MAX; bx.hi[k] = FLT_MAX; }
    if (fo->type == NDT_T_ORTHOTOPE && bs[NP] > 0) {
        double lo[NP], hi[NP];
        orthotope_reach<NP>(sc, fo, lo, hi);
        for (int k = 0; k < sc.n; ++k) {
            /* stored as float, rounded outwards */
            bx.lo[k] = __double2float_rd(lo[k]);
            bx.hi[k] = __double2float_ru(hi[k]);
        }
        r.boxed = 1;
        r.par_mask |= orthotope_parallel_lights<NP>(sc, fo);
    }
    else if (fo->type == NDT_T_HCUBE && bs[NP] > 0 && sc.any_boxed && hcube_of_orthotopes(sc, fo)) {
        /* an hcube reports what one of its faces reports (hcube.c:236-250: trace() over the face list, the cube
         * only replaces the object pointer): the union of the faces' boxes bounds every hit point, and a shadow
         * direction that must not be culled for one face must not be culled for the cube.  The cube's bounding
         * sphere is that of its corners, its box far smaller: of the rays that pass the sphere of a randomly
         * turned 6-cube one in eight meets the box (BASELINE config 3; each of the others cost a pass over
         * the cube's 472 faces). */
        double lo[NP], hi[NP];
        for (int k = 0; k < NP; ++k) { lo[k] = DBL_MAX; hi[k] = -DBL_MAX; }
        uint32_t pm = 0;
        for (int ch = fo->child_begin; ch < fo->child_begin + fo->child_count; ++ch) {
            double cl[NP], chh[NP];
            orthotope_reach<NP>(sc, sc.obj + ch, cl, chh);
            for (int k = 0; k < sc.n; ++k) { lo[k] = cl[k] < lo[k] ? cl[k] : lo[k]; hi[k] = chh[k] > hi[k] ? chh[k] : hi[k]; }
            pm |= orthotope_parallel_lights<NP>(sc, sc.obj + ch);
        }
        for (int k = 0; k < sc.n; ++k) {
            bx.lo[k] = __double2float_rd(lo[k]);
            bx.hi[k] = __double2float_ru(hi[k]);
        }
        r.boxed = 1;
        r.par_mask |= pm;
    }
    else if (bs[NP] > 0 && sc.any_boxed) {
        /* any other primitive with a bounding sphere: the sphere's own box.  A ray that misses it misses the
         * sphere, so the reference's pre-test (bounding.c:52-84) rejects the object before its intersect() is
         * called -- for every kind of query, hence no par_mask.  The margin (0.03) dwarfs what fp64 rounding can
         * move desc by at |coordinates| < 2e4 (ndt_b200_upload).  What it buys: the fp32 slab test and the bundle
         * cull run ahead of the fp64 sphere test for every record of a large leaf (BASELINE config 3: the kd
         * tree cannot separate 10 000 overlapping 6-D objects, a leaf holds them all). */
        const double rr = bs[NP];
        for (int k = 0; k < sc.n; ++k) {
            const double mg = 0.03 + 1e-9 * (fabs(bs[k]) + rr);
            bx.lo[k] = __double2float_rd(bs[k] - rr - mg);
            bx.hi[k] = __double2float_ru(bs[k] + rr + mg);
        }
        r.boxed = 1;
        r.par_mask = 0;
    }
    out[i] = r;
    box_out[i] = bx;
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
    if (!CNT) {
        ws.init(smem_raw + (threadIdx.x >> 5) * warp_smem_bytes<NP>(sc.any_boxed), a.leafrec, sc.any_boxed ? a.boxrec : nullptr, lane);
        ws.init_nested(sc.any_boxed);
    }

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
            if (active) active = primary_ray<NP>(sc, a.x0 + tx, a.y0 + ty, o, v);
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
                    ray_store<NP>(rays + (s - a.n0), sp.origin, sp.refl_dir, sp.refl_frac, depth - 1);
                }
            }
            if (q2) {
                const int s = wbase + __popc(b1) + __popc(b2 & lt);
                rec.child_refr = fits ? s : CHILD_BLACK;
                if (fits) {
                    ray_store<NP>(rays + (s - a.n0), sp.origin, sp.refr_dir, sp.refr_frac, depth - 1);
                }
            }
            rec_store(a.rec + (a.start + r), rec);
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
            z.child_refl = z.child_refr = CHILD_NONE; z.nrays = 0; z.flags = REC_UNTRACED;
            rec_store(a.rec + (a.start + r), z);
            if (tx < a.tw && ty < a.th) {       /* a pixel of the frame that is not traced (HIDEF_3D blanking rows) */
                const size_t p = (size_t)ty * a.tw + tx;
                if (a.out_hit) a.out_hit[p] = 0;
                if (a.out_id) a.out_id[p] = -1;
                if (a.out_depth) a.out_depth[p] = 0.0;
            }
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

/* ---------------------------------------------------------------------------
 * The wavefront: one bounce generation = four launches
 *   k_trace (radiance rays)  nearest hit of every ray of the generation -> HitRec
 *   k_shade<A>               per ray: hit point, normal, side tests (ndt.c:149-169);
 *                            queues one shadow ray per light that needs one
 *   k_trace (shadow rays)    -> HitRec per shadow ray
 *   k_shade<B>               per ray: the light loop again, in the reference's order,
 *                            now with the shadow answers; RayRec; reflection /
 *                            refraction rays appended to the next generation
 * k_trace holds nothing but the walk, the leaf loop and the primitives: measured on B200
 * the fused kernel (k_generation) was bound by instruction fetch -- SM instruction cache
 * 32 KB, hit rate 84 %, GPC instruction-cache requests at 73 % of peak -- because every ray
 * dragged the shading code (acos / pow / sin / asin expansions, 60 KB) through the cache
 * between two visits of the 20 KB leaf loop.  k_shade re-derives hit point and normal from
 * the winner's identity (materialise) instead of carrying them through memory, and runs the
 * light loop twice (A to emit the queries, B to consume the answers): both are a few hundred
 * flops per ray against the thousands of a traversal.
 * ------------------------------------------------------------------------- */
struct alignas(16) HitRec {
    double t;
    int32_t id, win, found, pad;
    double pad2;
};                            /* 32 bytes = two 16-byte words */
__device__ __forceinline__ void hit_store(HitRec *dst, double t, int id, int win, int found)
{
    double2 *d = reinterpret_cast<double2 *>(dst);
    d[0] = make_double2(t, __hiloint2double(win, id));
    d[1] = make_double2(__hiloint2double(0, found), 0.0);
}
__device__ __forceinline__ void hit_load(const HitRec *src, double &t, int &id, int &win, int &found)
{
    const double2 *s = reinterpret_cast<const double2 *>(src);
    const double2 a = s[0], b = s[1];
    t = a.x; id = __double2loint(a.y); win = __double2hiint(a.y); found = __double2loint(b.x);
}

/* ---- queue layout of the wavefront: word-major ------------------------------------------------------------
 * A queue entry (ray, record, hit, hit geometry) is a handful of 16-byte words.  Stored entry after entry, the
 * 32 lanes of a warp reading word c of 32 consecutive entries touch 32 different sectors 80-176 bytes apart:
 * ncu on k_shade (profiles/r02_ncu_k_shade_aos_*): 73 % of the stall samples long_scoreboard on exactly those
 * loads, L1 hit rate 54-61 %, 42 cycles per issued instruction.  Stored word after word -- word c of entry e at
 * base[c * stride + e] -- the same access is 512 contiguous bytes. */
template <int NP>
__device__ __forceinline__ void ray_store_s(double2 *base, size_t stride, size_t e, const double *o, const double *v,
                                            double frac, int depth, int aux = 0)
{
    double2 *d = base + e;
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) d[(size_t)(i / 2) * stride] = make_double2(o[i], o[i + 1]);
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) d[(size_t)(NP / 2 + i / 2) * stride] = make_double2(v[i], v[i + 1]);
    d[(size_t)NP * stride] = make_double2(frac, __hiloint2double(aux, depth));   /* frac | depth, aux (shadow queries: where the answer goes) */
}
template <int NP>
__device__ __forceinline__ void ray_load_s(const double2 *base, size_t stride, size_t e, double *o, double *v,
                                           double &frac, int &depth, int &aux)
{
    const double2 *d = base + e;
    double2 t[NP + 1];
    NDT_UNROLL
    for (int i = 0; i <= NP; ++i) t[i] = d[(size_t)i * stride];          /* all loads in flight before the first use */
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) { o[i] = t[i / 2].x; o[i + 1] = t[i / 2].y; }
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) { v[i] = t[NP / 2 + i / 2].x; v[i + 1] = t[NP / 2 + i / 2].y; }
    frac = t[NP].x;
    depth = __double2loint(t[NP].y);
    aux = __double2hiint(t[NP].y);
}
__device__ __forceinline__ void rec_load_s(RayRec &r, const double2 *base, size_t stride, size_t e)
{
    double2 *d = reinterpret_cast<double2 *>(&r);
    NDT_UNROLL
    for (int i = 0; i < (int)(sizeof(RayRec) / 16); ++i) d[i] = base[(size_t)i * stride + e];
}
__device__ __forceinline__ void rec_store_s(double2 *base, size_t stride, size_t e, const RayRec &r)
{
    const double2 *s = reinterpret_cast<const double2 *>(&r);
    NDT_UNROLL
    for (int i = 0; i < (int)(sizeof(RayRec) / 16); ++i) base[(size_t)i * stride + e] = s[i];
}
__device__ __forceinline__ void hit_store_s(double2 *base, size_t stride, size_t e, double t, int id, int win, int found)
{
    base[e] = make_double2(t, __hiloint2double(win, id));
    base[stride + e] = make_double2(__hiloint2double(0, found), 0.0);
}
__device__ __forceinline__ void hit_load_s(const double2 *base, size_t stride, size_t e, double &t, int &id, int &win, int &found)
{
    const double2 a = base[e], b = base[stride + e];
    t = a.x; id = __double2loint(a.y); win = __double2hiint(a.y); found = __double2loint(b.x);
}

/* ---- the generation loop lives on the device -------------------------------------------------------
 * Everything a launch needs to know about "which rays now" is read from WaveState in device memory, so the
 * host can enqueue a whole frame without knowing how many bounce generations it has: one CUDA graph per
 * frame whose body is a WHILE node (kernels.cu), or -- same kernels -- a host loop that reads `cont` back.
 * A generation larger than gen_cap is worked off in batches (the shadow queue and its answers are sized
 * for one batch).  Three groups of fields, each on its own cache lines: constant for a pass (k_begin),
 * constant for a launch (k_begin / k_next_gen), and the atomics of the running launch. */
constexpr int WAVE_MAX_GEN = 1024;
/* what every kernel of the loop reads and none of them writes: copied to shared memory at the top of each
 * kernel (wave_head_load), because read from global memory where they are used these values cost an L2 round
 * trip per ray (the compiler hoists ordinary loads into registers instead, +40-60 registers) */
struct WaveHead {
    /* the pass: written by k_begin */
    int n0;                  /* slots of generation 0 (tile padded to 8x4 blocks, or the sample count) */
    int x0, y0, tw, th, bpr; /* tile, and 8-pixel blocks per tile row (0: an explicit sample list) */
    int eye;                 /* eye_override of primary_ray (ANAGLYPH_3D passes) */
    int first;
    const double *samples_xy; /* generation 0 from an explicit list of pixel-space positions (ip, jp) instead of the
                                 tile's pixel grid: the sub-pixel samples of the recursive anti-aliasing */
    uint8_t *out_hit;
    int32_t *out_id;
    double *out_depth;
    double *out_f64;
    uint8_t *out_u8;
    /* the launch: written by k_begin and k_next_gen / k_resolve_next */
    int gen;                 /* generation of the current batch; 0: rays are generated from pixels */
    int start, count;        /* the batch: record slots [start, start + count) */
    int gen_start, gen_count; /* the generation the batch belongs to */
    int ngen;                /* finished generations (gstart / gcount filled) */
    int iters;               /* batches worked off */
    int cont;                /* the loop goes on (mirror of the WHILE node's condition, for the host loop) */
    int resolve_g;           /* the generation k_resolve folds next */
    int fail;                /* copy of pool_overflow | kd_fault << 8 when the loop ended */
};
static_assert(sizeof(WaveHead) == 120, "WaveHead is copied as 15 8-byte words");
struct WaveState : WaveHead {
    /* atomics of the running launch, one cache line each: every warp of the GPU hits them */
    alignas(128) int tail;   /* next free record slot */
    alignas(128) int next0;  /* work counter, radiance rays */
    alignas(128) int stail;  /* shadow queue tail */
    alignas(128) int next1;  /* work counter, shadow queries */
    alignas(128) int nextA;  /* work counters of k_shade<A> / k_shade<B> */
    alignas(128) int nextB;
    alignas(128) int nextL;  /* ... of k_light / k_libm */
    alignas(128) int nextM;
    alignas(128) int nextP0; /* ... of k_pre<0> / k_pre<1> */
    alignas(128) int nextP1;
    /* the walker lists k_pre hands to k_trace (WaveArgs::wl0 / wl1): warps whose 32 rays all walk claim 32 entries
     * from the front (wfull, a multiple of 32: their bundle stays one 8x4 pixel block), the others append theirs
     * from the back (wpart) */
    alignas(128) int wfull0;
    alignas(128) int wpart0;
    alignas(128) int wfull1;
    alignas(128) int wpart1;
    alignas(128) int nextR;  /* ... of k_resolve_dev / k_finish_dev */
    alignas(128) int nextF;
    alignas(128) int pool_overflow;  /* 1 record pool, 2 shadow queue, 3 more than WAVE_MAX_GEN generations */
    int kd_fault;            /* 1 traversal stack overflow, 2 staging copy timed out */
    alignas(128) int gstart[WAVE_MAX_GEN];
    int gcount[WAVE_MAX_GEN];
};
/* every thread of the block calls it before anything else */
__device__ __forceinline__ const WaveHead *wave_head_load(const WaveState *st)
{
    __shared__ __align__(16) unsigned long long sh_head[sizeof(WaveHead) / 8];
    if (threadIdx.x < sizeof(WaveHead) / 8)
        sh_head[threadIdx.x] = reinterpret_cast<const unsigned long long *>(static_cast<const WaveHead *>(st))[threadIdx.x];
    __syncthreads();
    return reinterpret_cast<const WaveHead *>(sh_head);
}

struct WaveArgs {            /* constant for the life of a graph: pool pointers and capacities */
    int cap;                 /* record pool capacity */
    int gen_cap;             /* rays per batch (multiple of 32) */
    int scap;                /* shadow queue capacity (gen_cap * non-ambient lights) */
    int pad;
    /* all word-major (see ray_store_s): */
    double2 *rec;            /* RayRec, 5 words, stride cap */
    double2 *rays;           /* RayIn<NP>, NP + 1 words, stride cap; slot s lives at entry s - n0 */
    double2 *hits;           /* HitRec, 2 words, stride cap, by slot */
    double2 *srays;          /* shadow queries, NP + 1 words, stride scap: frac = dist_limit, depth = 1 + light index for
                                the any-hit query of a DIRECTIONAL light, else 0 */
    double2 *shits;          /* [gen_cap * n_lights]: the answer to the shadow query of (ray, light) at [ray * n_lights + light],
                                ray counted from the start of the batch; the query carries that index (ray_aux).  The slot
                                is rewritten in place by k_light (LightGeo) and k_libm (LightTerm); entries without a
                                query are never read */
    double2 *hgeo;           /* NP words, stride gen_cap: hit point and normal of every shaded ray of the batch (k_shade<A>) */
    uint32_t *qmask;         /* [mw][gen_cap]: bit 0 = shaded, bit 1 + it = light it was asked (k_shade<A>) */
    uint32_t mw, nl_eff;     /* nl_eff = max(n_lights, 1) */
    WaveState *st;
    unsigned long long *stats; /* [0] shadow rays [1] flops [2] rays_ref [3] samples [4] hit pixels [5] traced pixels */
    uint32_t *mb_bits;
    uint32_t mb_stride, mb_words, mb_shift, pad2;
    const void *leafrec;
    const void *boxrec;
    /* k_pre -> k_trace: the rays of the batch that walk the tree (index within the batch) and the interval of each
     * inside the root box (tl, tu); [gen_cap] for the batch's own rays, [scap] for its shadow queries */
    int *wl0, *wl1;
    double2 *wt0, *wt1;
};

/* what k_begin gets by value */
struct WaveBegin {
    int n0, x0, y0, tw, th, bpr, eye, first;
    const double *samples_xy;
    uint8_t *out_hit;
    int32_t *out_id;
    double *out_depth;
    double *out_f64;
    uint8_t *out_u8;
};

/* the ray of slot start + r of the current batch */
template <int NP>
__device__ __forceinline__ bool wave_ray(const Scene &sc, const WaveArgs &a, const WaveHead *hd, int gen, int start, int count,
                                         int r, int lane, double *o, double *v, double &frac, int &depth, int &tx, int &ty)
{
    bool active = r < count;
    frac = 1.0;
    depth = sc.max_optic_depth;
    tx = ty = 0;
    if (gen == 0) {
        const double *sxy = hd->samples_xy;
        const int s = start + r;
        if (sxy) {
            if (active) primary_ray_at<NP>(sc, sxy[2 * (size_t)s], sxy[2 * (size_t)s + 1], o, v);
        } else {
            const int blk = s >> 5, bpr = hd->bpr;
            tx = (blk % bpr) * 8 + (s & 7);              /* slot s = pixel (s & 31) of block s >> 5, whichever lane asks */
            ty = (blk / bpr) * 4 + ((s >> 3) & 3);
            active = active && tx < hd->tw && ty < hd->th;
            if (active) active = primary_ray<NP>(sc, hd->x0 + tx, hd->y0 + ty, o, v, hd->eye);
        }
    } else if (active) {
        { int aux_; ray_load_s<NP>(a.rays, (size_t)a.cap, (size_t)(start + r - hd->n0), o, v, frac, depth, aux_); }
    }
    return active;
}

/* MODE 0: the batch's own rays, MODE 1: its shadow queries.
 * k_pre: what every query does before it walks (warp.cuh: pre_walk) -- the infinite objects and the root box --
 * for ALL rays of the batch, 12-32 warps per SM where k_trace has 8.  A ray that does not enter the root box gets
 * its final answer here; a walker gets the state of the walk so far (the hit record holds trace()'s result over the
 * infinite objects, wt its interval in the root box) and an entry in the walker list. */
#ifndef NDT_PRE_MIN_BLOCKS
#define NDT_PRE_MIN_BLOCKS 4
#endif
template <int NP, int MODE>
__global__ void __launch_bounds__(BLOCK, NDT_PRE_MIN_BLOCKS) k_pre(const Scene sc, const WaveArgs a)
{
    const int lane = threadIdx.x & 31;
    WaveState *st = a.st;
    const WaveHead *hd = wave_head_load(st);
    const int gen = hd->gen, start = hd->start;
    int count = hd->count;
    if (MODE == 1) {             /* written by k_shade<A>, complete before this launch started */
        count = *(volatile const int *)&st->stail;
        if (count > a.scap) count = a.scap;
    }
    int *wfull = MODE ? &st->wfull1 : &st->wfull0, *wpart = MODE ? &st->wpart1 : &st->wpart0;
    int *wl = MODE ? a.wl1 : a.wl0;
    double2 *wt = MODE ? a.wt1 : a.wt0;
    const int lcap = MODE ? a.scap : a.gen_cap;
    /* a static stride: what a ray costs here hardly varies (k_trace draws from a counter: there it varies a lot) */
    const int stride = (int)(gridDim.x * blockDim.x);
    for (int base = (int)(blockIdx.x * blockDim.x) + (threadIdx.x & ~31); base < count; base += stride) {
        const int r = base + lane;
        double o[NP], v[NP], limit = -1.0;
        bool want;
        int dir_light = -1, dest = 0;
        if (MODE == 0) {
            double frac; int depth, tx, ty;
            want = wave_ray<NP>(sc, a, hd, gen, start, count, r, lane, o, v, frac, depth, tx, ty);
        } else {
            want = r < count;
            if (want) {
                int code;
                ray_load_s<NP>(a.srays, (size_t)a.scap, (size_t)r, o, v, limit, code, dest);
                dir_light = code - 1;
            }
        }
        PreWalk pw;
        pw.walking = false;
        if (want) pre_walk<NP, true>(sc, o, v, limit, dir_light >= 0, pw);
        const bool walk = want && pw.walking;
        const unsigned wm = __ballot_sync(0xffffffffu, walk);
        int pos = 0;
        if (wm) {
            const int nw = __popc(wm);
            if (lane == 0) pos = nw == 32 ? atomicAdd(wfull, 32) : lcap - (atomicAdd(wpart, nw) + nw);
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(wm & ((1u << lane) - 1u));
        }
        if (want) {
            /* a walker's record is what the walk starts from, everybody else's is final */
            if (MODE == 0) hit_store_s(a.hits, (size_t)a.cap, (size_t)(start + r), pw.md, pw.id, pw.win, pw.ret);
            else hit_store_s(a.shits, (size_t)a.gen_cap * a.nl_eff, (size_t)dest, pw.md, pw.id, pw.win, pw.ret);
            if (walk) {
                wl[pos] = r;
                wt[pos] = make_double2(pw.tl, pw.tu);
            }
        }
    }
}

/* k_trace: the walk, for the rays k_pre listed */
template <int NP, int MODE, bool BIG>
__global__ void __launch_bounds__(BLOCK, NDT_TRACE_MIN_BLOCKS) k_trace(const Scene sc, const WaveArgs a)
{
    const int lane = threadIdx.x & 31;
    WaveState *st = a.st;
    const WaveHead *hd = wave_head_load(st);
    const int gen = hd->gen, start = hd->start;
    int count = hd->count;
    if (MODE == 1) {             /* written by k_shade<A>, complete before this launch started */
        count = *(volatile const int *)&st->stail;
        if (count > a.scap) count = a.scap;
    }
    /* the walker list of k_pre, complete before this launch started */
    const int n_full = *(volatile const int *)(MODE ? &st->wfull1 : &st->wfull0);
    const int n_part = *(volatile const int *)(MODE ? &st->wpart1 : &st->wpart0);
    const int n_walk = n_full + n_part;
    const int lcap = MODE ? a.scap : a.gen_cap;
    const int *wl = MODE ? a.wl1 : a.wl0;
    const double2 *wt = MODE ? a.wt1 : a.wt0;
    if ((int)(blockIdx.x * blockDim.x) >= n_walk && blockIdx.x > 0) return;   /* the grid is sized for a full batch */
    Mailbox mb;
    mb.bits = a.mb_bits; mb.stride = a.mb_stride;
    mb.slot = blockIdx.x * blockDim.x + threadIdx.x;
    mb.words = a.mb_words; mb.group_shift = a.mb_shift;
    mb.dirty = ~0ull;            /* first clear() wipes the whole column */
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpStage<NP> ws;
    ws.init(smem_raw + (threadIdx.x >> 5) * warp_smem_bytes<NP>(sc.any_boxed), a.leafrec, sc.any_boxed ? a.boxrec : nullptr, lane);
    ws.init_nested(sc.any_boxed);
    int kd_overflow = 0;
    int *next = MODE ? &st->next1 : &st->next0;

    /* Rays per draw of the work counter: one bundle of 32.  128 per draw measured 7-10 % slower (the heavy bundles
     * of a silhouette end up in one warp); spreading a launch that has fewer rays than the resident warps could
     * take thinner (16 ... 1 rays per draw) measured 6-17 % slower on every workload (one more trip to the counter
     * per draw, the same leaves staged by more warps): profiles/r02_experiments.md. */
    constexpr int draw = 32;
    while (true) {
        /* (fetching the counter one batch ahead was measured slower: profiles/r01_experiments.md) */
        int base = 0;
        if (lane == 0) base = atomicAdd(next, draw);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n_walk) break;
        const int i = base + lane;
        const bool want = i < n_walk;
        const int pos = i < n_full ? i : lcap - n_part + (i - n_full);
        double o[NP], v[NP], limit = -1.0;
        int dir_light = -1;          /* >= 0: the any-hit query of that DIRECTIONAL light */
        int dest = 0;                /* MODE 1: index of the answer in shits[] */
        int r = 0;
        PreWalk pw;
        pw.md = -1; pw.tl = pw.tu = 0; pw.id = pw.win = -1; pw.ret = 0; pw.walking = want;
        if (want) {
            r = wl[pos];
            const double2 tt = wt[pos];
            pw.tl = tt.x; pw.tu = tt.y;
            if (MODE == 0) {
                double frac; int depth, tx, ty;
                wave_ray<NP>(sc, a, hd, gen, start, count, r, lane, o, v, frac, depth, tx, ty);
                hit_load_s(a.hits, (size_t)a.cap, (size_t)(start + r), pw.md, pw.id, pw.win, pw.ret);
            } else {
                int code;
                ray_load_s<NP>(a.srays, (size_t)a.scap, (size_t)r, o, v, limit, code, dest);
                dir_light = code - 1;
                hit_load_s(a.shits, (size_t)a.gen_cap * a.nl_eff, (size_t)dest, pw.md, pw.id, pw.win, pw.ret);
            }
        }
        Hit T;
        trace_kd_resume<NP, BIG>(sc, ws, mb, want, o, v, limit, pw, T, kd_overflow, dir_light);
        if (ws.fault) break;         /* warp-uniform (warp.cuh) */
        if (want) {
            if (MODE == 0) hit_store_s(a.hits, (size_t)a.cap, (size_t)(start + r), T.t, T.id, T.win, T.found);
            else hit_store_s(a.shits, (size_t)a.gen_cap * a.nl_eff, (size_t)dest, T.t, T.id, T.win, T.found);
        }
    }
    if (kd_overflow) atomicMax(&st->kd_fault, 1);
    if (ws.fault) atomicMax(&st->kd_fault, 2);
}

/* ---- shading: four kernels around the shadow trace ------------------------------------------------------
 *   k_shade<NP,0>  (A, per ray)    hit point + normal of the winner (materialise), stored for the later phases;
 *                                  side tests (ndt.c:149-169); one shadow query per light that needs one; the
 *                                  primary ray's hit / id / depth outputs
 *   k_light<NP>    (per query)     the vector half of the lit term (light_geom, wave.cuh): does the light see
 *                                  the shaded point, acos argument, specular dot product -> three scalars
 *   k_libm         (per query)     acos / cos / pow on those scalars (light_libm): no vectors, ~40 registers,
 *                                  full occupancy -- the libm expansions were the long dependent FP64 chains
 *                                  of the old phase B, run there at 8-12 warps per SM
 *   k_shade<NP,1>  (B, per ray)    colour in the reference's light order from the (scale, rvn) pairs, RayRec,
 *                                  reflection / refraction rays of the next generation
 * All have fixed grids (CUDA graph) and draw 32 rays / queries per warp from a counter in WaveState: the cost
 * of a ray varies from a sky pixel to a glass sphere under three lights, and a static stride left the tail of
 * every launch to the warps that drew the expensive ones (profiles/r02_experiments.md). */
#ifndef NDT_SHADE_LOOP
#define NDT_SHADE_LOOP 1      /* 1: fixed grid, dynamic draw; 0: one thread per ray of a FULL batch, early exit */
#endif

/* the answer slot of a (ray, light) pair goes through three states: HitRec (k_trace<1>), LightGeo (k_light),
 * LightTerm (k_libm); all 32 bytes */
struct alignas(16) LightGeo { double q, ldist2, rvdot; int32_t lit, qok; };
struct alignas(16) LightTerm { double light_scale, rvn; int32_t lit, pad; double pad2; };

template <int NP> __device__ __forceinline__ void hgeo_store_s(double2 *base, size_t stride, size_t e, const double *Hp, const double *Hn)
{
    double2 *d = base + e;
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) d[(size_t)(i / 2) * stride] = make_double2(Hp[i], Hp[i + 1]);
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) d[(size_t)(NP / 2 + i / 2) * stride] = make_double2(Hn[i], Hn[i + 1]);
}
template <int NP> __device__ __forceinline__ void hgeo_load_s(const double2 *base, size_t stride, size_t e, double *Hp, double *Hn)
{
    const double2 *d = base + e;
    double2 t[NP];
    NDT_UNROLL
    for (int i = 0; i < NP; ++i) t[i] = d[(size_t)i * stride];
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) { Hp[i] = t[i / 2].x; Hp[i + 1] = t[i / 2].y; }
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) { Hn[i] = t[NP / 2 + i / 2].x; Hn[i + 1] = t[NP / 2 + i / 2].y; }
}

/* phase A, one ray */
template <int NP>
__device__ __forceinline__ void shade_a_one(const Scene &sc, const WaveArgs &a, WaveState *st, const WaveHead *hd, int r, int lane)
{
    const int gen = hd->gen, start = hd->start, count = hd->count;
    const int nl = sc.n_lights;
    double o[NP], v[NP], frac;
    int depth, tx, ty;
    /* the kernel is bound by the latency of its first loads (ncu: long_scoreboard at the first use of the
     * hit record): start them before the ray is rebuilt */
    Hit T0;
    T0.t = -1; T0.id = -1; T0.win = -1; T0.found = 0;
    if (r < count) hit_load_s(a.hits, (size_t)a.cap, (size_t)(start + r), T0.t, T0.id, T0.win, T0.found);
    const bool active = wave_ray<NP>(sc, a, hd, gen, start, count, r, lane, o, v, frac, depth, tx, ty);
    Tally<false> none;
    Shade<NP> S;
    RayRec rec;
    Spawn<NP> sp;
    shade_init<NP>(S, rec, sp);
    int p_hit = 0, p_id = -1;
    double p_dist = -1.0;
    uint32_t nsh = 0;
    if (active) {
        shade_setup<NP, false>(sc, S, -1, o, v, nsh, none);
        shade_after<NP, false>(sc, S, -1, T0, o, v, rec, p_hit, p_id, p_dist, none);
        if (S.shaded) hgeo_store_s<NP>(a.hgeo, (size_t)a.gen_cap, (size_t)r, S.Hp, S.Hn);
    }
    /* which lights were asked: bit 0 = the ray is shaded at all, bit 1 + it = light it has a query */
    uint32_t mword = (active && S.shaded) ? 1u : 0u;
    uint32_t *mrow = a.qmask + r;            /* word w at mrow[w * gen_cap] */
    int mw_idx = 0;
    if (__ballot_sync(FULL, active && S.shaded)) {
        for (int it = 0; it < nl; ++it) {
            const bool want = active && shade_setup<NP, false>(sc, S, it, o, v, nsh, none);
            /* slots of this light's queries: one ballot, one atomic per warp; queries of one
             * light from neighbouring pixels end up next to each other in the queue */
            const unsigned b = __ballot_sync(FULL, want);
            int wbase = 0;
            if (b) {
                if (lane == 0) wbase = atomicAdd(&st->stail, __popc(b));
                wbase = __shfl_sync(FULL, wbase, 0);
            }
            if (want) {
                const int slot = wbase + __popc(b & ((1u << lane) - 1u));
                if (slot < a.scap) {
                    /* depth = light index + 1 for the any-hit query of a DIRECTIONAL light */
                    ray_store_s<NP>(a.srays, (size_t)a.scap, (size_t)slot, S.ro, S.rv, S.limit,
                                    S.ltype == NDT_L_DIRECTIONAL ? 1 + it : 0, it * a.gen_cap + r);
                } else {
                    atomicExch(&st->pool_overflow, 2);   /* the render fails with NDT_B200_E_OVERFLOW */
                }
                mword |= 1u << ((it + 1) & 31);
            }
            if (((it + 2) & 31) == 0) {         /* the word is full */
                if (r < count) mrow[(size_t)mw_idx * a.gen_cap] = mword;
                ++mw_idx; mword = 0;
            }
        }
    }
    if (r < count) {
        mrow[(size_t)mw_idx * a.gen_cap] = mword;
        for (int w = mw_idx + 1; w < (int)a.mw; ++w) mrow[(size_t)w * a.gen_cap] = 0;
    }
    if (gen == 0 && hd->samples_xy == nullptr && r < count) {
        const int tw = hd->tw, th = hd->th;
        if (tx < tw && ty < th) {
            /* !active: a pixel of the frame that is not traced (HIDEF_3D blanking rows) */
            const size_t p = (size_t)ty * tw + tx;
            uint8_t *oh = hd->out_hit;
            int32_t *oi = hd->out_id;
            double *od = hd->out_depth;
            if (oh) oh[p] = active ? (uint8_t)p_hit : 0;
            if (oi) oi[p] = active ? p_id : -1;
            if (od) od[p] = (active && p_id >= 0 && p_dist > EPS) ? 1.0 / p_dist : 0.0;
        }
    }
}

/* phase B, one ray */
template <int NP>
__device__ __forceinline__ void shade_b_one(const Scene &sc, const WaveArgs &a, WaveState *st, const WaveHead *hd, int r, int lane,
                                            unsigned long long &shadow_total)
{
    const int gen = hd->gen, start = hd->start, count = hd->count;
    const int nl = sc.n_lights;
    double o[NP], v[NP], frac;
    int depth, tx, ty;
    Hit T0;
    T0.t = -1; T0.id = -1; T0.win = -1; T0.found = 0;
    uint32_t mword = 0;
    const uint32_t *mrow = a.qmask + r;
    if (r < count) {
        hit_load_s(a.hits, (size_t)a.cap, (size_t)(start + r), T0.t, T0.id, T0.win, T0.found);
        mword = mrow[0];
    }
    const bool active = wave_ray<NP>(sc, a, hd, gen, start, count, r, lane, o, v, frac, depth, tx, ty);
    Tally<false> none;
    Shade<NP> S;
    RayRec rec;
    Spawn<NP> sp;
    shade_init<NP>(S, rec, sp);
    uint32_t nsh = 0;
    if (active && (mword & 1u)) {
        /* what shade_after(-1) left in S, without the intersection */
        S.shaded = true;
        S.oid = T0.id;
        hgeo_load_s<NP>(a.hgeo, (size_t)a.gen_cap, (size_t)r, S.Hp, S.Hn);
        const ndt_flat_object *fo = sc.obj + S.oid;
        S.hr = NDT_LDG(&fo->rgb[0]); S.hg = NDT_LDG(&fo->rgb[1]); S.hb = NDT_LDG(&fo->rgb[2]);
        rec.h[0] = NDT_LDG(&fo->refl[0]); rec.h[1] = NDT_LDG(&fo->refl[1]); rec.h[2] = NDT_LDG(&fo->refl[2]);
        if (sc.specular) { S.rr = rec.h[0]; S.rg = rec.h[1]; S.rb = rec.h[2]; }
        S.transparent = (NDT_LDG(&fo->flags) & NDT_OF_TRANSPARENT) != 0;
        S.clr0 = S.hr * sc.ambient[0];                               /* ndt.c:88-92 */
        S.clr1 = S.hg * sc.ambient[1];
        S.clr2 = S.hb * sc.ambient[2];
        int mw_idx = 0;
        for (int it = 0; it < nl; ++it) {                            /* ndt.c:103-310, the reference's order */
            const ndt_flat_light *L = sc.lights + it;
            if (NDT_LDG(&L->type) == NDT_L_AMBIENT) {                /* ndt.c:105-111 */
                S.clr0 += S.hr * NDT_LDG(&L->rgb[0]); S.clr1 += S.hg * NDT_LDG(&L->rgb[1]); S.clr2 += S.hb * NDT_LDG(&L->rgb[2]);
            } else if (mword & (1u << ((it + 1) & 31))) {
                ++nsh;
                const size_t e = (size_t)it * a.gen_cap + r;
                const double2 t0 = a.shits[e], t1 = a.shits[(size_t)a.gen_cap * a.nl_eff + e];
                if (__double2loint(t1.y)) light_apply<NP>(sc, S, L, t0.x, t0.y);      /* lit (k_light), scale and rvn (k_libm) */
            }
            if (((it + 2) & 31) == 0) mword = mrow[(size_t)(++mw_idx) * a.gen_cap];
        }
    }
    if (active) {
        shade_finish<NP, false>(sc, S, v, frac, depth, rec, sp, none);
        rec.nrays = 1u + nsh;
    }
    /* hand out slots of the next generation: two ballots, one atomic per warp */
    const bool q1 = active && sp.want_refl == 1;
    const bool q2 = active && sp.want_refr == 1;
    const unsigned b1 = __ballot_sync(FULL, q1);
    const unsigned b2 = __ballot_sync(FULL, q2);
    const int total = __popc(b1) + __popc(b2);
    int wbase = 0;
    if (total > 0) {
        if (lane == 0) wbase = atomicAdd(&st->tail, total);
        wbase = __shfl_sync(FULL, wbase, 0);
    }
    const bool fits = wbase + total <= a.cap;
    if (total > 0 && !fits && lane == 0) atomicExch(&st->pool_overflow, 1);
    const unsigned lt = (1u << lane) - 1u;
    const int n0 = hd->n0;
    if (active) {
        if (sp.want_refl == 2) rec.child_refl = CHILD_BLACK;
        if (sp.want_refr == 2) rec.child_refr = CHILD_BLACK;
        if (q1) {
            const int s = wbase + __popc(b1 & lt);
            rec.child_refl = fits ? s : CHILD_BLACK;
            if (fits) ray_store_s<NP>(a.rays, (size_t)a.cap, (size_t)(s - n0), sp.origin, sp.refl_dir, sp.refl_frac, depth - 1);
        }
        if (q2) {
            const int s = wbase + __popc(b1) + __popc(b2 & lt);
            rec.child_refr = fits ? s : CHILD_BLACK;
            if (fits) ray_store_s<NP>(a.rays, (size_t)a.cap, (size_t)(s - n0), sp.origin, sp.refr_dir, sp.refr_frac, depth - 1);
        }
        rec_store_s(a.rec, (size_t)a.cap, (size_t)(start + r), rec);
    } else if (gen == 0 && r < count) {
        /* padding lane of a partial 8x4 block: keep the record defined */
        RayRec z;
        z.clr[0] = z.clr[1] = z.clr[2] = 0.0; z.alpha = 0.0;
        z.h[0] = z.h[1] = z.h[2] = 0.0;
        z.child_refl = z.child_refr = CHILD_NONE; z.nrays = 0; z.flags = REC_UNTRACED;
        rec_store_s(a.rec, (size_t)a.cap, (size_t)(start + r), z);
    }
    shadow_total += active ? nsh : 0;
}

/* Inside the drawing loop ptxas hoists loop-invariant scene loads (lights, materials) out of the loop and
 * keeps them in registers: +40 to +60 registers and one CTA per SM less than the loop-free form (ptxas -v;
 * k_shade<4,B> 3.3 -> 4.6 ms on BASELINE config 1).  The launch bounds pin the occupancy. */
template <int NP, int PHASE> __host__ __device__ constexpr int shade_min_blocks()
{
#ifdef NDT_SHADE_A_BLOCKS
    if (PHASE == 0) return NDT_SHADE_A_BLOCKS;
#endif
#ifdef NDT_SHADE_MIN_BLOCKS
    return NDT_SHADE_MIN_BLOCKS;
#else
#ifdef NDT_SHADE_NO_BOUNDS
    return 1;
#else
    return PHASE == 0 ? (NP <= 4 ? 4 : (NP <= 8 ? 3 : 2)) : (NP <= 6 ? 4 : (NP <= 8 ? 3 : 2));
#endif
#endif
}
template <int NP, int PHASE>
__global__ void __launch_bounds__(BLOCK, shade_min_blocks<NP, PHASE>()) k_shade(const Scene sc, const WaveArgs a)
{
    const int lane = threadIdx.x & 31;
    const WaveHead *hd = wave_head_load(a.st);
    const int count = hd->count;
    unsigned long long shadow_total = 0;
#if defined(NDT_SHADE_STATIC)
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r - lane < count; r += gridDim.x * blockDim.x) {
        if (PHASE == 0) shade_a_one<NP>(sc, a, a.st, hd, r, lane);
        else shade_b_one<NP>(sc, a, a.st, hd, r, lane, shadow_total);
    }
#elif NDT_SHADE_LOOP
    /* 128 rays per draw: the round trip of the atomic is exposed (nothing else to do in the warp), one per
     * 32 rays was 10 % of the stall samples */
    int *next = PHASE ? &a.st->nextB : &a.st->nextA;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(next, 128);
        base = __shfl_sync(FULL, base, 0);
        if (base >= count) break;
        for (int k = 0; k < 4 && base + 32 * k < count; ++k) {
            if (PHASE == 0) shade_a_one<NP>(sc, a, a.st, hd, base + 32 * k + lane, lane);
            else shade_b_one<NP>(sc, a, a.st, hd, base + 32 * k + lane, lane, shadow_total);
        }
    }
#else
    {
        const int r = blockIdx.x * blockDim.x + threadIdx.x;
        if (r - lane >= count) return;
        if (PHASE == 0) shade_a_one<NP>(sc, a, a.st, hd, r, lane);
        else shade_b_one<NP>(sc, a, a.st, hd, r, lane, shadow_total);
    }
#endif
    if (PHASE == 0) return;
    for (int d = 16; d > 0; d >>= 1) shadow_total += __shfl_down_sync(FULL, shadow_total, d);
    if (lane == 0 && shadow_total) atomicAdd(&a.stats[0], shadow_total);
}

/* the vector half of the lit term, one shadow query per thread */
template <int NP>
__global__ void __launch_bounds__(BLOCK, 3) k_light(const Scene sc, const WaveArgs a)
{
    const int lane = threadIdx.x & 31;
    WaveState *st = a.st;
    const WaveHead *hd = wave_head_load(st);
    const int gen = hd->gen, start = hd->start, count = hd->count;
    int nq = *(volatile const int *)&st->stail;
    if (nq > a.scap) nq = a.scap;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&st->nextL, 128);
        base = __shfl_sync(FULL, base, 0);
        if (base >= nq) break;
        for (int k = 0; k < 4; ++k) {
        const int j = base + 32 * k + lane;
        if (j >= nq) continue;
        double ro[NP], rv[NP], limit;
        int code;
        int aux;
        ray_load_s<NP>(a.srays, (size_t)a.scap, (size_t)j, ro, rv, limit, code, aux);
        const int it = aux / a.gen_cap, r = aux - it * a.gen_cap;
        const size_t sstride = (size_t)a.gen_cap * a.nl_eff;
        Hit T;
        hit_load_s(a.shits, sstride, (size_t)aux, T.t, T.id, T.win, T.found);
        const ndt_flat_light *L = sc.lights + it;
        const int ltype = NDT_LDG(&L->type);
        LightGeo out;
        out.q = 0.0; out.ldist2 = 1.0; out.rvdot = 0.0; out.lit = 0; out.qok = 0;
        /* the cheap rejections first (ndt.c:217, 241): most shadowed pairs stop here */
        double t0; int oid, owin, ofound;
        hit_load_s(a.hits, (size_t)a.cap, (size_t)(start + r), t0, oid, owin, ofound);
        const bool maybe = ltype == NDT_L_DIRECTIONAL ? !T.found : (T.found && T.id == oid);
        if (maybe) {
            double Hp[NP], Hn[NP], o[NP], v[NP], frac, lv[NP];
            int depth, tx, ty;
            hgeo_load_s<NP>(a.hgeo, (size_t)a.gen_cap, (size_t)r, Hp, Hn);
            /* the ray's direction (ndt.c:296 needs -look): rebuilt like every other kernel does */
            wave_ray<NP>(sc, a, hd, gen, start, count, r, (start + r) & 31, o, v, frac, depth, tx, ty);
            if (ltype != NDT_L_DIRECTIONAL) {
                /* light_vec and ldist2 as shade_setup computed them (ndt.c:194-197): the query carries the light's
                 * position (ro) and the unit vector (rv); |Hp - lgt_pos|^2 is the same subtraction and dot product */
                double d[NP];
                vsub<NP>(Hp, ro, d);
                out.ldist2 = vdot<NP>(d, d);
                vcopy<NP>(lv, rv);
            }
            Tally<false> none;
            double q, rvdot; int qok;
            if (light_geom<NP, false>(sc, L, ltype, oid, Hp, Hn, ro, rv, lv, T, v, q, qok, rvdot, none)) {
                out.q = q; out.qok = qok; out.rvdot = rvdot; out.lit = 1;
            }
        }
        a.shits[aux] = make_double2(out.q, out.ldist2);
        a.shits[sstride + aux] = make_double2(out.rvdot, __hiloint2double(out.qok, out.lit));
        }
    }
}

/* trace_kd (object.c:683) for an explicit list of rays: the probe behind
 * ndt_b200_trace_rays, used by the per-primitive known-answer tests */
template <int NP>
__global__ void __launch_bounds__(BLOCK, NDT_MIN_BLOCKS)
k_trace_rays(const Scene sc, int n_rays, const double *o_in, const double *v_in, const double *limits,
             int32_t *found, int32_t *ids, double *ts, double *hits, double *normals,
             uint32_t *mb_bits, uint32_t mb_stride, uint32_t mb_words, uint32_t mb_shift, int *overflow,
             const void *leafrec, const void *boxrec)
{
    const int lane = threadIdx.x & 31;
    Mailbox mb;
    mb.bits = mb_bits; mb.stride = mb_stride;
    mb.slot = blockIdx.x * blockDim.x + threadIdx.x;
    mb.words = mb_words; mb.group_shift = mb_shift;
    mb.dirty = ~0ull;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpStage<NP> ws;
    ws.init(smem_raw + (threadIdx.x >> 5) * warp_smem_bytes<NP>(sc.any_boxed), leafrec, sc.any_boxed ? boxrec : nullptr, lane);
    ws.init_nested(sc.any_boxed);
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
        trace_kd_warp<NP, true>(sc, ws, mb, want, o, v, (want && limits) ? limits[r] : -1.0, T, ovf, -1);
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
    int (*trace_blocks_per_sm)(int boxed);
    void (*trace)(int mode, int blocks, cudaStream_t st, const Scene &sc, const WaveArgs &a);
    void (*shade)(int phase, int blocks, cudaStream_t st, const Scene &sc, const WaveArgs &a);

    int (*blocks_per_sm)(bool cnt, int boxed);
    void (*generation)(bool cnt, int blocks, cudaStream_t st, const Scene &sc, const GenArgs &a);
    void (*pack_leaf)(cudaStream_t st, const Scene &sc, int n_refs, void *out, void *box_out, int id_base);
    void (*trace_rays)(int blocks, cudaStream_t st, const Scene &sc, int n_rays, const double *o, const double *v,
                       const double *limits, int32_t *found, int32_t *ids, double *ts, double *hits, double *normals,
                       uint32_t *mb_bits, uint32_t mb_stride, uint32_t mb_words, uint32_t mb_shift, int *overflow,
                       const void *leafrec, const void *boxrec);
    /* for the graph nodes of the device-side generation loop (kernels.cu) */
    const void *(*pre_fn)(int mode);
    void (*pre)(int mode, int blocks, cudaStream_t st, const Scene &sc, const WaveArgs &a);
    int (*pre_grid)(int sm_count);
    const void *(*trace_fn)(int mode, int stage);   /* stage = Scene::any_boxed: bit 2 selects the big-leaf instantiation */
    const void *(*shade_fn)(int phase);
    size_t (*trace_smem_bytes)(int boxed);
    int (*shade_grid)(int sm_count, int gen_cap);   /* blocks of a k_shade launch */
    const void *(*light_fn)();
    void (*light)(int blocks, cudaStream_t st, const Scene &sc, const WaveArgs &a);
};
const NpOps *ndt_np_ops(int np);     /* NULL: dimension not built */
