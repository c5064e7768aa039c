/*
 * warp.cuh -- the warp-synchronous form of the nearest-hit query (CUDA only).
 *
 * core.cuh's trace_kd walks one ray at a time: id -> bounding sphere ->
 * object record -> geometry is a chain of dependent loads per object, and ncu
 * showed the kernel latency bound on exactly that chain (profiles/r01_*).  Here
 * the 32 rays of a warp walk the kd-tree on their own but visit LEAVES
 * together:
 *
 *   1. every lane advances its traversal (kd-tree.c:482-568, same near/far rule
 *      and deferred *t_ptr guards as trace_kd) until it stands on a leaf;
 *   2. the warp picks the leaf of its lowest waiting lane; all lanes on that
 *      leaf take part (8x4 pixel blocks: one leaf for the whole warp 9 times
 *      out of 10), the others wait their turn;
 *   3. the leaf's records stream through shared memory in chunks of 32: one
 *      TMA bulk copy (cp.async.bulk + mbarrier) per chunk, double buffered, from
 *      an array packed in leaf order at upload time (LeafRec: bounding sphere,
 *      id, type, geometry offset = everything the loop needs before it touches
 *      the geometry);
 *   4. BROAD phase, no branches: each lane runs the ray-only part of the
 *      bounding-sphere pre-test (bounding.c:52-84) on the 32 staged records
 *      and keeps a 32-bit candidate mask;
 *   5. NARROW phase over the union of the masks, in list order: mailbox,
 *      the min_dist part of the pre-test (bounding.c:43-50), the per-type
 *      intersection, trace()'s EPSILON hysteresis and early break
 *      (object.c:692-747).
 *
 * Why this is the reference's result bit for bit.  trace() is sequential:
 * min_dist, the mailbox and the break depend on everything before.  Only
 * step 4 is hoisted out of that order, and it is a pure function of ray and
 * sphere.  An object failing it is rejected by the reference at the same test
 * whenever it gets there, in this leaf or any later one, so it never needs a
 * mailbox bit; objects passing it are handled strictly in list order with the
 * reference's own state.
 */
#pragma once
#include "wave.cuh"

namespace ndt {

#ifndef NDT_BROAD_UNROLL
#define NDT_BROAD_UNROLL 1    /* instruction-fetch bound: 1 measured faster than 2, 4, 8 */
#endif
#define NDT_STR2(x) #x
#define NDT_STR(x) NDT_STR2(x)
#define NDT_BROAD_LOOP _Pragma(NDT_STR(unroll NDT_BROAD_UNROLL))

constexpr unsigned FULL = 0xffffffffu;
constexpr int CHUNK = 32;

/* one leaf reference, packed at upload time (k_pack_leaf) */
template <int NP> struct LeafRec {
    double c[NP];         /* bounding sphere centre */
    double r2, r;         /* radius^2, radius (<= 0: no pre-test, object.c:618) */
    int32_t id;           /* kd item id = objects[] index */
    uint32_t tfa;         /* type | flags << 4 | n_axes << 8 | (16-byte units of the geometry block, 0 = not staged) << 16 */
    uint32_t geom_off;
    int32_t report_id;
    /* the box cull (see box_hit): */
    uint32_t boxed;       /* 1: lo/hi bound every point the primitive can report as a hit */
    uint32_t par_mask;    /* bit L: DIRECTIONAL light L runs (nearly) parallel to the primitive's flat, its shadow
                             queries must not be culled */
    uint32_t pad[2];
};                        /* 8 NP + 48 bytes, a multiple of 16 */

/* the boxes of the box cull, a second stream parallel to LeafRec[] that is only staged for scenes that
 * have boxed primitives (Scene::any_boxed) */
template <int NP> struct BoxRec {
    float lo[NP], hi[NP];
};                        /* 8 NP bytes, a multiple of 16 */

/* doubles of the largest geometry block staged in shared memory (ndt_flat.h layouts):
 * an orthotope / hcylinder with NP axes, or a facet */
template <int NP> __host__ __device__ constexpr int geom_max_doubles()
{
    return (((NP + NP * NP + 3 * NP + 1) > (6 * NP + 8) ? (NP + NP * NP + 3 * NP + 1) : (6 * NP + 8)) + 1) & ~1;
}
/* doubles of an object's geometry block; 0 = nothing to stage (hcube) */
__host__ __device__ inline int geom_block_doubles(int type, int m, int np)
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

/* per warp: two LeafRec chunks, two geometry blocks, four mbarriers and -- only for scenes with boxed
 * primitives -- two BoxRec chunks (dynamic shared memory: the launch asks for what the scene needs).
 * `mode` is Scene::any_boxed: bit 0 = boxes are staged, bit 1 = a second, identical staging area for the
 * face lists nested in hcubes (warp_nested). */
template <int NP> __host__ __device__ constexpr int warp_stage_bytes(bool boxed)
{
    return 2 * CHUNK * (int)sizeof(LeafRec<NP>) + 2 * geom_max_doubles<NP>() * 8 + 32 +
           (boxed ? 2 * CHUNK * (int)sizeof(BoxRec<NP>) + NP * 16 : 0);     /* + the ray bundle's bounds */
}
template <int NP> __host__ __device__ constexpr int warp_smem_bytes(int mode)
{
    return warp_stage_bytes<NP>((mode & 1) != 0) * ((mode & 2) ? 2 : 1);
}

/* ---- mbarrier + TMA bulk copy (PTX ISA: cp.async.bulk, mbarrier) ------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
/* bounded wait: a wedged copy must surface as an error, never as a hung GPU */
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (int spin = 0; spin < (1 << 22); ++spin)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}

/* per-warp staging area and its pipeline state */
template <int NP> struct WarpStage {
    /* one base pointer each and arithmetic on the buffer index: an array of pointers indexed by a
     * run-time value lives in local memory, and every buf[s] in the leaf loop was a local load
     * (ncu: long_scoreboard on exactly those lines) */
    LeafRec<NP> *buf0;    /* two chunks of CHUNK records */
    BoxRec<NP> *bbuf0;    /* the chunks' boxes (only present / filled when boxes != NULL) */
    float4 *bundle;       /* per axis: (min o, max o, min 1/v, max 1/v) over the warp's walking rays (bundle cull) */
    double *gbuf0;        /* two staged geometry blocks of geom_max_doubles<NP>() */
    uint64_t *bar;        /* bar[0..1]: record chunks, bar[2..3]: geometry blocks */
    __device__ __forceinline__ LeafRec<NP> *buf(int s) const { return buf0 + s * CHUNK; }
    __device__ __forceinline__ BoxRec<NP> *bbuf(int s) const { return bbuf0 + s * CHUNK; }
    __device__ __forceinline__ double *gbuf(int g) const { return gbuf0 + g * geom_max_doubles<NP>(); }
    uint32_t phase;       /* bit s = parity the next wait on bar[s] uses */
    const LeafRec<NP> *stream;
    const BoxRec<NP> *boxes;   /* NULL: the scene has no boxed primitive */
    int lane;
    int fault;            /* a bounded wait ran out */
    unsigned char *nest;  /* the second staging area (hcube face lists), NULL: none */
    uint32_t nphase;      /* its barrier parities */

    /* lay the pointers over a staging area (no barrier initialisation) */
    __device__ __forceinline__ void place(unsigned char *smem)
    {
        buf0 = reinterpret_cast<LeafRec<NP> *>(smem);
        gbuf0 = reinterpret_cast<double *>(buf0 + 2 * CHUNK);
        bar = reinterpret_cast<uint64_t *>(gbuf0 + 2 * geom_max_doubles<NP>());
        bbuf0 = reinterpret_cast<BoxRec<NP> *>(bar + 4);        /* present only when boxes != NULL */
        bundle = reinterpret_cast<float4 *>(bbuf0 + 2 * CHUNK); /* likewise */
    }
    /* mode & 2 (Scene::any_boxed): the launch reserved a second area behind the first */
    __device__ __forceinline__ void init_nested(int mode)
    {
        nest = nullptr;
        nphase = 0;
        if (mode & 2) {
            nest = reinterpret_cast<unsigned char *>(buf0) + warp_stage_bytes<NP>(true);
            if (lane == 0) {
                uint64_t *nb = reinterpret_cast<uint64_t *>(nest + 2 * CHUNK * sizeof(LeafRec<NP>) + 2 * geom_max_doubles<NP>() * 8);
                mbar_init(nb, 1);
                mbar_init(nb + 1, 1);
                mbar_init(nb + 2, 1);
                mbar_init(nb + 3, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncwarp();
        }
    }

    __device__ __forceinline__ void init(unsigned char *smem, const void *leafrec, const void *boxrec, int lane_)
    {
        nest = nullptr;
        nphase = 0;
        buf0 = reinterpret_cast<LeafRec<NP> *>(smem);
        boxes = static_cast<const BoxRec<NP> *>(boxrec);
        gbuf0 = reinterpret_cast<double *>(buf0 + 2 * CHUNK);
        bar = reinterpret_cast<uint64_t *>(gbuf0 + 2 * geom_max_doubles<NP>());
        bbuf0 = reinterpret_cast<BoxRec<NP> *>(bar + 4);        /* present only when boxes != NULL */
        bundle = reinterpret_cast<float4 *>(bbuf0 + 2 * CHUNK); /* likewise */
        phase = 0;
        stream = static_cast<const LeafRec<NP> *>(leafrec);
        lane = lane_;
        fault = 0;
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_init(bar + 1, 1);
            mbar_init(bar + 2, 1);
            mbar_init(bar + 3, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    /* lane 0 starts the copy of `cnt` records at stream[first] into buffer s */
    __device__ __forceinline__ void fetch(int s, int first, int cnt)
    {
        if (lane == 0 && !fault) {
            const uint32_t bytes = (uint32_t)cnt * (uint32_t)sizeof(LeafRec<NP>);
            const uint32_t bbytes = boxes ? (uint32_t)cnt * (uint32_t)sizeof(BoxRec<NP>) : 0u;
            mbar_expect_tx(bar + s, bytes + bbytes);
            bulk_g2s(buf(s), stream + first, bytes, bar + s);
            if (boxes) bulk_g2s(bbuf(s), boxes + first, bbytes, bar + s);
        }
    }
    /* lane 0 starts the copy of a geometry block (n16 16-byte units at geom[off]) into gbuf[g] */
    __device__ __forceinline__ void fetch_geom(int g, const double *geom, uint32_t off, uint32_t n16)
    {
        if (lane == 0 && !fault) {
            mbar_expect_tx(bar + 2 + g, n16 * 16u);
            bulk_g2s(gbuf(g), geom + off, n16 * 16u, bar + 2 + g);
        }
    }
    /* all 32 lanes wait on the same barrier, so `fault` stays warp-uniform; once set,
     * nothing is fetched or waited for any more and the launch reports the error */
    __device__ __forceinline__ void wait(int s)
    {
        if (fault) return;
        if (!mbar_wait(bar + s, (phase >> s) & 1u)) fault = 1;
        fault = __any_sync(FULL, fault) ? 1 : 0;
        phase ^= 1u << s;
    }
};

/* ---- the box cull -------------------------------------------------------------------------
 * Not in the reference: a test that can only REMOVE intersection tests whose answer is "no hit".
 * The bounding spheres of ndt are loose for flat primitives (a 5-face of the 8-cube has the
 * sphere of the whole cube): in BASELINE config 2, 43 % of the objects of a leaf pass the sphere
 * test, 99 % of those then miss (tools/workload_stats.py).  An orthotope can only report a hit
 * point p0 + sum s_a b_a + d with s_a in [-EPS, len_a + EPS] (orthotope.c:122-148) and |d|^2 <=
 * 2 EPS (the quadratic's "qc -= EPSILON", orthotope.c:199, and its closest-approach fallback
 * |dist| <= EPSILON, :262): k_pack_leaf bounds that set by an axis-aligned box with a margin, and
 * a ray that misses the box cannot hit.  trace() has no state that a missing "no hit" would
 * change (object.c:715-733: only a hit touches min_dist or breaks), and a culled object needs no
 * mailbox bit for the same reason as an object failing the sphere test (header of this file).
 *
 * The one way a "no hit" differs from the reference's answer is its NaN / infinity path: for a
 * ray (nearly) parallel to the flat, |P|^2 < EPS, orthotope.c:236-242 divides by a qb that can be
 * exactly zero and then reports a hit at t = +-inf.  For every kind of query but one that answer
 * is inert (NaN / inf distances are never accepted over a finite one, never "< dist_limit", and
 * an accepted inf never reaches the kd result: trace_kd_warp's `lmd < lt`).  The exception is the
 * any-hit query of a DIRECTIONAL light (dist_limit == 0.0 breaks on ANY reported hit,
 * object.c:729).  All queries of one such light share one direction, so k_pack_leaf evaluates
 * |P|^2 for (primitive, light) with the intersection's own arithmetic (axes_PQ) and flags the
 * pairs below EPS: those are never culled.
 * ------------------------------------------------------------------------------------------- */
/* The slab test runs in fp32 (full-rate pipe, native min/max that drop a NaN operand) and BEFORE the
 * fp64 sphere test, which then only sees the ~1 % of boxed objects that pass.  Rounding: every fp32 step
 * moves a box face by at most ~2^-22 * max(|coordinate|, |origin|) in position space; the box carries
 * 0.03 of margin where 0.0142 is needed, and ndt_b200_upload switches the cull off for scenes whose
 * coordinates exceed 2e4 (error 5e-3). */
template <int NP> __device__ __forceinline__ bool box_hit(const float *lo, const float *hi, const float *of, const float *vif)
{
    float tmin = 0.0f, tmax = FLT_MAX;
    NDT_UNROLL
    for (int i = 0; i < NP; i += 4) {
        const float4 l4 = *reinterpret_cast<const float4 *>(lo + i);
        const float4 h4 = *reinterpret_cast<const float4 *>(hi + i);
        const float l[4] = { l4.x, l4.y, l4.z, l4.w }, h[4] = { h4.x, h4.y, h4.z, h4.w };
        NDT_UNROLL
        for (int k = 0; k < 4; ++k) {
            if (i + k < NP) {
                const float t1 = __fmul_rn(__fsub_rn(l[k], of[i + k]), vif[i + k]);
                const float t2 = __fmul_rn(__fsub_rn(h[k], of[i + k]), vif[i + k]);
                tmin = fmaxf(tmin, fminf(t1, t2));      /* fminf / fmaxf drop a NaN operand (0 * inf): no constraint */
                tmax = fminf(tmax, fmaxf(t1, t2));
            }
        }
    }
    /* one more ulp-scale allowance on the comparison itself */
    return tmin <= tmax * 1.000001f + 1e-30f;
}

/* box_hit on a box held in registers (the sparse-warp form of the broad phase): the same operations in the same order */
template <int NP> __device__ __forceinline__ bool box_hit_regs(const float *lo, const float *hi, const float *of, const float *vif)
{
    float tmin = 0.0f, tmax = FLT_MAX;
    NDT_UNROLL
    for (int i = 0; i < NP; ++i) {
        const float t1 = __fmul_rn(__fsub_rn(lo[i], of[i]), vif[i]);
        const float t2 = __fmul_rn(__fsub_rn(hi[i], of[i]), vif[i]);
        tmin = fmaxf(tmin, fminf(t1, t2));
        tmax = fminf(tmax, fmaxf(t1, t2));
    }
    return tmin <= tmax * 1.000001f + 1e-30f;
}

/* BUNDLE CULL: box_hit for all rays of the warp at once, one RECORD per lane instead of one ray per lane.
 * The 32 rays of a warp are neighbours (an 8x4 pixel block, or queue neighbours spawned by one), 99 % of a
 * leaf's boxes are missed by each of them, and mostly by all of them: per axis the bundle is described by the
 * intervals [o_lo, o_hi] and [vi_lo, vi_hi] of its origins and reciprocal directions (trace_kd_warp), and a
 * record is dropped for the whole warp when the interval version of the slab test already fails.
 * Exactness: the bounds below are built from the SAME fp32 operations as box_hit applied to interval end
 * points; IEEE rounding is monotone, so tmin_lb <= every ray's computed tmin and tmax_ub >= every ray's
 * computed tmax: a record dropped here would have failed box_hit for every ray of the bundle (the per-ray
 * result is unchanged, bit for bit).  Axes on which the rays do not all run the same way (or some are
 * parallel and some not) give no constraint; if all are parallel (1/v = +-inf) and every origin lies on one
 * side of the slab, tmin / tmax take the +-inf box_hit computes for each of them.  tests/test_bundle_math.py
 * restates both functions in numpy fp32 and checks the implication on random bundles. */
template <int NP> __device__ __forceinline__ bool bundle_hit(const float *lo, const float *hi, const float4 *bd)
{
    float tmin = 0.0f, tmax = FLT_MAX;
    /* rolled: once per chunk and warp, not worth instruction-cache space (unrolled, the hit rate of the SM
     * instruction cache fell from 98 % to 80 %) */
    NDT_NO_UNROLL
    for (int i = 0; i < NP; ++i) {
        const float4 b = bd[i];                         /* o_lo, o_hi, vi_lo, vi_hi: warp-uniform */
        const float l = lo[i], h = hi[i];
        if (b.z > 0.0f && b.w < FLT_MAX) {              /* all rays go up this axis: near = (l-o) vi, far = (h-o) vi */
            const float a = __fsub_rn(l, b.y), d = __fsub_rn(h, b.x);
            tmin = fmaxf(tmin, fminf(__fmul_rn(a, b.z), __fmul_rn(a, b.w)));
            tmax = fminf(tmax, fmaxf(__fmul_rn(d, b.z), __fmul_rn(d, b.w)));
        } else if (b.w < 0.0f && b.z > -FLT_MAX) {      /* all go down: near = (h-o) vi, far = (l-o) vi */
            const float a = __fsub_rn(l, b.y), d = __fsub_rn(h, b.x);
            tmin = fmaxf(tmin, fminf(__fmul_rn(d, b.z), __fmul_rn(d, b.w)));
            tmax = fminf(tmax, fmaxf(__fmul_rn(a, b.z), __fmul_rn(a, b.w)));
        } else if (b.z == b.w && (b.z > FLT_MAX || b.z < -FLT_MAX)) {   /* all parallel to the slab */
            /* box_hit's own values: (l-o) inf and (h-o) inf are both -inf behind the slab, both +inf before it */
            if (b.x > h) tmax = -INFINITY;
            if (b.y < l) tmin = INFINITY;
        }
    }
    return tmin <= tmax * 1.000001f + 1e-30f;
}
/* order-preserving float <-> int, for redux.sync (integer only) */
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

/* the ray-only half of bounding.c:34-85: desc = (v.oc)^2 - |oc|^2 + r^2 */
template <int NP> __device__ __forceinline__ bool bsphere_ray_part(const double *c, double r2, const double *o, const double *v)
{
    double oc[NP];
    vsub<NP>(o, c, oc);
    double oc2 = vdot<NP>(oc, oc);
    double voc = vdot<NP>(v, oc);
    double voc2 = voc * voc;
    double desc = voc2 - oc2 + r2;
    return !(desc < 0.0 || (voc > 0.0 && voc2 > desc));
}
/* the min_dist half (bounding.c:43-50): true = rejected */
template <int NP> __device__ __forceinline__ bool bsphere_far(const double *c, double r, const double *o, double min_dist)
{
    if (!(min_dist > 0)) return false;
    double oc[NP];
    vsub<NP>(o, c, oc);
    double oc2 = vdot<NP>(oc, oc);
    double mr = min_dist + r;
    return oc2 > mr * mr;
}

template <int NP> __device__ __forceinline__ void lds_vec(double *d, const double *s)
{
    NDT_UNROLL
    for (int i = 0; i < NP; i += 2) {
        double2 t = *reinterpret_cast<const double2 *>(s + i);
        d[i] = t.x; d[i + 1] = t.y;
    }
}

/* trace() for the lists that are not worth staging -- the infinite objects
 * (kd-tree.c:592-594) and the faces nested in an hcube (hcube.c:236-250): one
 * out-of-line copy of core.cuh's scalar loop, so that the hot leaf code above
 * stays small in the instruction cache.  The ray is passed through local
 * memory (copies, so that the caller's o/v stay in registers). */
template <int NP>
__device__ __noinline__ double trace_list_slow_impl(const Scene &sc, const int32_t *ids, int cnt, int base,
                                                    const double *o_in, const double *v_in, double dist_limit,
                                                    int *out_id, int *out_win)
{
    double o[NP], v[NP];
    vcopy<NP>(o, o_in);
    vcopy<NP>(v, v_in);
    Tally<false> none;
    int oid, owin;
    const double md = trace_list<NP, false>(sc, ids, cnt, (Mailbox *)nullptr, o, v, dist_limit, oid, owin, none, base);
    *out_id = oid;
    *out_win = owin;
    return md;
}
template <int NP>
__device__ __forceinline__ double trace_list_slow(const Scene &sc, const int32_t *ids, int cnt, int base,
                                                  const double *o, const double *v, double dist_limit,
                                                  int *out_id, int *out_win)
{
    double ot[NP], vt[NP];
    vcopy<NP>(ot, o);
    vcopy<NP>(vt, v);
    return trace_list_slow_impl<NP>(sc, ids, cnt, base, ot, vt, dist_limit, out_id, out_win);
}

template <int NP, bool NESTED, bool BIG>
__device__ __forceinline__ double warp_leaf(const Scene &sc, WarpStage<NP> &ws, int first, int count, bool mine,
                                            Mailbox &mb, const double *o, const double *v, const float *obox,
                                            const float *vbox, uint32_t keep_mask, double dist_limit,
                                            int &out_id, int &out_win);

/* The face list nested in an hcube (hcube.c:236-250: trace() over hcube->obj with no mailbox, no limit, its own
 * min_dist) for the lanes with `mine`, through the SECOND staging area of the warp: the same record / box streams,
 * bundle cull, fp32 slab test, sphere pre-test and staged narrow phase as a kd leaf gets.  The scalar loop
 * (core.cuh: trace_list) tested the bounding sphere of every face for every ray that passed the cube's own
 * sphere -- 472 faces for a 6-cube, 98 % of the FP64 work of BASELINE config 3.  Out of line: one copy, and
 * the caller's registers (its ray, its leaf state) are not the nested loop's problem. */
template <int NP>
__device__ __noinline__ double warp_nested(const Scene &sc, unsigned char *area, float4 *bundle, int lane, int first,
                                           int count, bool mine, const double *o_in, const double *v_in,
                                           const float *obox, const float *vbox, uint32_t keep_mask,
                                           uint32_t *phase_io, int *fault_io, int *out_win)
{
    WarpStage<NP> wn;
    wn.place(area);
    wn.bundle = bundle;                 /* the rays of this call are a subset of the bundle's: the cull stays valid */
    wn.stream = static_cast<const LeafRec<NP> *>(sc.nrec);
    wn.boxes = static_cast<const BoxRec<NP> *>(sc.nbox);
    wn.lane = lane;
    wn.phase = *phase_io;
    wn.fault = *fault_io;
    wn.nest = nullptr;
    wn.nphase = 0;
    double o[NP], v[NP];
    vcopy<NP>(o, o_in);
    vcopy<NP>(v, v_in);
    Mailbox none;
    none.bits = nullptr; none.stride = 0; none.slot = 0; none.words = 0; none.group_shift = 0; none.dirty = 0;
    int oid, owin;
    const double md = warp_leaf<NP, true, true>(sc, wn, first, count, mine, none, o, v, obox, vbox, keep_mask, -1.0, oid, owin);
    *phase_io = wn.phase;
    *fault_io = wn.fault;
    *out_win = owin;
    return md;
}

/* The face list of an hcube for ONE ray (lane L's), the 32 lanes side by side on 32 faces: what warp_nested does
 * for a bundle, for the incoherent warps of the bounce generations, where one or two lanes ask for a cube and
 * walking its list in step leaves 30 lanes idle.
 *   1. every lane runs the slab test and the ray half of the sphere pre-test for its face (boxes and records
 *      straight from global memory, 15 rounds for the 472 faces of a 6-cube); the faces that pass are
 *      appended, in list order, to a list in shared memory;
 *   2. whenever 32 are listed (and at the end) each lane intersects ONE of them -- a pure function of
 *      (face, ray): the nested trace() has no mailbox and no limit, and every face is an orthotope, whose
 *      intersect() overwrites hit and normal before it reads them;
 *   3. the results are folded in list order with trace()'s own rule (object.c:715-727) and, in front of it, the
 *      min_dist half of the pre-test (bounding.c:43-50), which only decides whether a face is asked at all.
 * Returns the nested trace()'s min_dist (< 0: nothing accepted) and the winning face, identical in every lane. */
template <int NP>
__device__ __noinline__ double hcube_one_ray(const Scene &sc, int32_t *list, int lane, int L, int first, int count,
                                             const double *o_in, const double *v_in, const float *obox,
                                             const float *vbox, uint32_t keep_mask, int *out_win)
{
    const LeafRec<NP> *recs = static_cast<const LeafRec<NP> *>(sc.nrec) + first;
    const BoxRec<NP> *boxes = static_cast<const BoxRec<NP> *>(sc.nbox) + first;
    double o[NP], v[NP];
    float of[NP], vif[NP];
    NDT_UNROLL
    for (int i = 0; i < NP; ++i) {
        o[i] = __shfl_sync(FULL, o_in[i], L);
        v[i] = __shfl_sync(FULL, v_in[i], L);
        of[i] = __shfl_sync(FULL, obox[i], L);
        vif[i] = __shfl_sync(FULL, vbox[i], L);
    }
    const uint32_t km = __shfl_sync(FULL, keep_mask, L);
    double in_min = -1;
    int win = -1, n_list = 0;
    const int rounds = (count + 31) >> 5;
    for (int rd = 0; rd <= rounds; ++rd) {
        if (rd < rounds) {
            const int f = rd * 32 + lane;
            bool pass = false;
            if (f < count) {
                const double *rp = reinterpret_cast<const double *>(recs + f);
                const uint2 bx = *reinterpret_cast<const uint2 *>(rp + NP + 4);        /* boxed, par_mask */
                pass = true;
                if (bx.x && !(bx.y & km)) {
                    float lo[NP], hi[NP];
                    NDT_UNROLL
                    for (int i = 0; i < NP; i += 2) {
                        const float2 l2 = *reinterpret_cast<const float2 *>(boxes[f].lo + i);
                        const float2 h2 = *reinterpret_cast<const float2 *>(boxes[f].hi + i);
                        lo[i] = l2.x; lo[i + 1] = l2.y; hi[i] = h2.x; hi[i + 1] = h2.y;
                    }
                    pass = box_hit_regs<NP>(lo, hi, of, vif);
                }
                if (pass) {
                    double c[NP];
                    NDT_UNROLL
                    for (int i = 0; i < NP; ++i) c[i] = rp[i];
                    const double r2 = rp[NP], r = rp[NP + 1];
                    pass = !(r > 0) || bsphere_ray_part<NP>(c, r2, o, v);
                }
            }
            const unsigned m = __ballot_sync(FULL, pass);
            if (pass) list[n_list + __popc(m & ((1u << lane) - 1u))] = f;
            n_list += __popc(m);
            __syncwarp();
        }
        /* a full batch, or what is left after the last round */
        while (n_list >= 32 || (rd == rounds && n_list > 0)) {
            const int nb = n_list < 32 ? n_list : 32;
            bool ret = false;
            double dist = -1, oc2 = 0, rad = 0;
            int f = -1;
            if (lane < nb) {
                f = list[lane];
                const double *rp = reinterpret_cast<const double *>(recs + f);
                const int4 meta = *reinterpret_cast<const int4 *>(rp + NP + 2);
                double c[NP], oc[NP];
                NDT_UNROLL
                for (int i = 0; i < NP; ++i) c[i] = rp[i];
                rad = rp[NP + 1];
                vsub<NP>(o, c, oc);
                oc2 = vdot<NP>(oc, oc);                 /* bsphere_far's own arithmetic */
                ndt_flat_object fo;
                fo.type = NDT_T_ORTHOTOPE;
                fo.flags = (int32_t)(((uint32_t)meta.y >> 4) & 0xfu);
                fo.n_axes = (int32_t)(((uint32_t)meta.y >> 8) & 0xffu);
                fo.geom_off = (uint32_t)meta.z;
                fo.report_id = meta.w;
                Tally<false> none;
                double res[NP], nrm[NP];
                vzero<NP>(res);
                vzero<NP>(nrm);
                ret = intersect_prim<NP, false, LdGlobal>(sc, fo, sc.geom + fo.geom_off, o, v, res, nrm, none);
                if (ret) dist = vdist<NP>(o, res);
            }
            __syncwarp();
            /* the faces that were not listed failed the pre-test; the listed ones that report no hit change nothing */
            for (unsigned hm = __ballot_sync(FULL, ret); hm; hm &= hm - 1) {
                const int src = __ffs(hm) - 1;
                const double d = __shfl_sync(FULL, dist, src), q = __shfl_sync(FULL, oc2, src), r = __shfl_sync(FULL, rad, src);
                const int fid = __shfl_sync(FULL, f, src);
                if (r > 0 && in_min > 0) {              /* bounding.c:43-50 with the min_dist of this point of the list */
                    const double mr = in_min + r;
                    if (q > mr * mr) continue;
                }
                if (d > EPS && (d + EPS < in_min || in_min < 0)) {
                    in_min = d;
                    win = fid;
                }
            }
            /* move the rest of the list to the front */
            int keep = -1;
            if (lane + nb < n_list) keep = list[lane + nb];
            __syncwarp();
            if (lane + nb < n_list) list[lane] = keep;
            n_list -= nb;
            __syncwarp();
        }
    }
    *out_win = win >= 0 ? first + win + sc.n_items : -1;
    return in_min;
}

/* One hcube of a kd leaf whose face list goes through the second staging area (warp_nested): the whole warp
 * stages it for the lanes that ask (`want`: live, and a candidate of the broad phase). */
template <int NP>
__device__ __forceinline__ void warp_hcube(const Scene &sc, WarpStage<NP> &ws, const double *rp, const int4 meta, bool want,
                                           Mailbox &mb, const double *o, const double *v, const float *obox, const float *vbox,
                                           uint32_t keep_mask, double dist_limit, double &min_dist, int &out_id, int &out_win,
                                           bool &live)
{
    const int id = meta.x;
    bool test = want;
    if (test) {                                /* object.c:706-713, bounding.c:43-50 */
        uint32_t *mword = mb.word((uint32_t)id >> 5);
        const uint32_t mcur = *mword, mbit = 1u << (id & 31);
        if (mcur & mbit) test = false;
        else {
            *mword = mcur | mbit;
            mb.dirty |= 1ull << (((uint32_t)id >> 5) >> mb.group_shift);
        }
        const double2 rr = *reinterpret_cast<const double2 *>(rp + NP);
        if (test && rr.y > 0) {
            double c[NP];
            lds_vec<NP>(c, rp);
            if (bsphere_far<NP>(c, rr.y, o, min_dist)) test = false;
        }
    }
    const unsigned asking = __ballot_sync(FULL, test);
    if (asking) {
        /* nested trace() (hcube.c:236-250): no mailbox, no limit, own min_dist */
        const ndt_flat_object *top = sc.obj + id;
        const int cb = NDT_LDG(&top->child_begin), cc = NDT_LDG(&top->child_count);
        double ot[NP], vt[NP];
        vcopy<NP>(ot, o);
        vcopy<NP>(vt, v);
        double in_min = -1;
        int cwin = -1;
#ifndef NDT_HCUBE_ONE_RAY_MAX
#define NDT_HCUBE_ONE_RAY_MAX 6       /* up to this many lanes asking: one ray at a time across the lanes */
#endif
        if (__popc(asking) <= NDT_HCUBE_ONE_RAY_MAX) {
            /* the second staging area is idle here: its first 256 bytes hold the candidate list (63 entries at most) */
            for (unsigned am = asking; am; am &= am - 1) {
                const int L = __ffs(am) - 1;
                int w1 = -1;
                const double m1 = hcube_one_ray<NP>(sc, reinterpret_cast<int32_t *>(ws.nest), ws.lane, L, cb - sc.n_items, cc, ot, vt,
                                                    obox, vbox, keep_mask, &w1);
                if (ws.lane == L) { in_min = m1; cwin = w1; }
            }
        } else {
            uint32_t ph = ws.nphase;
            int fl = ws.fault;
            in_min = warp_nested<NP>(sc, ws.nest, ws.bundle, ws.lane, cb - sc.n_items, cc, test, ot, vt,
                                     obox, vbox, keep_mask, &ph, &fl, &cwin);
            ws.nphase = ph;
            ws.fault = fl;
        }
        if (test && !(in_min < 0)) {
            const double dist = in_min;
            if (dist > EPS && (dist + EPS < min_dist || min_dist < 0)) {
                min_dist = dist;
                out_id = meta.w;
                out_win = cwin;
            }
            if (dist_limit == 0.0 || dist < dist_limit) live = false;
        }
    }
}

/* trace() (object.c:692-747) of one leaf for the lanes with `mine`, all 32
 * lanes of the warp taking part in the staging.  Per lane on return: min_dist
 * (<0: nothing accepted), out_id, out_win.  NESTED: the face list of an hcube (warp_nested): every record is
 * an orthotope, no mailbox, dist_limit -1. */
template <int NP, bool NESTED, bool BIG>
__device__ __forceinline__ double warp_leaf(const Scene &sc, WarpStage<NP> &ws, int first, int count, bool mine,
                                            Mailbox &mb, const double *o, const double *v, const float *obox,
                                            const float *vbox, uint32_t keep_mask, double dist_limit,
                                            int &out_id, int &out_win)
{
    double min_dist = -1;
    out_id = -1;
    out_win = -1;
    bool live = mine && !ws.fault;  /* false once this lane's trace() has hit its break */
    if (ws.fault) return min_dist;
    /* trace() hands one calloc'ed hit/normal pair to every object of the list (object.c:700-703):
     * what vectNd_copy leaves in the pad lane carries over from one object to the next */
    double res[NP], nrm[NP];
    vzero<NP>(res);
    vzero<NP>(nrm);
    const int nch = (count + CHUNK - 1) / CHUNK;
    int s = 0;
    /* both buffers are free here: every warp_leaf drains what it started */
    ws.fetch(0, first, count < CHUNK ? count : CHUNK);
    for (int ch = 0; ch < nch; ++ch, s ^= 1) {
        const int cnt = (count - ch * CHUNK) < CHUNK ? (count - ch * CHUNK) : CHUNK;
        if (ch + 1 < nch) {
            const int rest = count - (ch + 1) * CHUNK;
            ws.fetch(s ^ 1, first + (ch + 1) * CHUNK, rest < CHUNK ? rest : CHUNK);
        }
        ws.wait(s);
        const LeafRec<NP> *rec = ws.buf(s);
        const BoxRec<NP> *brec = ws.bbuf(s);

        /* broad phase */
        unsigned cand = 0;
        unsigned surv = cnt >= 32 ? ~0u : (1u << cnt) - 1u;
#ifndef NDT_NO_BUNDLE_CULL
        if (ws.boxes) {
            /* bundle cull, all 32 lanes: lane k tests record k against the warp's ray bundle */
            const uint32_t kmw = __reduce_or_sync(FULL, live ? keep_mask : 0u);
            bool keep = false;
            if (ws.lane < cnt) {
                const double *rp = reinterpret_cast<const double *>(rec + ws.lane);
                const uint2 bx = *reinterpret_cast<const uint2 *>(rp + NP + 4);        /* boxed, par_mask */
                keep = !(bx.x && !(bx.y & kmw)) || bundle_hit<NP>(brec[ws.lane].lo, brec[ws.lane].hi, ws.bundle);
            }
            surv = __ballot_sync(FULL, keep);
        }
#endif
        /* A SPARSE warp (few live rays: the late bounce generations, warps drawn short by k_trace) turns the
         * slab test round as well: one ray at a time, lane k testing record k for it -- R rounds instead of one
         * round per surviving record.  Same function of the same (ray, record) pair, so the same candidates. */
        unsigned own = surv;            /* the records this lane's ray still has to look at */
        const int n_live = BIG ? __popc(__ballot_sync(FULL, live)) : 32;
#ifndef NDT_NO_SPARSE_BROAD
        const bool sparse = BIG && ws.boxes && n_live * 3 < __popc(surv) * 2;
#else
        const bool sparse = false;
#endif
        if (sparse) {
            float lo[NP], hi[NP];
            uint2 bx = make_uint2(0u, 0u);
            const bool have = ws.lane < cnt && ((surv >> ws.lane) & 1u);
            if (have) {
                bx = *reinterpret_cast<const uint2 *>(reinterpret_cast<const double *>(rec + ws.lane) + NP + 4);
                NDT_UNROLL
                for (int i = 0; i < NP; i += 4) {
                    const float4 l4 = *reinterpret_cast<const float4 *>(brec[ws.lane].lo + i);
                    const float4 h4 = *reinterpret_cast<const float4 *>(brec[ws.lane].hi + i);
                    lo[i] = l4.x; hi[i] = h4.x;
                    if (i + 1 < NP) { lo[i + 1] = l4.y; hi[i + 1] = h4.y; }
                    if (i + 2 < NP) { lo[i + 2] = l4.z; hi[i + 2] = h4.z; }
                    if (i + 3 < NP) { lo[i + 3] = l4.w; hi[i + 3] = h4.w; }
                }
            } else {
                NDT_UNROLL
                for (int i = 0; i < NP; ++i) { lo[i] = 0.0f; hi[i] = 0.0f; }
            }
            float of[NP], vif[NP];
            NDT_UNROLL
            for (int i = 0; i < NP; ++i) { of[i] = live ? obox[i] : 0.0f; vif[i] = live ? vbox[i] : 0.0f; }
            NDT_NO_UNROLL
            for (unsigned am = __ballot_sync(FULL, live); am; am &= am - 1) {
                const int L = __ffs(am) - 1;
                float ol[NP], vl[NP];
                NDT_UNROLL
                for (int i = 0; i < NP; ++i) { ol[i] = __shfl_sync(FULL, of[i], L); vl[i] = __shfl_sync(FULL, vif[i], L); }
                const uint32_t kml = __shfl_sync(FULL, keep_mask, L);
                const bool pass = have && (!(bx.x && !(bx.y & kml)) || box_hit_regs<NP>(lo, hi, ol, vl));
                const unsigned m = __ballot_sync(FULL, pass);
                if (ws.lane == L) own = m;
            }
        }
        if (live) {
            if (ws.boxes) {
                /* boxed scene: the fp32 slab test first, the sphere test for what is left */
                float of[NP], vif[NP];
                NDT_UNROLL
                for (int i = 0; i < NP; ++i) { of[i] = obox[i]; vif[i] = vbox[i]; }
                NDT_BROAD_LOOP
                for (unsigned sm = own; sm; sm &= sm - 1) {
                    const int k = __ffs(sm) - 1;
                    const double *rp = reinterpret_cast<const double *>(rec + k);
                    const uint2 bx = *reinterpret_cast<const uint2 *>(rp + NP + 4);    /* boxed, par_mask */
                    bool pass = true;
                    /* (a sparse warp's `own` already went through the slab test) */
                    if (!sparse && bx.x && !(bx.y & keep_mask)) pass = box_hit<NP>(brec[k].lo, brec[k].hi, of, vif);
                    if (pass) {
                        double c[NP];
                        lds_vec<NP>(c, rp);
                        const double2 rr = *reinterpret_cast<const double2 *>(rp + NP);   /* r2, r */
                        pass = !(rr.y > 0) || bsphere_ray_part<NP>(c, rr.x, o, v);
                    }
                    cand |= (pass ? 1u : 0u) << k;
                }
            } else {
                NDT_BROAD_LOOP
                for (int k = 0; k < cnt; ++k) {
                    const double *rp = reinterpret_cast<const double *>(rec + k);
                    double c[NP];
                    lds_vec<NP>(c, rp);
                    const double2 rr = *reinterpret_cast<const double2 *>(rp + NP);   /* r2, r */
                    const bool pass = !(rr.y > 0) || bsphere_ray_part<NP>(c, rr.x, o, v);
                    cand |= (pass ? 1u : 0u) << k;
                }
            }
        }

        /* narrow phase, list order, the warp in step on the union of the masks; the geometry
         * block of the next candidate is copied into shared memory while this one is tested */
        unsigned un = __reduce_or_sync(FULL, cand);
        int gs = 0;
        if (un) {
            const int4 m0 = *reinterpret_cast<const int4 *>(reinterpret_cast<const double *>(rec + (__ffs(un) - 1)) + NP + 2);
            if ((uint32_t)m0.y >> 16) ws.fetch_geom(0, sc.geom, (uint32_t)m0.z, (uint32_t)m0.y >> 16);
        }
        while (un) {
            const int k = __ffs(un) - 1;
            un &= un - 1;
            const unsigned alive = __ballot_sync(FULL, live);
            if (!alive) un = 0;             /* every lane has left its trace(): stop after draining this copy */
            if (un) {
                const int4 m1 = *reinterpret_cast<const int4 *>(reinterpret_cast<const double *>(rec + (__ffs(un) - 1)) + NP + 2);
                if ((uint32_t)m1.y >> 16) ws.fetch_geom(gs ^ 1, sc.geom, (uint32_t)m1.z, (uint32_t)m1.y >> 16);
            }
            const double *rp = reinterpret_cast<const double *>(rec + k);
            const int4 meta = *reinterpret_cast<const int4 *>(rp + NP + 2);
            const bool staged = ((uint32_t)meta.y >> 16) != 0;
            if (staged) ws.wait(2 + gs);
            const int id = meta.x;
            if (BIG && !NESTED && ((uint32_t)meta.y & 0xfu) == NDT_T_HCUBE && ws.nest) {
                /* (warp-uniform branch: the record is the same for all lanes) */
                warp_hcube<NP>(sc, ws, rp, meta, live && ((cand >> k) & 1u), mb, o, v, obox, vbox, keep_mask, dist_limit,
                               min_dist, out_id, out_win, live);
            } else if (live && ((cand >> k) & 1u)) {
                bool skip = false;
                if (!NESTED) {                         /* object.c:706-713 */
                    uint32_t *mword = mb.word((uint32_t)id >> 5);
                    const uint32_t mcur = *mword, mbit = 1u << (id & 31);
                    if (mcur & mbit) skip = true;
                    else {
                        *mword = mcur | mbit;
                        mb.dirty |= 1ull << (((uint32_t)id >> 5) >> mb.group_shift);
                    }
                }
                if (!skip) {
                    const double2 rr = *reinterpret_cast<const double2 *>(rp + NP);
                    if (rr.y > 0) {
                        double c[NP];
                        lds_vec<NP>(c, rp);
                        skip = bsphere_far<NP>(c, rr.y, o, min_dist);
                    }
                }
                if (!skip) {
                    ndt_flat_object fo;
                    fo.type = NESTED ? (int32_t)NDT_T_ORTHOTOPE : (int32_t)((uint32_t)meta.y & 0xfu);
                    fo.flags = (int32_t)(((uint32_t)meta.y >> 4) & 0xfu);
                    fo.n_axes = (int32_t)(((uint32_t)meta.y >> 8) & 0xffu);
                    fo.geom_off = (uint32_t)meta.z;
                    fo.report_id = meta.w;
                    Tally<false> none;
                    bool ret;
                    double dist = -1;
                    int win = id;
                    if (staged) {
                        ret = intersect_prim<NP, false, LdShared>(sc, fo, ws.gbuf(gs), o, v, res, nrm, none);
                        if (ret) dist = vdist<NP>(o, res);
                    } else if (NESTED || fo.type != NDT_T_HCUBE) {
                        ret = false;        /* unreachable: ndt_b200_upload refuses blocks that cannot be staged */
                    } else {
                        /* nested trace() (hcube.c:236-250): no mailbox, no limit, own min_dist */
                        const ndt_flat_object *top = sc.obj + id;
                        int cid, cwin;
                        const double in_min = trace_list_slow<NP>(sc, (const int32_t *)nullptr, NDT_LDG(&top->child_count),
                                                                  NDT_LDG(&top->child_begin), o, v, -1.0, &cid, &cwin);
                        ret = !(in_min < 0);
                        if (ret) { dist = in_min; win = cwin; }
                    }
                    if (ret) {
                        if (dist > EPS && (dist + EPS < min_dist || min_dist < 0)) {
                            min_dist = dist;
                            out_id = fo.report_id;
                            out_win = win;
                        }
                        if (!NESTED && (dist_limit == 0.0 || dist < dist_limit)) live = false;
                    }
                }
            }
            __syncwarp();       /* gbuf[gs] is free again */
            gs ^= 1;
        }
        __syncwarp();           /* everyone is done with buf[s] before it is refilled */
        if (ch + 1 < nch && !__ballot_sync(FULL, live)) {
            ws.wait(s ^ 1);     /* drain the copy already in flight, then leave */
            break;
        }
    }
    return min_dist;
}

/* What every query does before it walks the tree: the infinite objects (kd-tree.c:592-594) and the root box
 * (aabb_intersect, kd-tree.c:84-127).  Nine rays in ten of BASELINE config 2 and most of config 1's end here, so
 * the wavefront runs this half in a kernel of its own at three times k_trace's occupancy (k_pre, gen.cuh) and
 * hands k_trace the walkers only; md / id / win / ret / tl / tu are the state the walk picks up. */
struct PreWalk {
    double md;            /* trace()'s min_dist over the infinite objects (< 0: none) */
    double tl, tu;        /* the ray's interval inside the root box, widened by EPS */
    int id, win, ret;
    bool walking;
};
/* SLIM: the form k_pre runs -- the root box straight from the ray's registers (unrolled) and, when every infinite
 * object is an hplane, a cylinder or an hcylinder (Scene::inf_hplanes), trace() specialised for those types and inlined;
 * otherwise and in
 * k_trace_rays / k_generation the rolled loops and the out-of-line list of round 1. */
template <int NP, bool SLIM>
__device__ __forceinline__ void pre_walk(const Scene &sc, const double *o, const double *v, double dist_limit, bool only_found,
                                         PreWalk &p)
{
    p.walking = false;
    p.tl = p.tu = 0;
    if (SLIM && sc.inf_hplanes == 1) {
        Tally<false> none;
        p.md = trace_list<NP, false, 1u << NDT_T_HPLANE>(sc, sc.inf, sc.n_inf, (Mailbox *)nullptr, o, v, dist_limit, p.id, p.win, none);
    } else if (SLIM && sc.inf_hplanes == 2) {
        constexpr unsigned INF_TYPES = (1u << NDT_T_HPLANE) | (1u << NDT_T_CYLINDER) | (1u << NDT_T_HCYLINDER);
        Tally<false> none;
        p.md = trace_list<NP, false, INF_TYPES>(sc, sc.inf, sc.n_inf, (Mailbox *)nullptr, o, v, dist_limit, p.id, p.win, none);
    } else {
        double ot[NP], vt[NP];
        vcopy<NP>(ot, o);
        vcopy<NP>(vt, v);
        p.md = trace_list_slow_impl<NP>(sc, sc.inf, sc.n_inf, 0, ot, vt, dist_limit, &p.id, &p.win);
    }
    p.ret = !(p.md < 0);
    if ((only_found && p.ret) || sc.n_nodes <= 0) return;
    const double *lo = sc.aabb, *hi = sc.aabb + NP;
    double l = -DBL_MAX, u = DBL_MAX;
    bool behind = false;
    if (SLIM) {
        NDT_UNROLL
        for (int i = 0; i < NP; ++i) {
            if (i < sc.n && !behind) {
                const double vi = v[i], oi = o[i];
                if (!(fabs(vi) < EPS2)) {
                    double a = (NDT_LDG(lo + i) - oi) / vi;
                    double b = (NDT_LDG(hi + i) - oi) / vi;
                    if (a > b) { double x = a; a = b; b = x; }
                    if (a > l) l = a;
                    if (b < u) u = b;
                    if (u < -EPS) behind = true;
                }
            }
        }
    } else {
        double o_dyn[NP], v_dyn[NP];
        NDT_UNROLL
        for (int i = 0; i < NP; ++i) { o_dyn[i] = o[i]; v_dyn[i] = v[i]; }
        NDT_NO_UNROLL
        for (int i = 0; i < sc.n && !behind; ++i) {
            const double vi = v_dyn[i], oi = o_dyn[i];
            if (!(fabs(vi) < EPS2)) {
                double a = (NDT_LDG(lo + i) - oi) / vi;
                double b = (NDT_LDG(hi + i) - oi) / vi;
                if (a > b) { double x = a; a = b; b = x; }
                if (a > l) l = a;
                if (b < u) u = b;
                if (u < -EPS) behind = true;
            }
        }
    }
    if (behind) return;
    l -= EPS;
    u += EPS;
    p.tl = l; p.tu = u;
    p.walking = (u >= -EPS) && (l <= u);
}

/* trace_kd (object.c:683) for the 32 rays of a warp, from where pre_walk left off: lanes with `walking` walk
 * the tree, the others take part in the staging only.  Same contract as core.cuh's trace_kd. */
template <int NP, bool BIG>
__device__ __forceinline__ void trace_kd_resume(const Scene &sc, WarpStage<NP> &ws, Mailbox &mb, bool walking,
                                                const double *o, const double *v, double dist_limit, const PreWalk &pw,
                                                Hit &out, int &overflow, int dir_light)
{
    /* dir_light >= 0: the any-hit query of DIRECTIONAL light dir_light (ndt.c:241-249): the caller consumes
     * nothing but the return value */
    const bool only_found = dir_light >= 0;
    /* primitives this query must not box-cull (see box_hit): bit 31 of par_mask is always set */
    const uint32_t keep_mask = dir_light < 0 ? 0u : (dir_light < 31 ? 1u << dir_light : 0x80000000u);

    /* per-axis values the walk indexes by the split dimension live in local memory; the loops
     * that fill them stay rolled (one copy of the fp64 division sequence instead of NP or 2 NP:
     * this per-ray prologue runs once per query and only costs instruction fetches) */
    double o_dyn[NP], v_dyn[NP], vinv[NP];
    float obox[NP], vbox[NP];       /* the ray in fp32 for the box cull (local memory; registers only inside the broad loop) */
    double t = DBL_MAX, md = pw.md;
    int ret = pw.ret;
    out.id = pw.id;
    out.win = pw.win;
    double tl = pw.tl, tu = pw.tu;
    if (md > EPS) t = md;
    if (walking) {
        NDT_UNROLL
        for (int i = 0; i < NP; ++i) { o_dyn[i] = o[i]; v_dyn[i] = v[i]; }
        mb.clear();
        /* only the rays that enter the root box (12 % in config 2) need the per-axis
         * reciprocals of the walk and of the box cull */
        NDT_NO_UNROLL
        for (int i = 0; i < NP; ++i) {
            double vi = v_dyn[i], r;
            if (vi < EPS2 && vi >= 0.0) r = INV_EPS2;
            else if (vi > -EPS2 && vi <= 0.0) r = -INV_EPS2;
            else r = 1.0 / vi;
            vinv[i] = r;
            if (sc.any_boxed) {     /* the slab test wants the true reciprocal (+-inf for a zero component) */
                obox[i] = (float)o_dyn[i];
                vbox[i] = (float)(1.0 / vi);
            }
        }
    }

#ifndef NDT_NO_BUNDLE_CULL
    /* the bundle of the rays that walk the tree (see bundle_hit): per axis the extremes of origin and
     * reciprocal direction, 4 NP integer reductions per 32 rays, kept in shared memory */
    if (sc.any_boxed && __ballot_sync(FULL, walking)) {
        NDT_NO_UNROLL
        for (int i = 0; i < NP; ++i) {
            const int fo = f2ord(walking ? obox[i] : 0.0f), fv = f2ord(walking ? vbox[i] : 0.0f);
            const int olo = __reduce_min_sync(FULL, walking ? fo : INT_MAX);
            const int ohi = __reduce_max_sync(FULL, walking ? fo : INT_MIN);
            const int vlo = __reduce_min_sync(FULL, walking ? fv : INT_MAX);
            const int vhi = __reduce_max_sync(FULL, walking ? fv : INT_MIN);
            if (ws.lane == 0) ws.bundle[i] = make_float4(ord2f(olo), ord2f(ohi), ord2f(vlo), ord2f(vhi));
        }
        __syncwarp();
    }
#endif

    double lt = DBL_MAX;
    int lret = 0, lid = -1, lwin = -1;
    int s_node[KD_STACK];
    double s_tl[KD_STACK], s_tu[KD_STACK], s_guard[KD_STACK];
    int sp = 0, ni = 0, dim = -1;
    bool have = true, after_leaf = false;

    while (true) {
        /* 1. advance to the next leaf (kd_node_intersect, kd-tree.c:482-568) */
        int leaf_first = 0, leaf_count = 0, leaf_node = -1;
        while (walking && leaf_node < 0) {
            if (!after_leaf) {
                if (!have) {
                    if (sp == 0) { walking = false; break; }
                    --sp;
                    ni = s_node[sp]; tl = s_tl[sp]; tu = s_tu[sp];
                    if (!(lt > s_guard[sp])) continue;
                    have = true;
                }
                if (ni < 0 || tu < 0.0) { have = false; continue; }
                const ndt_flat_node *nd = sc.nodes + ni;
                dim = NDT_LDG(&nd->dim);
                const int lcount = NDT_LDG(&nd->leaf_count);
                after_leaf = true;
                if (lcount > 0) {
                    leaf_node = ni;
                    leaf_first = NDT_LDG(&nd->leaf_begin);
                    leaf_count = lcount;
                }
                continue;
            }
            after_leaf = false;
            if (dim < 0) { have = false; continue; }
            const ndt_flat_node *nd = sc.nodes + ni;
            int nr = NDT_LDG(&nd->left), fr = NDT_LDG(&nd->right);
            const double b = NDT_LDG(&nd->boundary);
            const double vi = vinv[dim], oi = o_dyn[dim];
            if (vi < EPS2) { int x = nr; nr = fr; fr = x; }
            if (-INV_EPS2 <= vi && vi <= INV_EPS2) {
                double tp = (b - oi) * vi;
                if (tu < tp - EPS && lt > tl) {
                    ni = nr;
                } else if (tl > tp + EPS && lt > tl) {
                    ni = fr;
                } else {
                    if (sp >= KD_STACK) { overflow = 1; have = false; continue; }
                    s_node[sp] = fr; s_tl[sp] = tp - EPS; s_tu[sp] = tu; s_guard[sp] = tp; ++sp;
                    if (lt > tl) { ni = nr; tu = tp + EPS; }
                    else have = false;
                }
            } else {
                bool go_near = (oi < b + EPS) && (lt > tl);
                if (oi > b - EPS) {
                    if (sp >= KD_STACK) { overflow = 1; have = false; continue; }
                    s_node[sp] = fr; s_tl[sp] = tl; s_tu[sp] = tu; s_guard[sp] = tl; ++sp;
                }
                if (go_near) ni = nr;
                else have = false;
            }
        }

        /* 2. leaves, one distinct leaf at a time */
        unsigned waiting = __ballot_sync(FULL, leaf_node >= 0);
        if (!waiting) break;
        while (waiting) {
            const int src = __ffs(waiting) - 1;
            const int L = __shfl_sync(FULL, leaf_node, src);
            const int first = __shfl_sync(FULL, leaf_first, src);
            const int count = __shfl_sync(FULL, leaf_count, src);
            const bool mine = leaf_node == L;
            waiting &= ~__ballot_sync(FULL, mine);
            int oid, owin;
            const double lmd = warp_leaf<NP, false, BIG>(sc, ws, first, count, mine, mb, o, v, obox, vbox, keep_mask, dist_limit, oid, owin);
            if (mine && !(lmd < 0)) {
                lret = 1;
                if (lmd < lt) {          /* trace sets t only when min_dist > EPS, which holds here */
                    lt = lmd;
                    lid = oid;
                    lwin = owin;
                }
                /* the DIRECTIONAL shadow test consumes only the return value (ndt.c:241-249),
                 * an OR over the leaves (kd-tree.c:594,607,616): final once one reports a hit */
                if (only_found) walking = false;
            }
        }
    }

    if (lret) {
        if (!ret || (lt > EPS && lt + EPS < t)) {   /* kd-tree.c:612-617 */
            out.id = lid;
            out.win = lwin;
            ret |= lret;
            md = lt;
        }
    }
    out.found = ret;
    out.t = md;
}

/* the whole query (k_trace_rays, k_generation): lanes with !want take part in the staging only */
template <int NP, bool BIG>
__device__ __forceinline__ void trace_kd_warp(const Scene &sc, WarpStage<NP> &ws, Mailbox &mb, bool want,
                                              const double *o, const double *v, double dist_limit,
                                              Hit &out, int &overflow, int dir_light)
{
    PreWalk pw;
    pw.md = -1; pw.tl = pw.tu = 0; pw.id = pw.win = -1; pw.ret = 0; pw.walking = false;
    if (want) pre_walk<NP, false>(sc, o, v, dist_limit, dir_light >= 0, pw);
    trace_kd_resume<NP, BIG>(sc, ws, mb, pw.walking, o, v, dist_limit, pw, out, overflow, dir_light);
}

/* process_ray (wave.cuh) for a warp: every lane runs the same sequence of
 * queries (the ray, then one per light), lanes without one idle through it */
template <int NP>
__device__ __forceinline__ void process_ray_warp(const Scene &sc, WarpStage<NP> &ws, Mailbox &mb, bool active,
                                                 const double *src, const double *look, double frac, int depth,
                                                 RayRec &rec, Spawn<NP> &sp, int &prim_hit, int &prim_id,
                                                 double &prim_dist, uint32_t &n_shadow, int &overflow)
{
    Tally<false> none;
    Shade<NP> S;
    shade_init<NP>(S, rec, sp);
    n_shadow = 0;
    const int nl = sc.n_lights;
    for (int it = -1; it < nl; ++it) {
        const bool want = active && shade_setup<NP, false>(sc, S, it, src, look, n_shadow, none);
        if (!__ballot_sync(FULL, want)) {
            if (it < 0) break;      /* nobody traces the ray itself: nothing is shaded */
            continue;
        }
        Hit T;
        trace_kd_warp<NP, true>(sc, ws, mb, want, S.ro, S.rv, S.limit, T, overflow, S.ltype == NDT_L_DIRECTIONAL ? it : -1);
        if (want) shade_after<NP, false>(sc, S, it, T, src, look, rec, prim_hit, prim_id, prim_dist, none);
        if (it < 0 && !__ballot_sync(FULL, active && S.shaded)) break;
    }
    if (active) shade_finish<NP, false>(sc, S, look, frac, depth, rec, sp, none);
}

} /* namespace ndt */
