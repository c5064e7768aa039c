/*
 * np_inst.cu -- one padded dimension's kernels (gen.cuh); built once per NP with -DNDT_NP=N.
 */
#include "gen.cuh"

#ifndef NDT_NP
#error "build with -DNDT_NP=<4|6|8|10|12>"
#endif
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

namespace {
constexpr int NP = NDT_NP;

template <bool CNT> size_t generation_smem(int boxed) { return CNT ? 0 : (size_t)(BLOCK / 32) * warp_smem_bytes<NP>(boxed); }

template <bool CNT> int occ(int boxed)
{
    int b = 0;
    const size_t sm = generation_smem<CNT>(boxed);
    if (sm > 48 * 1024) cudaFuncSetAttribute(k_generation<NP, CNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_generation<NP, CNT>, BLOCK, sm) != cudaSuccess || b < 1) b = 1;
    return b;
}
int blocks_per_sm(bool cnt, int boxed) { return cnt ? occ<true>(boxed) : occ<false>(boxed); }

void generation(bool cnt, int blocks, cudaStream_t st, const Scene &sc, const GenArgs &a)
{
    if (cnt) k_generation<NP, true><<<blocks, BLOCK, generation_smem<true>(sc.any_boxed), st>>>(sc, a);
    else k_generation<NP, false><<<blocks, BLOCK, generation_smem<false>(sc.any_boxed), st>>>(sc, a);
}

size_t trace_smem(int boxed) { return (size_t)(BLOCK / 32) * warp_smem_bytes<NP>(boxed); }
template <bool BIG> int trace_occ(int boxed)
{
    int b0 = 0, b1 = 0;
    const size_t sm = trace_smem(boxed);
    if (sm > 48 * 1024) {
        cudaFuncSetAttribute(k_trace<NP, 0, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute(k_trace<NP, 1, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_trace<NP, 0, BIG>, BLOCK, sm) != cudaSuccess || b0 < 1) b0 = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_trace<NP, 1, BIG>, BLOCK, sm) != cudaSuccess || b1 < 1) b1 = 1;
    return b0 < b1 ? b0 : b1;
}
int trace_blocks_per_sm(int boxed) { return (boxed & 4) ? trace_occ<true>(boxed) : trace_occ<false>(boxed); }
void trace(int mode, int blocks, cudaStream_t st, const Scene &sc, const WaveArgs &a)
{
    const size_t sm = trace_smem(sc.any_boxed);
    if (sc.any_boxed & 4) {
        if (mode) k_trace<NP, 1, true><<<blocks, BLOCK, sm, st>>>(sc, a);
        else k_trace<NP, 0, true><<<blocks, BLOCK, sm, st>>>(sc, a);
    } else {
        if (mode) k_trace<NP, 1, false><<<blocks, BLOCK, sm, st>>>(sc, a);
        else k_trace<NP, 0, false><<<blocks, BLOCK, sm, st>>>(sc, a);
    }
}
void shade(int phase, int blocks, cudaStream_t st, const Scene &sc, const WaveArgs &a)
{
    if (phase) k_shade<NP, 1><<<blocks, BLOCK, 0, st>>>(sc, a);
    else k_shade<NP, 0><<<blocks, BLOCK, 0, st>>>(sc, a);
}

const void *pre_fn(int mode) { return mode ? (const void *)k_pre<NP, 1> : (const void *)k_pre<NP, 0>; }
void pre(int mode, int blocks, cudaStream_t st, const Scene &sc, const WaveArgs &a)
{
    if (mode) k_pre<NP, 1><<<blocks, BLOCK, 0, st>>>(sc, a);
    else k_pre<NP, 0><<<blocks, BLOCK, 0, st>>>(sc, a);
}
int pre_grid(int sm_count)
{
    int b0 = 0, b1 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_pre<NP, 0>, BLOCK, 0) != cudaSuccess || b0 < 1) b0 = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_pre<NP, 1>, BLOCK, 0) != cudaSuccess || b1 < 1) b1 = 1;
    return (b0 < b1 ? b0 : b1) * sm_count;
}
const void *trace_fn(int mode, int stage)
{
    if (stage & 4) return mode ? (const void *)k_trace<NP, 1, true> : (const void *)k_trace<NP, 0, true>;
    return mode ? (const void *)k_trace<NP, 1, false> : (const void *)k_trace<NP, 0, false>;
}
const void *shade_fn(int phase) { return phase ? (const void *)k_shade<NP, 1> : (const void *)k_shade<NP, 0>; }
int shade_grid(int sm_count, int gen_cap)
{
#if NDT_SHADE_LOOP
    int b0 = 0, b1 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_shade<NP, 0>, BLOCK, 0) != cudaSuccess || b0 < 1) b0 = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_shade<NP, 1>, BLOCK, 0) != cudaSuccess || b1 < 1) b1 = 1;
    (void)gen_cap;
    return (b0 > b1 ? b0 : b1) * sm_count * 2;
#else
    (void)sm_count;
    return (gen_cap + BLOCK - 1) / BLOCK;
#endif
}

const void *light_fn() { return (const void *)k_light<NP>; }
void light(int blocks, cudaStream_t st, const Scene &sc, const WaveArgs &a) { k_light<NP><<<blocks, BLOCK, 0, st>>>(sc, a); }

void pack_leaf(cudaStream_t st, const Scene &sc, int n_refs, void *out, void *box_out, int id_base)
{
    k_pack_leaf<NP><<<(n_refs + 255) / 256, 256, 0, st>>>(sc, n_refs, (LeafRec<NP> *)out, (BoxRec<NP> *)box_out, id_base);
}

void trace_rays(int blocks, cudaStream_t st, const Scene &sc, int n_rays, const double *o, const double *v,
                const double *limits, int32_t *found, int32_t *ids, double *ts, double *hits, double *normals,
                uint32_t *mb_bits, uint32_t mb_stride, uint32_t mb_words, uint32_t mb_shift, int *overflow,
                const void *leafrec, const void *boxrec)
{
    const size_t sm = (size_t)(BLOCK / 32) * warp_smem_bytes<NP>(sc.any_boxed);
    if (sm > 48 * 1024) cudaFuncSetAttribute(k_trace_rays<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k_trace_rays<NP><<<blocks, BLOCK, sm, st>>>(
        sc, n_rays, o, v, limits, found, ids, ts, hits, normals, mb_bits, mb_stride, mb_words, mb_shift, overflow, leafrec, boxrec);
}
}  // namespace

extern const NpOps CAT(ndt_np_ops_, NDT_NP) = { trace_blocks_per_sm, trace, shade, blocks_per_sm, generation, pack_leaf, trace_rays,
                                                     pre_fn, pre, pre_grid, trace_fn, shade_fn, trace_smem, shade_grid, light_fn, light };
