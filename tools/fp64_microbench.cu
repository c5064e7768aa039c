// tools/fp64_microbench.cu -- B200 FP64 pipe: dependent-issue latency and throughput
// as a function of warps per SM sub-partition and independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH> __global__ void k(double *sink, int iters, long long *cyc)
{
    double a[CH];
    for (int c = 0; c < CH; ++c) a[c] = threadIdx.x * 1e-9 + 1.0 + c;
    const double m = 1.0000001, add = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) a[c] = __dadd_rn(__dmul_rn(a[c], m), add);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; ++c) s += a[c];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH> void run(int warps_per_sm, double *sink, long long *cyc)
{
    const int iters = 4096;
    k<CH><<<148, warps_per_sm * 32>>>(sink, iters, cyc);
    k<CH><<<148, warps_per_sm * 32>>>(sink, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_iter = (double)h / iters;
    // per SMSP: warps_per_sm/4 warps, each CH*2 FP64 instr per iter
    double inst_per_cyc_smsp = (warps_per_sm / 4.0) * CH * 2 / per_iter;
    printf("warps/SM %2d chains %d: %7.2f cycles/iter  -> %5.3f FP64 warp-instr/cycle/SMSP (peak 0.5), dependent pair latency %.1f\n",
           warps_per_sm, CH, per_iter, inst_per_cyc_smsp, per_iter);
}
int main()
{
    double *sink; long long *cyc;
    cudaMalloc(&sink, 64); cudaMalloc(&cyc, 64);
    for (int w : {4, 8, 12, 16, 24, 32}) {
        run<1>(w, sink, cyc); run<2>(w, sink, cyc); run<4>(w, sink, cyc); run<8>(w, sink, cyc);
    }
    return 0;
}
