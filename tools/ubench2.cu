#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n)
{
    __shared__ double sh[64];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    double x = 1.0 + threadIdx.x * 1e-3;
    long long t0, t1;
    // (a) dependent double shuffles, runtime source lane, all warps
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) x = __shfl_sync(0xffffffffu, x, (i * 7 + n) & 31) + 1e-9;
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 64;
    // (b) the same but only inside the warp that "owns" step i (others skip), barrier each step
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { if (wrp == (i & 15)) { x = __shfl_sync(0xffffffffu, x, (i * 7 + n) & 31) + 1e-9; if (lane == 0) sh[i & 31] = x; } __syncthreads(); x += sh[i & 31] * 1e-9; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 64;
    // (c) owner does shfl + division + lane-0 store, barrier
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { if (wrp == (i & 15)) { double p = __shfl_sync(0xffffffffu, x, (i * 7 + n) & 31); double r = 1.0 / (p + 2.0); if (lane == 0) { sh[i & 31] = r; sh[32 + (i & 31)] = p; } } __syncthreads(); x += sh[i & 31] * 1e-9; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 64;
    // (d) as (c) but the owner is always warp 0
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { if (wrp == 0) { double p = __shfl_sync(0xffffffffu, x, (i * 7 + n) & 31); double r = 1.0 / (p + 2.0); if (lane == 0) { sh[i & 31] = r; sh[32 + (i & 31)] = p; } } __syncthreads(); x += sh[i & 31] * 1e-9; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 64;
    out[threadIdx.x] = x;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
    for (int nt : {32, 512}) {
        k<<<1, nt>>>(out, cyc, 3); cudaDeviceSynchronize();
        k<<<1, nt>>>(out, cyc, 3); cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads=%3d: dep shfl(double)+add %lld cyc | rotating-owner shfl+publish+barrier+read %lld | + division %lld | fixed owner (warp 0) %lld\n", nt, h[0], h[1], h[2], h[3]);
    }
    return 0;
}
