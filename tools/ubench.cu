// Micro-benchmarks that size the EKF elimination's critical path on B200 (results quoted in DESIGN.md).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, double x0)
{
    __shared__ double sh[64];
    double x = x0 + threadIdx.x * 1e-9;
    long long t0, t1;
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = fma(x, 1.0000001, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 256;
    // dependent double division
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) x = 1.0 / (x + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 64;
    // dependent rcp.approx + 2 Newton steps
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { double d = x + 1.5, r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); r = fma(fma(-d, r, 1.0), r, r); r = fma(fma(-d, r, 1.0), r, r); x = r; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 64;
    // __syncthreads round
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 64;
    // smem publish -> barrier -> read (the per-step handshake)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { if ((threadIdx.x >> 5) == (i & 15)) sh[threadIdx.x & 31] = x; __syncthreads(); x += sh[(i * 7) & 31]; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 64;
    // dependent sqrt
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) x = sqrt(x + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0) / 64;
    // throughput: independent DFMAs, all warps
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
                                    a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9); }
    __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    out[threadIdx.x] = x + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
    for (int nt : {32, 512}) {
        k_lat<<<1, nt>>>(out, cyc, 1.0); cudaDeviceSynchronize();
        k_lat<<<1, nt>>>(out, cyc, 1.0); cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads=%3d  DFMA dep latency %lld cyc | 1.0/x %lld | rcp.approx+2 Newton %lld | __syncthreads %lld | publish+barrier+read %lld | sqrt %lld | "
               "2048 DFMA/thread x %d threads in %lld cyc = %.1f DFMA/clk/SM\n", nt, h[0], h[1], h[2], h[3], h[4], h[5], nt, h[6], 2048.0 * nt / h[6]);
    }
    return 0;
}
