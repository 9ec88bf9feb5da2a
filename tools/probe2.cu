// tools/probe2.cu -- round-1 probes that size the next EKF kernel and the host round trips (results quoted in DESIGN.md):
//   * DMMA (mma.sync m8n8k4 f64) throughput / dependent latency vs DFMA
//   * cluster size 8 vs 16 (non-portable): can it launch with 512 threads + 100 KB, cluster.sync cost, DSMEM read latency
//   * fp64 rsqrt / rcp chains
//   * host round trips: kernel -> D2H copy -> stream sync, vs kernel writing mapped pinned memory + host polling
//   * pinned H2D of one stereo pair (2 x 361 KB)
#include <cstdio>
#include <chrono>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma(double* out, long long* cyc, double x0)
{
    double a = x0 + threadIdx.x * 1e-9, b = 1.0000001;
    double c[16];
    for (int i = 0; i < 16; i++) c[i] = i;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) dmma(c[0], c[1], a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 256;          // dependent latency
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma(c[2 * j], c[2 * j + 1], a, b);
    }
    __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;                  // 256 * 8 DMMA per warp
    double s = 0; for (int i = 0; i < 16; i++) s += c[i];
    out[threadIdx.x] = s;
}
__global__ void k_chain(double* out, long long* cyc, double x0)
{
    double x = x0 + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) x = rsqrt(x + 2.0);
    long long t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 64;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { double d = x + 1.5, r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); r = fma(fma(-d, r, 1.0), r, r); r = fma(fma(-d, r, 1.0), r, r); x = r; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 64;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) x = __shfl_sync(0xffffffffu, x, (i * 5) & 31) + 1.0;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 64;
    out[threadIdx.x] = x;
}

// cluster probe: cluster.sync latency, DSMEM dependent-read latency, DSMEM block read
__global__ void __launch_bounds__(512) k_cluster(double* out, long long* cyc)
{
    extern __shared__ double sm[];
    cg::cluster_group cl = cg::this_cluster();
    const int r = cl.block_rank(), nb = cl.num_blocks();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 4095);
    cl.sync();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 32; i++) cl.sync();
    long long t1 = clock64();
    if (r == 0 && threadIdx.x == 0) cyc[0] = (t1 - t0) / 32;
    // dependent DSMEM reads (pointer chase through the neighbour's shared memory)
    const double* rem = cl.map_shared_rank(sm, (r + 1) % nb);
    double idx = 0;
    t0 = clock64();
    if (threadIdx.x == 0) {
#pragma unroll 1
        for (int i = 0; i < 64; i++) idx = rem[(int)idx];
    }
    t1 = clock64();
    if (r == 0 && threadIdx.x == 0) cyc[1] = (t1 - t0) / 64;
    // block read: all 512 threads pull 4096 doubles (32 KB) from the neighbour
    cl.sync();
    t0 = clock64();
    double acc = 0;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) acc += rem[i];
    __syncthreads();
    t1 = clock64();
    if (r == 0 && threadIdx.x == 0) cyc[2] = t1 - t0;
    // same volume from global (L2)
    double* g = out + 8192 + (size_t)((r + 1) % nb) * 4096;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[8192 + (size_t)r * 4096 + i] = sm[i];
    cl.sync();
    t0 = clock64();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) acc += g[i];
    __syncthreads();
    t1 = clock64();
    if (r == 0 && threadIdx.x == 0) cyc[3] = t1 - t0;
    cl.sync();
    out[r * 512 + threadIdx.x] = acc + idx;
}

__global__ void k_flag(volatile int* hostFlag, double* hostRes, int v) { hostRes[threadIdx.x] = v; __threadfence_system(); if (threadIdx.x == 0) *hostFlag = v; }
__global__ void k_res(double* res, int v) { res[threadIdx.x] = v; }

static int try_cluster(int size, size_t smem, double* out, long long* cyc)
{
    cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (size > 8) { cudaError_t e = cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); if (e) { printf("cluster %d: non-portable attr: %s\n", size, cudaGetErrorString(e)); return 1; } }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(size); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = size; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int nclusters = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_cluster, &cfg);
    printf("cluster %2d x 512 thr x %zu KB smem: maxActiveClusters=%d (%s)\n", size, smem / 1024, nclusters, cudaGetErrorString(e));
    for (int rep = 0; rep < 2; rep++) {
        e = cudaLaunchKernelEx(&cfg, k_cluster, out, cyc);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e || e2) { printf("  launch failed: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError(); return 1; }
    }
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("  cluster.sync %lld cyc | DSMEM dependent read %lld cyc | 32 KB block read: DSMEM %lld cyc, L2 %lld cyc\n", h[0], h[1], h[2], h[3]);
    // launch overhead of a cluster launch: 200 back-to-back
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 200; i++) cudaLaunchKernelEx(&cfg, k_cluster, out, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("  back-to-back launches: %.2f us each (kernel body included)\n", ms * 1000 / 200);
    return 0;
}

int main()
{
    double* out; long long* cyc; cudaMalloc(&out, (8192 + 16 * 4096) * 8 + 16 * 512 * 8); cudaMalloc(&cyc, 64);
    long long h[8];
    for (int nt : {32, 128, 512}) {
        k_dmma<<<1, nt>>>(out, cyc, 1.0); cudaDeviceSynchronize();
        k_dmma<<<1, nt>>>(out, cyc, 1.0); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("DMMA m8n8k4 threads=%3d: dependent latency %lld cyc | %d warps x 2048 DMMA in %lld cyc = %.1f FMA/clk/SM\n", nt, h[0], nt / 32, h[1],
               256.0 * 2048.0 * (nt / 32) / h[1]);
    }
    k_chain<<<1, 32>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("rsqrt(double) dependent %lld cyc | rcp.approx+2 Newton %lld | shfl+add %lld\n", h[0], h[1], h[2]);
    try_cluster(8, 100 * 1024, out, cyc);
    try_cluster(16, 100 * 1024, out, cyc);
    try_cluster(16, 48 * 1024, out, cyc);

    // host round trips
    using clk = std::chrono::steady_clock;
    cudaStream_t s; cudaStreamCreate(&s);
    double* dres; cudaMalloc(&dres, 256);
    double* hres; cudaMallocHost(&hres, 256);
    int* hflag; cudaHostAlloc(&hflag, 64, cudaHostAllocMapped);
    double* hmap; cudaHostAlloc(&hmap, 256, cudaHostAllocMapped);
    int* dflag; double* dmap; cudaHostGetDevicePointer(&dflag, hflag, 0); cudaHostGetDevicePointer(&dmap, hmap, 0);
    for (int mode = 0; mode < 2; mode++) {
        double best = 1e9, sum = 0; const int reps = 200;
        for (int i = 1; i <= reps + 20; i++) {
            auto t0 = clk::now();
            if (mode == 0) { k_res<<<1, 32, 0, s>>>(dres, i); cudaMemcpyAsync(hres, dres, 24, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }
            else { k_flag<<<1, 32, 0, s>>>(dflag, dmap, i); while (*(volatile int*)hflag != i) { } }
            double us = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
            if (i > 20) { sum += us; if (us < best) best = us; }
        }
        printf("round trip (%s): mean %.2f us, best %.2f us\n", mode == 0 ? "kernel + 24 B D2H copy + stream sync" : "kernel writes mapped pinned memory, host polls", sum / reps, best);
    }
    // H2D of a stereo pair from pinned memory
    unsigned char* hp; cudaMallocHost(&hp, 2 * 360960); unsigned char* dp; cudaMalloc(&dp, 2 * 360960);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a, s);
        for (int i = 0; i < 50; i++) { cudaMemcpyAsync(dp, hp, 360960, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(dp + 360960, hp + 360960, 360960, cudaMemcpyHostToDevice, s); }
        cudaEventRecord(b, s); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep == 2) printf("pinned H2D of one stereo pair (2 x 361 KB): %.2f us (%.1f GB/s)\n", ms * 1000 / 50, 2 * 360960 / (ms * 1e-3 / 50) / 1e9);
    }
    auto t0 = clk::now();
    for (int i = 0; i < 200; i++) cudaMemcpyAsync(dres, hres, 24, cudaMemcpyHostToDevice, s);
    double us = std::chrono::duration<double, std::micro>(clk::now() - t0).count();
    cudaStreamSynchronize(s);
    printf("host cost of one small cudaMemcpyAsync H2D call: %.2f us\n", us / 200);
    return 0;
}
