// tools/ubench_elim2.cu -- cycle breakdown of ek2_block_eliminate (hybvio_b200/csrc/ekf_cluster2.cuh) on one CTA:
// where does a block step go: rows solve, barrier, warp 0's look-ahead tile + diagonal factorisation, barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ihybvio_b200/csrc -o tools/ubench_elim2 tools/ubench_elim2.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ unsigned long long g_acc[8][16];       // [mark transition][warp]; accumulated with fire-and-forget reductions
#define EK2_ELIM_DECL long long elim_t_ = clock64();
#define EK2_ELIM_MARK(i)                                                                                     \
    do { const long long t_ = clock64(); if (lane == 0) atomicAdd(&g_acc[(i)][wrp], (unsigned long long)(t_ - elim_t_)); elim_t_ = t_; } while (0)
#include "ekf_cluster2.cuh"

__global__ void __launch_bounds__(512) k_elim(const double* Tin, double* Tout, int n, int ncols, int W, long long* total)
{
    extern __shared__ double T[];
    __shared__ double s_linv[EK2_LINV_DOUBLES];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    for (int i = tid; i < n * W; i += 512) T[i] = Tin[i];
    if (lane == 0) for (int i = 0; i < 8; i++) g_acc[i][wrp] = 0;
    __syncthreads();
    const long long t0 = clock64();
    const bool ok = ek2_block_eliminate(T, W, n, ncols, wrp, lane, s_linv, &s_bad);
    __syncthreads();
    if (tid == 0) { total[0] = clock64() - t0; total[1] = ok; }
    __syncthreads();
    for (int i = tid; i < n * W; i += 512) Tout[i] = T[i];
}

int main()
{
    for (int n : {8, 20, 40, 84}) {
        const int B = 20, ncols = n + B + 1, W = ek2_pad4mod16(ncols);
        std::vector<double> M((size_t)n * n), T((size_t)n * W, 0.0);
        srand(1);
        for (auto& x : M) x = rand() / (double)RAND_MAX - 0.5;
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double s = 0; for (int k = 0; k < n; k++) s += M[i * n + k] * M[j * n + k]; T[(size_t)i * W + j] = s + (i == j ? 25.0 : 0.0); }
        for (int i = 0; i < n; i++) for (int j = n; j < ncols; j++) T[(size_t)i * W + j] = rand() / (double)RAND_MAX;
        double* d; double* dout; long long* tot; cudaMalloc(&d, T.size() * 8); cudaMalloc(&dout, T.size() * 8); cudaMalloc(&tot, 16);
        cudaMemcpy(d, T.data(), T.size() * 8, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(k_elim, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int rep = 0; rep < 3; rep++) k_elim<<<1, 512, T.size() * 8>>>(d, dout, n, ncols, W, tot);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; unsigned long long acc[8][16];
        cudaMemcpy(h, tot, 16, cudaMemcpyDeviceToHost);
        cudaMemcpyFromSymbol(acc, g_acc, sizeof(acc));
        // check against a host Cholesky: Y_out = L^-1 Y
        std::vector<double> R(T.size()), L((size_t)n * n, 0.0);
        cudaMemcpy(R.data(), dout, T.size() * 8, cudaMemcpyDeviceToHost);
        for (int j = 0; j < n; j++) {
            double dj = T[(size_t)j * W + j];
            for (int k = 0; k < j; k++) dj -= L[j * n + k] * L[j * n + k];
            L[j * n + j] = sqrt(dj);
            for (int i = j + 1; i < n; i++) { double v = T[(size_t)j * W + i]; for (int k = 0; k < j; k++) v -= L[i * n + k] * L[j * n + k]; L[i * n + j] = v / L[j * n + j]; }
        }
        double worst = 0.0;
        for (int c = n; c < ncols; c++) {
            std::vector<double> z(n);
            for (int i = 0; i < n; i++) { double v = T[(size_t)i * W + c]; for (int k = 0; k < i; k++) v -= L[i * n + k] * z[k]; z[i] = v / L[i * n + i]; }
            for (int i = 0; i < n; i++) worst = fmax(worst, fabs(z[i] - R[(size_t)i * W + c]));
        }
        printf("max |L^-1 Y - host| = %.3e  ", worst);
        const int MT = (n + EK2_EB - 1) / EK2_EB;      // blocks
        printf("n=%2d (%2d blocks): total %6lld cyc = %5.2f us @1.965GHz (%s, ok=%lld) | per block step, warp 0: rows-solve %llu | wait barrier %llu | look-ahead tile %llu | diag factor %llu | wait barrier %llu | loop top %llu"
               "  || warp 5 (worker): rows-solve %llu | barrier %llu | trailing tiles %llu | idle until barrier %llu\n",
               n, MT, h[0], h[0] / 1965.0, cudaGetErrorString(e), h[1],
               acc[1][0] / MT, acc[2][0] / MT, acc[3][0] / (MT > 1 ? MT - 1 : 1), acc[4][0] / (MT > 1 ? MT - 1 : 1), acc[5][0] / (MT > 1 ? MT - 1 : 1), acc[0][0] / MT,
               acc[1][5] / MT, acc[2][5] / MT, acc[3][5] / (MT > 1 ? MT - 1 : 1), (acc[4][5] + acc[5][5]) / (MT > 1 ? MT - 1 : 1));
        cudaFree(d); cudaFree(dout); cudaFree(tot);
    }
    return 0;
}
