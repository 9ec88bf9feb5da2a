// Times the per-pivot cost of elim_group<> (hybvio_b200/csrc/ekf_elim.cuh) on one CTA of 512 threads.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../hybvio_b200/csrc/ekf_elim.cuh"

__global__ void __launch_bounds__(512) k_elim(double* out, long long* cyc, int n, int W, int reps)
{
    __shared__ double s_elim[ELIM_SMEM_DOUBLES];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    double t[ELIM_RA][2][ELIM_CJ];
    for (int aa = 0; aa < ELIM_RA; aa++) for (int sr = 0; sr < 2; sr++) for (int bb = 0; bb < ELIM_CJ; bb++) {
        int i = elim_row(wrp, aa, sr), j = lane + 32 * bb;
        t[aa][sr][bb] = (i < n && j < W) ? ((i == j) ? 100.0 + i : 1.0 / (1 + i + j)) : 0.0;   // diagonally dominant SPD-ish
    }
    __syncthreads();
    long long t0 = clock64();
    bool ok = true;
    for (int rep = 0; rep < reps; rep++) {
    if (rep > 0) for (int aa = 0; aa < ELIM_RA; aa++) for (int sr = 0; sr < 2; sr++) for (int bb = 0; bb < ELIM_CJ; bb++) { int i = elim_row(wrp, aa, sr), j = lane + 32 * bb; t[aa][sr][bb] = (i < n && j < W) ? ((i == j) ? 100.0 + i : 1.0 / (1 + i + j)) : 0.0; }
    ok = ok && elim_dispatch(t, n, W, lane, wrp, s_elim);
    __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = ok; }
    double s = 0; for (int aa = 0; aa < ELIM_RA; aa++) for (int sr = 0; sr < 2; sr++) for (int bb = 0; bb < ELIM_CJ; bb++) s += t[aa][sr][bb];
    out[tid] = s;
}
int main(int argc, char** argv)
{
    int reps = argc > 1 ? atoi(argv[1]) : 1;
    double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
    for (int n : {8, 20, 40, 84}) {
        int W = n + 21;
        for (int r = 0; r < 2; r++) { k_elim<<<1, 512>>>(out, cyc, n, W, reps); cudaDeviceSynchronize(); }
        long long h[2]; cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("n=%2d W=%3d: %lld cycles total, %lld per pivot (ok=%lld)\n", n, W, h[0], h[0] / n / reps, h[1]);
    }
    return 0;
}
