// tests/emu/cuda_emu.h -- a very small host emulation of the CUDA execution model, for testing kernel LOGIC without a
// GPU (this container has none; GPU minutes are scarce). Test infrastructure only -- never part of the product.
//
// Every CUDA thread of ONE CTA is an OS thread; __syncthreads / __syncwarp are std::barrier, full-warp shuffles go
// through a per-warp exchange buffer. `__shared__` becomes `static` (one CTA at a time per process; cluster kernels are
// emulated with one forked process per CTA and MAP_SHARED global memory, see emu_cluster.h). Data races between
// barriers are NOT detected; what this catches is indexing, staging, ownership and protocol mistakes.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define HV_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __cluster_dims__(...)
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct emu_dim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local emu_dim3 threadIdx, blockIdx;
inline emu_dim3 blockDim, gridDim;
struct alignas(16) double2 { double x, y; };

namespace emu {
struct Cta {
    int nthreads, nwarps;
    std::barrier<> bar;
    std::vector<std::unique_ptr<std::barrier<>>> wbar;
    std::vector<double> xch, xch2;     // nwarps x 32 exchange slots
    explicit Cta(int nt) : nthreads(nt), nwarps((nt + 31) / 32), bar(nt), xch((size_t)((nt + 31) / 32) * 32), xch2((size_t)((nt + 31) / 32) * 32)
    {
        for (int w = 0; w < nwarps; w++) wbar.emplace_back(new std::barrier<>(std::min(32, nt - 32 * w)));
    }
};
inline Cta* cta = nullptr;
inline unsigned block_y = 0, block_z = 0;      // blockIdx.y / .z of the CTA launch_cta runs (kernels with 2-D / 3-D grids)

template <class F>
void launch_cta(int nthreads, unsigned bx, F&& body)
{
    Cta c(nthreads);
    cta = &c;
    blockDim.x = nthreads; blockDim.y = blockDim.z = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t] { threadIdx.x = t; threadIdx.y = threadIdx.z = 0; blockIdx.x = bx; blockIdx.y = block_y; blockIdx.z = block_z; body(); });
    for (auto& x : th) x.join();
    cta = nullptr;
}
}  // namespace emu

inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __syncthreads() { emu::cta->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::cta->wbar[threadIdx.x >> 5]->arrive_and_wait(); }
// full-warp shuffle: EVERY lane of the warp must call it (what the kernels written for the emulator do)
inline double __shfl_sync(unsigned, double v, int src)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* slot = emu::cta->xch.data() + (size_t)w * 32;
    slot[lane] = v;
    emu::cta->wbar[w]->arrive_and_wait();
    const double r = slot[src & 31];
    emu::cta->wbar[w]->arrive_and_wait();
    return r;
}
// mma.sync.m8n8k4.f64: a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1] (g = lane >> 2, t = lane & 3); every lane calls it
inline void emu_dmma(double& c0, double& c1, double a, double b)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double* A = emu::cta->xch.data() + (size_t)w * 32;
    double* B = emu::cta->xch2.data() + (size_t)w * 32;
    A[lane] = a; B[lane] = b;
    emu::cta->wbar[w]->arrive_and_wait();
    for (int k = 0; k < 4; k++) { c0 = std::fma(A[g * 4 + k], B[(2 * t) * 4 + k], c0); c1 = std::fma(A[g * 4 + k], B[(2 * t + 1) * 4 + k], c1); }
    emu::cta->wbar[w]->arrive_and_wait();
}
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
template <class T> inline T min(T a, T b) { return a < b ? a : b; }
template <class T> inline T max(T a, T b) { return a > b ? a : b; }
using std::fma;

// ---- additions for the tracker kernels (lk.cu): vector types, read-only loads, rounding-mode intrinsics (the host FPU rounds to
// nearest even; build with -ffp-contract=off so that no multiply-add is contracted), integer warp collectives, atomics
#include <climits>
#include <cfloat>
struct float2 { float x, y; };
struct short2 { short x, y; };
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline float __ll2float_rn(long long v) { return (float)v; }
inline int __float2int_rn(float v) { return (int)std::lrintf(v); }
inline int __float2int_rd(float v) { return (int)std::floor(v); }
inline int __shfl_down_sync(unsigned, int v, int delta)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* slot = emu::cta->xch.data() + (size_t)w * 32;
    slot[lane] = (double)v;
    emu::cta->wbar[w]->arrive_and_wait();
    const int r = lane + delta < 32 ? (int)slot[lane + delta] : v;
    emu::cta->wbar[w]->arrive_and_wait();
    return r;
}
inline int __reduce_add_sync(unsigned, int v)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* slot = emu::cta->xch.data() + (size_t)w * 32;
    slot[lane] = (double)v;
    emu::cta->wbar[w]->arrive_and_wait();
    long long s = 0;
    for (int i = 0; i < 32; i++) s += (long long)slot[i];
    emu::cta->wbar[w]->arrive_and_wait();
    return (int)s;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
struct uchar4 { unsigned char x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
// prmt.b32 in its default mode: selector nibble n picks byte (n & 7) of {b, a} (a = bytes 0-3, b = bytes 4-7); bit 3 replicates the sign
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned n = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(src >> (8 * (n & 7))) & 0xff;
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
inline unsigned char* emu_dynamic_smem = nullptr;   // stands in for `extern __shared__` arrays (set by the harness before a launch)
