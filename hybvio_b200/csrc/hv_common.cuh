// hybvio_b200 -- shared device-side types for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HV_MAX_LEVELS 6      // pyrLKMaxLevel <= 5 (HybVIO default 3, codegen/parameter_definitions.c:334-344)
#define HV_PYR_TILE 64       // level-0 tile edge of the fused pyramid kernel; must be divisible by 2^maxLevel

// One pyramid level in HBM. Levels are stored UNPADDED (the reference pads every level by winSize on all
// sides, OCV/video/src/lkpyramid.cpp:761-808; here the reflect-101 / zero border is applied by index
// arithmetic in the LK kernel instead, which saves 0.79 MB of writes per 752x480 image).
struct HvLevel {
    uint8_t* gray;    // w x h, row pitch gpitch bytes (multiple of 4; level 0: w when w % 4 == 0, else / coarser levels: multiple of 128)
    short2*  deriv;   // w x h (Ix, Iy) Scharr x32, row pitch dpitch elements (multiple of 32 => 128 B)
    int w, h;
    int gpitch;
    int dpitch;
};

struct HvPyrDesc {
    HvLevel lv[HV_MAX_LEVELS];
    int nlevels;
    int win;
};

// BORDER_REFLECT_101 (OCV/core/src/copy.cpp:1136-1181)
__host__ __device__ __forceinline__ int hv_reflect101(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// ---- Lucas-Kanade launch description (lk.cu)
#define LK_WARPS_PER_CTA 4
#define LK_MAX_JOBS 8

struct LkJob {
    int prevIdx, nextIdx;       // pyramid descriptor indices
    int n;                      // number of features
    int useInitial;             // OPTFLOW_USE_INITIAL_FLOW
    const float2* prevPts;      // device
    float2* nextPts;            // device; in: initial guess (if useInitial), out: end point
    uint8_t* status;            // device; OpenCV status 1 = ok, 0 = failed
    int32_t* trackStatus;       // device, optional; tracker::Feature::Status (src/tracker/track.hpp:9-20)
    const float2* initPts;      // device, optional: the initial guess lives here instead of in nextPts (which is then only written)
};

struct LkLaunch {
    const HvPyrDesc* table;
    LkJob jobs[LK_MAX_JOBS];
    int njobs;
    int maxLevel;
    int maxIter;                // already clamped to [0,100]
    double eps2;                // already clamped and squared
    float minEig;
    // Completion signal for the host-buffer API (single job, CTA-per-feature kernel): every CTA bumps *doneCounter after
    // its results are visible system-wide; the one that reaches doneTarget stores seq into *hostFlag (mapped pinned host
    // memory), which the host polls instead of a D2H copy + stream synchronisation. NULL: no signal.
    unsigned* doneCounter;
    unsigned doneTarget, seq;
    volatile unsigned* hostFlag;
    int prefetch;               // CTA-per-feature kernel: request the search region of a level with cp.async before the template patch is loaded
};


// Kernel choice of hv_launch_lk: up to 640 features in a launch (one session) a CTA of 4 warps per feature (minimal latency; the
// only kernel that raises the host flag of the polled path), above that a warp per feature (throughput). Both produce identical bits.
inline bool hv_lk_uses_cta_kernel(long long totalFeatures) { return totalFeatures <= 640; }
cudaError_t hv_launch_lk(const LkLaunch& L, int win, cudaStream_t stream);

// ---- corner detection launch description (gftt.cu)
struct GfttArgs {
    const uint8_t* gray; int pitch, w, h;
    int cell;                 // bs
    float k0, k1;             // [1 2 1] * scale as the fp32 kernel OpenCV builds (k0 = 2 s, k1 = s)
    float minResponse;
    float* kp;                // cellsX * cellsY * (x, y, response * 16); may be mapped pinned host memory
    unsigned* doneCounter; unsigned doneTarget, seq; volatile unsigned* hostFlag;      // polled completion (like the LK kernel), optional
};

cudaError_t hv_launch_gftt(const GfttArgs& a, cudaStream_t stream);

// ---- frame ingest (ingest.cu)
#define HV_REMAP_INVALID (-32768)
struct HvRemapEntry { short x0, y0; float xfrac, yfrac; };      // 12 bytes per output pixel (hv_remap_entry of the C ABI)
cudaError_t hv_launch_gray(const uint8_t* src, int srcPitch, int channels, int w, int h, const float coeff[4], uint8_t* dst, int dstPitch, cudaStream_t s);
cudaError_t hv_launch_remap(const uint8_t* src, int srcPitch, int w, int h, const HvRemapEntry* table, uint8_t* dst, int dstPitch, cudaStream_t s);
