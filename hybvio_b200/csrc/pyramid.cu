// hybvio_b200/csrc/pyramid.cu -- fused optical-flow pyramid for sm_100a.
//
// Replaces, for tracker::ImagePyramid::Factory::compute (src/tracker/image_pyramid.cpp:40-48), the reference's
//   cv::buildOpticalFlowPyramid      OCV/video/src/lkpyramid.cpp:726-822
//     cv::pyrDown (8U, 5x5)          OCV/imgproc/src/pyramids.cpp:745-900  ((sum+128)>>8, reflect-101 in the ROI)
//     calcScharrDeriv                OCV/video/src/lkpyramid.cpp:70-151    (int16 Ix,Iy = 32x gradient)
// (OCV = 3rdparty/mobile-cv-suite/opencv/modules). All arithmetic is integer and bit-exact.
//
// ONE launch builds every level of every image in the batch. A CTA owns one 64x64 level-0 tile and the
// matching 32x32 / 16x16 / 8x8 ... tiles of the coarser levels. It stages the level-0 region it needs
// (tile + halo: 22 px for 4 levels) in shared memory with 32-bit coalesced loads, then computes each coarser
// level's region (tile + halo) from the finer one entirely in shared memory -- the halo is recomputed
// redundantly by neighbouring CTAs so no grid-wide dependency between levels exists. Each level's tile is
// written once to HBM: gray as uchar4, Scharr (Ix,Iy) as 16-byte int4 (4 pixels). Level-0 gray is the
// input image itself (the H2D copy lands directly in the level-0 buffer) and is never rewritten.
//
// HBM traffic per 752x480 image, maxLevel 3 (the algorithmic bytes of SURVEY.md 8(d)): read 360,960 +
// write gray L1-3 118,440 + write deriv L0-3 1,917,600 = 2,397,000 B. Halo re-reads hit L2 (the whole input
// is 361 KB).
#include "hv_common.cuh"

#ifndef PYR_NT
#define PYR_NT 256
#endif

#define PYR_MAX_BATCH 60
struct PyrBuildList {
    const HvPyrDesc* table;   // device-resident descriptors of all pyramids of the context
    int n;                    // images in this launch
    unsigned short idx[PYR_MAX_BATCH];   // descriptor index per blockIdx.z
    // Frame already in HBM (e.g. decoded on the device): read level 0 from src[z] (row pitch srcPitch[z] bytes)
    // and let each CTA also write its own 64x64 tile into the level-0 buffer. NULL: the frame was copied (H2D)
    // straight into the level-0 buffer, which is then read in place.
    const uint8_t* src[PYR_MAX_BATCH];
    int srcPitch[PYR_MAX_BATCH];
};

struct Span { int o0, o1;   // owned output range [o0, o1) at this level
              int s0, s1; };  // stored (shared-memory) range [s0, s1] inclusive, in level coordinates

// Stored span of level k given the stored span of level k+1 (or none for the top level).
__device__ __forceinline__ void span_close(Span& s, int a, int b, int len)
{
    // a..b = unreflected needed range; close it under reflect-101
    int s0 = max(a, 0), s1 = min(b, len - 1);
    if (a < 0) { if (-a > len - 1) { s0 = 0; s1 = len - 1; } else s1 = max(s1, -a); }
    if (b > len - 1) { int r = 2 * (len - 1) - b; if (r < 0) { s0 = 0; s1 = len - 1; } else s0 = min(s0, r); }
    s.s0 = s0; s.s1 = s1;
}

__device__ __forceinline__ int halo_of(int k, int top) { int h = 1; for (int i = top; i > k; --i) h = 2 * h + 2; return h; }

// Scharr + store for the owned tile of level k. buf holds the stored region [sx.s0a.., sy.s0..] with pitch bp.
__device__ __forceinline__ void emit_level(const HvLevel& L, const uint8_t* buf, int bp, int bx0, int by0,
                                           const Span& sx, const Span& sy, bool writeGray, int itemW)
{
    const int w = L.w, h = L.h;
    const int tw = sx.o1 - sx.o0, th = sy.o1 - sy.o0;
    const int itemsX = (tw + itemW - 1) / itemW;
    const int nItems = itemsX * th;
    for (int it = threadIdx.x; it < nItems; it += PYR_NT) {
        const int iy = it / itemsX, ix = it - iy * itemsX;
        const int y = sy.o0 + iy, gx = sx.o0 + ix * itemW;
        const uint8_t* r0 = buf + (hv_reflect101(y - 1, h) - by0) * bp - bx0;
        const uint8_t* r1 = buf + (y - by0) * bp - bx0;
        const uint8_t* r2 = buf + (hv_reflect101(y + 1, h) - by0) * bp - bx0;
        int t0[6], t1[6], c[4];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            if (i < itemW + 2) {
                // columns past the owned range (partial item at the right edge) are clamped: their results are discarded
                int x = hv_reflect101(min(gx - 1 + i, sx.o1), w);
                int a = r0[x], b = r1[x], d = r2[x];
                t0[i] = (a + d) * 3 + b * 10;
                t1[i] = d - a;
                if (i >= 1 && i <= 4) c[i - 1] = b;
            }
        }
        short2 g[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < itemW) {
                g[i].x = (short)(t0[i + 2] - t0[i]);
                g[i].y = (short)((t1[i] + t1[i + 2]) * 3 + t1[i + 1] * 10);
            }
        }
        short2* drow = L.deriv + (size_t)y * L.dpitch + gx;
        uint8_t* grow = L.gray + (size_t)y * L.gpitch + gx;
        if (itemW == 4 && gx + 3 < sx.o1) {
            int4 v;
            v.x = *reinterpret_cast<int*>(&g[0]); v.y = *reinterpret_cast<int*>(&g[1]);
            v.z = *reinterpret_cast<int*>(&g[2]); v.w = *reinterpret_cast<int*>(&g[3]);
            *reinterpret_cast<int4*>(drow) = v;
            if (writeGray) *reinterpret_cast<uchar4*>(grow) = make_uchar4(c[0], c[1], c[2], c[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < itemW && gx + i < sx.o1) { drow[i] = g[i]; if (writeGray) grow[i] = (uint8_t)c[i]; }
        }
    }
}

__global__ void __launch_bounds__(PYR_NT) hv_pyr_fused_kernel(PyrBuildList list)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const HvPyrDesc& P = list.table[list.idx[blockIdx.z]];
    const int nl = P.nlevels, top = nl - 1;
    const int tx = blockIdx.x, ty = blockIdx.y;
    if (tx * HV_PYR_TILE >= P.lv[0].w || ty * HV_PYR_TILE >= P.lv[0].h) return;

    // ---- geometry (identical in every thread)
    Span sx[HV_MAX_LEVELS], sy[HV_MAX_LEVELS];
    int bufOff[HV_MAX_LEVELS], bufPitch[HV_MAX_LEVELS];
    {
        int off = 0;
        for (int k = 0; k < nl; k++) {
            int rw = (HV_PYR_TILE >> k) + 2 * halo_of(k, top) + 4;   // +4: slack for 4-byte aligned start
            rw = (rw + 3) & ~3;
            bufOff[k] = off; bufPitch[k] = rw;
            off += rw * rw; off = (off + 15) & ~15;
        }
        int ax = 0, bx = 0, ay = 0, by = 0;
        for (int k = top; k >= 0; k--) {
            const int w = P.lv[k].w, h = P.lv[k].h;
            sx[k].o0 = (tx * HV_PYR_TILE) >> k; sx[k].o1 = min(w, ((tx + 1) * HV_PYR_TILE) >> k);
            sy[k].o0 = (ty * HV_PYR_TILE) >> k; sy[k].o1 = min(h, ((ty + 1) * HV_PYR_TILE) >> k);
            int nax = sx[k].o0 - 1, nbx = sx[k].o1, nay = sy[k].o0 - 1, nby = sy[k].o1;
            if (k < top) { nax = min(nax, 2 * ax - 2); nbx = max(nbx, 2 * bx + 2); nay = min(nay, 2 * ay - 2); nby = max(nby, 2 * by + 2); }
            span_close(sx[k], nax, nbx, w);
            span_close(sy[k], nay, nby, h);
            ax = sx[k].s0; bx = sx[k].s1; ay = sy[k].s0; by = sy[k].s1;
        }
    }

    // ---- stage the level-0 region with 32-bit loads (gpitch is a multiple of 4: rows are 4-byte aligned)
    const uint8_t* ext = list.src[blockIdx.z];
    {
        const HvLevel& L0 = P.lv[0];
        const int x0a = sx[0].s0 & ~3;
        const int words = ((sx[0].s1 - x0a) >> 2) + 1, rows = sy[0].s1 - sy[0].s0 + 1;
        uint8_t* b0 = smem + bufOff[0];
        const int bp = bufPitch[0];
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
        const bool aligned = ext == nullptr || ((((size_t)ext) | (size_t)list.srcPitch[blockIdx.z]) & 3) == 0;
        const uint8_t* base = ext ? ext : L0.gray;
        const int pitch = ext ? list.srcPitch[blockIdx.z] : L0.gpitch;
        if (aligned && (ext == nullptr || x0a + words * 4 <= pitch)) {
            for (int r = wrp; r < rows; r += PYR_NT / 32) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(base + (size_t)(sy[0].s0 + r) * pitch + x0a);
                uint32_t* dst = reinterpret_cast<uint32_t*>(b0 + r * bp);
                for (int c = lane; c < words; c += 32) dst[c] = __ldg(src + c);
            }
        } else {   // arbitrary external pitch / alignment: byte loads, never past the last image column
            const int nbytes = min(words * 4, L0.w - x0a);
            for (int r = wrp; r < rows; r += PYR_NT / 32) {
                const uint8_t* src = base + (size_t)(sy[0].s0 + r) * pitch + x0a;
                for (int c = lane; c < nbytes; c += 32) b0[r * bp + c] = __ldg(src + c);
            }
        }
        sx[0].s0 = x0a;   // stored origin is the aligned one
    }
    __syncthreads();
    emit_level(P.lv[0], smem + bufOff[0], bufPitch[0], sx[0].s0, sy[0].s0, sx[0], sy[0], ext != nullptr, 4);

    // ---- coarser levels: 5x5 [1 4 6 4 1]^2, (sum+128)>>8, reflect-101 inside the finer level
    for (int k = 1; k < nl; k++) {
        const uint8_t* src = smem + bufOff[k - 1];
        const int sp = bufPitch[k - 1], sx0 = sx[k - 1].s0, sy0 = sy[k - 1].s0;
        const int sw = P.lv[k - 1].w, sh = P.lv[k - 1].h;
        uint8_t* dst = smem + bufOff[k];
        const int dp = bufPitch[k];
        const int cw = sx[k].s1 - sx[k].s0 + 1, ch = sy[k].s1 - sy[k].s0 + 1;
        for (int it = threadIdx.x; it < cw * ch; it += PYR_NT) {
            const int iy = it / cw, ix = it - iy * cw;
            const int cx = sx[k].s0 + ix, cy = sy[k].s0 + iy;
            int xs[5];
#pragma unroll
            for (int i = 0; i < 5; i++) xs[i] = hv_reflect101(2 * cx - 2 + i, sw) - sx0;
            int acc = 0;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const uint8_t* row = src + (hv_reflect101(2 * cy - 2 + j, sh) - sy0) * sp;
                int hsum = row[xs[2]] * 6 + (row[xs[1]] + row[xs[3]]) * 4 + row[xs[0]] + row[xs[4]];
                acc += hsum * (j == 2 ? 6 : (j == 1 || j == 3) ? 4 : 1);
            }
            dst[iy * dp + ix] = (uint8_t)((acc + 128) >> 8);
        }
        __syncthreads();
        emit_level(P.lv[k], dst, dp, sx[k].s0, sy[k].s0, sx[k], sy[k], true, min(4, HV_PYR_TILE >> k));
    }
}

// Host-side launch helper (called from capi.cu). Shared memory is sized for the deepest pyramid in the list.
size_t hv_pyr_smem_bytes(int nlevels)
{
    size_t off = 0;
    for (int k = 0; k < nlevels; k++) {
        int h = 1; for (int i = nlevels - 1; i > k; --i) h = 2 * h + 2;
        int rw = (HV_PYR_TILE >> k) + 2 * h + 4; rw = (rw + 3) & ~3;
        off += (size_t)rw * rw; off = (off + 15) & ~(size_t)15;
    }
    return off;
}

cudaError_t hv_launch_pyr_fused(const HvPyrDesc* table, const unsigned short* idx, const uint8_t* const* src, const int* srcPitch,
                                int n, int w0, int h0, int maxNlevels, cudaStream_t stream)
{
    static bool attrSet = false;
    size_t smem = hv_pyr_smem_bytes(maxNlevels);
    if (!attrSet) {
        cudaError_t e = cudaFuncSetAttribute(hv_pyr_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attrSet = true;
    }
    for (int base = 0; base < n; base += PYR_MAX_BATCH) {
        PyrBuildList list;
        list.table = table; list.n = min(PYR_MAX_BATCH, n - base);
        for (int i = 0; i < list.n; i++) {
            list.idx[i] = idx[base + i];
            list.src[i] = src ? src[base + i] : nullptr;
            list.srcPitch[i] = src ? srcPitch[base + i] : 0;
        }
        dim3 grid((w0 + HV_PYR_TILE - 1) / HV_PYR_TILE, (h0 + HV_PYR_TILE - 1) / HV_PYR_TILE, list.n);
        hv_pyr_fused_kernel<<<grid, PYR_NT, smem, stream>>>(list);
    }
    return cudaGetLastError();
}
