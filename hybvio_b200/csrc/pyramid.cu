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
//
// The first generation of this kernel (round 1, removed) evaluated every output pixel on its own: 18 byte loads with a reflect-101
// per tap for a Scharr item, 25 for a pyrDown pixel -- 36,000 warp instructions per CTA (ncu), i.e. issue-bound (8 % of the HBM
// roofline at 32 images per launch; 24.9 us per stereo pair against 16.8 us for this one). hv_pyr_fused2_kernel keeps the data flow and
// removes the instructions: a thread walks DOWN a 4-pixel-wide strip with the three input rows rolling through registers (one aligned 32-bit
// + two byte loads per row instead of 18 byte loads), the Scharr arithmetic runs on two 16-bit lanes per register with biases
// chosen so that every lane stays in [0, 65535] and the bias is exactly 0x8000 (removed by one XOR), pyrDown is separable along a
// vertical strip of outputs (5 loads per source row shared by the outputs of the strip), reflection is one abs + one min
// (levels of at least 8 x 8 pixels; smaller levels and partial 4-pixel items take the first-generation code), and the Scharr
// pass of level k shares its barrier interval with the pyrDown pass that produces level k + 1.
#include "hv_common.cuh"
#include <stdlib.h>
#include <string.h>

#ifndef PYR_NT
#define PYR_NT 256
#endif

#define PYR_MAX_BATCH 32
// 128-byte TMA descriptor (CUtensorMap) as an opaque blob, so that the device part also compiles on the host emulator
struct alignas(64) HvTmap { unsigned long long q[16]; };
struct PyrBuildList {
    const HvPyrDesc* table;   // device-resident descriptors of all pyramids of the context
    int n;                    // images in this launch
    unsigned short idx[PYR_MAX_BATCH];   // descriptor index per blockIdx.z
    // Frame already in HBM (e.g. decoded on the device): read level 0 from src[z] (row pitch srcPitch[z] bytes)
    // and let each CTA also write its own 64x64 tile into the level-0 buffer. NULL: the frame was copied (H2D)
    // straight into the level-0 buffer, which is then read in place.
    const uint8_t* src[PYR_MAX_BATCH];
    int srcPitch[PYR_MAX_BATCH];
    // TMA staging of the level-0 region (cp.async.bulk.tensor.2d + mbarrier): one descriptor per image over its level-0 source (u8,
    // w x h, row pitch a multiple of 16 bytes), box = (pitch of the level-0 shared-memory buffer) x (tile + 2 halo rows). useTma[z] == 0:
    // the source does not meet TMA's alignment rules (or HV_PYR_NO_TMA=1): staged with 32-bit loads instead.
    unsigned char useTma[PYR_MAX_BATCH];
    HvTmap tmap[PYR_MAX_BATCH];
};

#ifndef HV_EMU
__device__ __forceinline__ unsigned pyr_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pyr_mbar_init(unsigned long long* bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(pyr_smem_u32(bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // the init must be visible to the async (TMA) proxy
}
__device__ __forceinline__ void pyr_tma_load_2d(void* dst, const HvTmap* map, int x, int y, unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pyr_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(pyr_smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(pyr_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pyr_mbar_wait(unsigned long long* bar, unsigned phase)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(pyr_smem_u32(bar)), "r"(phase) : "memory");
    } while (!done);
}
#endif

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
__device__ __forceinline__ void emit_item(const HvLevel& L, const uint8_t* buf, int bp, int bx0, int by0,
                                          const Span& sx, const Span& sy, bool writeGray, int itemW, int ix, int iy)
{
    const int w = L.w, h = L.h;
    {
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

__device__ __forceinline__ void emit_level(const HvLevel& L, const uint8_t* buf, int bp, int bx0, int by0,
                                           const Span& sx, const Span& sy, bool writeGray, int itemW)
{
    const int tw = sx.o1 - sx.o0, th = sy.o1 - sy.o0;
    const int itemsX = (tw + itemW - 1) / itemW;
    const int nItems = itemsX * th;
    for (int it = threadIdx.x; it < nItems; it += PYR_NT) {
        const int iy = it / itemsX, ix = it - iy * itemsX;
        emit_item(L, buf, bp, bx0, by0, sx, sy, writeGray, itemW, ix, iy);
    }
}

// ------------------------------------------------------------------------------------------------ second generation
#ifdef HV_EMU
#define PYR_ALIGNED(p, n) do { if (((uintptr_t)(p)) % (n)) { fprintf(stderr, "misaligned %d-byte access\n", (int)(n)); abort(); } } while (0)
#else
#define PYR_ALIGNED(p, n) do { } while (0)
#endif

// BORDER_REFLECT_101 for -len < p < 2 len - 1 (one reflection), len >= 2
__device__ __forceinline__ int pyr_reflect1(int p, int len) { p = abs(p); return min(p, 2 * len - 2 - p); }

// Scharr of a strip of the owned tile: columns gx .. gx + 3 (all owned; gx - bx0 is a multiple of 4), rows y0 .. y0 + RB - 1
// (those below sy.o1 are skipped). Per register two 16-bit lanes = two neighbouring columns:
//   A = (gx - 1, gx), B = (gx + 1, gx + 2), C = (gx + 3, gx + 4);   t0 = 3 (above + below) + 10 centre  (<= 4080),
//   t1 = below - above + 2048  (in [1793, 2303]);   Ix = t0[i + 2] - t0[i] + 0x8000,   Iy = 3 (t1[i] + t1[i + 2]) + 10 t1[i + 1]
//   = Iy_true + 16 * 2048 = Iy_true + 0x8000: both in [28688, 36848], and x + 0x8000 mod 2^16 = x ^ 0x8000.
// PLAIN: all RB + 2 input rows lie inside the image and all RB output rows are owned -- no reflection, no row test, the row
// addresses are compile-time multiples of the pitch off one base.
template <int RB, bool PLAIN>
__device__ __forceinline__ void emit_strip(const HvLevel& L, const uint8_t* buf, int bp, int bx0, int by0, const Span& sx, const Span& sy,
                                           bool writeGray, int ix, int iy)
{
    const int w = L.w, h = L.h;
    const int gx = sx.o0 + 4 * ix, y0 = sy.o0 + RB * iy;
    const int xc = gx - bx0, xl = pyr_reflect1(gx - 1, w) - bx0, xr = pyr_reflect1(gx + 4, w) - bx0;
    uint32_t a[3], b[3], d[3], wb = 0, wd = 0;
    const uint8_t* base = buf + (y0 - 1 - by0) * bp;                   // row y0 - 1 (PLAIN)
    auto load = [&](int y, uint32_t (&v)[3], uint32_t& word) {
        const uint8_t* row = PLAIN ? base + (y - (y0 - 1)) * bp : buf + (pyr_reflect1(y, h) - by0) * bp;
        PYR_ALIGNED(row + xc, 4);
        word = *reinterpret_cast<const uint32_t*>(row + xc);
        const uint32_t l = row[xl], r = row[xr];
        v[0] = __byte_perm(word, l, 0x5054);      // (l, c0)
        v[1] = __byte_perm(word, 0u, 0x4241);     // (c1, c2)
        v[2] = __byte_perm(word, r, 0x5453);      // (c3, r)
    };
    load(y0 - 1, a, wd);
    load(y0, b, wb);
    short2* drow = L.deriv + (size_t)y0 * L.dpitch + gx;
    uint8_t* grow = L.gray + (size_t)y0 * L.gpitch + gx;
    const int dpitch = L.dpitch, gpitch = L.gpitch;
    PYR_ALIGNED(drow, 16); PYR_ALIGNED(L.deriv + dpitch, 16);
    PYR_ALIGNED(grow, 4); PYR_ALIGNED(L.gray + gpitch, 4);
#pragma unroll
    for (int r = 0; r < RB; r++) {
        const int y = y0 + r;
        if (!PLAIN && y >= sy.o1) break;
        load(y + 1, d, wd);
        uint32_t t0[3], t1[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            t0[i] = (a[i] + d[i]) * 3u + b[i] * 10u;
            t1[i] = d[i] + 0x08000800u - a[i];
        }
        const uint32_t gx01 = (t0[1] + 0x80008000u - t0[0]) ^ 0x80008000u;
        const uint32_t gx23 = (t0[2] + 0x80008000u - t0[1]) ^ 0x80008000u;
        const uint32_t m01 = __byte_perm(t1[0], t1[1], 0x5432), m23 = __byte_perm(t1[1], t1[2], 0x5432);
        const uint32_t gy01 = ((t1[0] + t1[1]) * 3u + m01 * 10u) ^ 0x80008000u;
        const uint32_t gy23 = ((t1[1] + t1[2]) * 3u + m23 * 10u) ^ 0x80008000u;
        int4 v;
        v.x = (int)__byte_perm(gx01, gy01, 0x5410); v.y = (int)__byte_perm(gx01, gy01, 0x7632);
        v.z = (int)__byte_perm(gx23, gy23, 0x5410); v.w = (int)__byte_perm(gx23, gy23, 0x7632);
        *reinterpret_cast<int4*>(drow) = v;
        if (writeGray) *reinterpret_cast<uint32_t*>(grow) = wb;
        drow += dpitch; grow += gpitch;
#pragma unroll
        for (int i = 0; i < 3; i++) { a[i] = b[i]; b[i] = d[i]; }
        wb = wd;
    }
}

// pyrDown (5x5 [1 4 6 4 1]^2, (sum + 128) >> 8) of RS vertically adjacent outputs (cx, cy0 .. cy0 + RS - 1), rows >= cyEnd skipped:
// the horizontal sums of the 2 RS + 3 source rows are formed once and shared. A vertical sum is at most 65,280.
template <int RS>
__device__ __forceinline__ void pyrdown_strip(const uint8_t* src, int sp, int sx0, int sy0, int sw, int sh, uint8_t* dst, int dp, int dx0, int dy0,
                                              int cx, int cy0, int cyEnd)
{
    int xs[5];
#pragma unroll
    for (int i = 0; i < 5; i++) xs[i] = pyr_reflect1(2 * cx - 2 + i, sw) - sx0;
    int hs[2 * RS + 3];
#pragma unroll
    for (int j = 0; j < 2 * RS + 3; j++) {
        const int yy = min(2 * cy0 - 2 + j, 2 * cyEnd);       // rows past the last output's window: clamped, results unused
        const uint8_t* row = src + (pyr_reflect1(yy, sh) - sy0) * sp;
        hs[j] = row[xs[2]] * 6 + (row[xs[1]] + row[xs[3]]) * 4 + row[xs[0]] + row[xs[4]];
    }
    uint8_t* out = dst + (cy0 - dy0) * dp + (cx - dx0);
#pragma unroll
    for (int r = 0; r < RS; r++)
        if (cy0 + r < cyEnd)
            out[r * dp] = (uint8_t)((hs[2 * r] + hs[2 * r + 4] + (hs[2 * r + 1] + hs[2 * r + 3]) * 4 + hs[2 * r + 2] * 6 + 128) >> 8);
}

// pyrDown of the stored region of level k, one output pixel per work item (first generation; any level size)
__device__ __forceinline__ void pyrdown_generic(const uint8_t* src, int sp, int sx0, int sy0, int sw, int sh, uint8_t* dst, int dp, int dx0,
                                                const Span& dsx, const Span& dsy)
{
    const int cw = dsx.s1 - dsx.s0 + 1, ch = dsy.s1 - dsy.s0 + 1;
    for (int it = threadIdx.x; it < cw * ch; it += PYR_NT) {
        const int iy = it / cw, ix = it - iy * cw;
        const int cx = dsx.s0 + ix, cy = dsy.s0 + iy;
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
        dst[iy * dp + cx - dx0] = (uint8_t)((acc + 128) >> 8);
    }
}

// The same for TWO neighbouring outputs (cx odd, cx + 1) whose 7 source columns 2 cx - 2 .. 2 cx + 4 need no reflection: they are the
// first 7 bytes of two aligned 32-bit words q0..q7 (the buffer origin and 2 cx - 2 are multiples of 4), and the horizontal pass runs on
// two 16-bit lanes (lo = cx, hi = cx + 1): 1 q0 + 4 q1 + 6 q2 + 4 q3 + 1 q4 | 1 q2 + 4 q3 + 6 q4 + 4 q5 + 1 q6 (<= 4080 per lane), the
// vertical pass as well (<= 65,408 with the rounding term: no carry between the lanes).
template <int RS, bool PLAIN>
__device__ __forceinline__ void pyrdown_pair(const uint8_t* src, int sp, int sx0, int sy0, int sh, uint8_t* dst, int dp, int dx0, int dy0,
                                             int cx, int cy0, int cyEnd)
{
    const int xo = 2 * cx - 2 - sx0;
    const uint8_t* base = src + (2 * cy0 - 2 - sy0) * sp + xo;        // PLAIN: all 2 RS + 3 source rows inside the image, all RS outputs wanted
    uint32_t hs[2 * RS + 3];
#pragma unroll
    for (int j = 0; j < 2 * RS + 3; j++) {
        const int yy = min(2 * cy0 - 2 + j, 2 * cyEnd);
        const uint8_t* row = PLAIN ? base + j * sp : src + (pyr_reflect1(yy, sh) - sy0) * sp + xo;
        PYR_ALIGNED(row, 4);
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(row), w1 = *reinterpret_cast<const uint32_t*>(row + 4);
        const uint32_t e0 = __byte_perm(w0, 0u, 0x4240), o0 = __byte_perm(w0, 0u, 0x4341);     // (q0, q2), (q1, q3)
        const uint32_t e1 = __byte_perm(w1, 0u, 0x4240), o1 = __byte_perm(w1, 0u, 0x4341);     // (q4, q6), (q5, q7)
        const uint32_t t2 = __byte_perm(e0, e1, 0x5432), t3 = __byte_perm(o0, o1, 0x5432);     // (q2, q4), (q3, q5)
        hs[j] = e0 + e1 + (o0 + t3) * 4u + t2 * 6u;
    }
    uint8_t* out = dst + (cy0 - dy0) * dp + (cx - dx0);
#pragma unroll
    for (int r = 0; r < RS; r++)
        if (PLAIN || cy0 + r < cyEnd) {
            const uint32_t v = hs[2 * r] + hs[2 * r + 4] + (hs[2 * r + 1] + hs[2 * r + 3]) * 4u + hs[2 * r + 2] * 6u + 0x00800080u;
            out[r * dp] = (uint8_t)(v >> 8);
            out[r * dp + 1] = (uint8_t)(v >> 24);
        }
}

// pyrDown of the stored region of level k, second generation: work item = (pair of columns (2 p + 1, 2 p + 2), strip of RS rows)
template <int RS>
__device__ __forceinline__ void pyrdown_fast(const uint8_t* src, int sp, int sx0, int sy0, int sw, int sh, uint8_t* dst, int dp, int dx0,
                                             const Span& dsx, const Span& dsy)
{
    const int p0 = (dsx.s0 - 1) >> 1, np = ((dsx.s1 - 1) >> 1) - p0 + 1;
    const int ch = dsy.s1 - dsy.s0 + 1, ns = (ch + RS - 1) / RS;
    for (int it = threadIdx.x; it < np * ns; it += PYR_NT) {
        const int is = it / np, ip = it - is * np;
        const int ca = 2 * (p0 + ip) + 1, cy0 = dsy.s0 + RS * is;
        const bool va = ca >= dsx.s0, vb = ca + 1 <= dsx.s1;
        if (va && vb && ca >= 1 && 2 * ca + 4 <= sw - 1) {
            if (cy0 >= 1 && 2 * cy0 + 2 * RS <= sh - 1 && cy0 + RS <= dsy.s1 + 1) pyrdown_pair<RS, true>(src, sp, sx0, sy0, sh, dst, dp, dx0, dsy.s0, ca, cy0, dsy.s1 + 1);
            else pyrdown_pair<RS, false>(src, sp, sx0, sy0, sh, dst, dp, dx0, dsy.s0, ca, cy0, dsy.s1 + 1);
        } else {
            if (va) pyrdown_strip<RS>(src, sp, sx0, sy0, sw, sh, dst, dp, dx0, dsy.s0, ca, cy0, dsy.s1 + 1);
            if (vb) pyrdown_strip<RS>(src, sp, sx0, sy0, sw, sh, dst, dp, dx0, dsy.s0, ca + 1, cy0, dsy.s1 + 1);
        }
    }
}

// Scharr pass of the second generation: full 4-pixel items as strips of RB rows, a partial item column (level width not a multiple
// of 4 at the right image border) through the first-generation item
template <int RB>
__device__ __forceinline__ void emit_fast(const HvLevel& L, const uint8_t* buf, int bp, int bx0, int by0, const Span& sx, const Span& sy, bool writeGray)
{
    const int tw = sx.o1 - sx.o0, th = sy.o1 - sy.o0;
    const int full = tw >> 2, ns = (th + RB - 1) / RB;
    for (int it = threadIdx.x; it < full * ns; it += PYR_NT) {
        const int iy = it / full, ix = it - iy * full;
        const int y0 = sy.o0 + RB * iy;
        if (y0 >= 1 && y0 + RB <= L.h - 1 && y0 + RB <= sy.o1) emit_strip<RB, true>(L, buf, bp, bx0, by0, sx, sy, writeGray, ix, iy);
        else emit_strip<RB, false>(L, buf, bp, bx0, by0, sx, sy, writeGray, ix, iy);
    }
    if (tw & 3)
        for (int iy = threadIdx.x; iy < th; iy += PYR_NT) emit_item(L, buf, bp, bx0, by0, sx, sy, writeGray, 4, full, iy);
}

__device__ __forceinline__ void pyr_body(const PyrBuildList& list, uint8_t* smem)
{
    const HvPyrDesc& P = list.table[list.idx[blockIdx.z]];
    const int nl = P.nlevels, top = nl - 1;
    const int tx = blockIdx.x, ty = blockIdx.y;
    if (tx * HV_PYR_TILE >= P.lv[0].w || ty * HV_PYR_TILE >= P.lv[0].h) return;

#ifndef HV_EMU
    __shared__ __align__(8) unsigned long long s_bar;
    if (list.useTma[blockIdx.z] && threadIdx.x == 0) pyr_mbar_init(&s_bar);      // made visible to the others by the geometry barrier below
#endif
    // ---- geometry (identical in every thread)
    Span sx[HV_MAX_LEVELS], sy[HV_MAX_LEVELS];
    int bufOff[HV_MAX_LEVELS], bufPitch[HV_MAX_LEVELS];
    int ox[HV_MAX_LEVELS];      // level coordinate of column 0 of the level's shared-memory buffer
    {
        // ~450 instructions when every thread derives it for itself: a tenth of the first generation's work, but more than a third of
        // the second generation's. One thread per axis and one for the buffer layout, handed to the others through shared memory.
        __shared__ Span g_sx[HV_MAX_LEVELS], g_sy[HV_MAX_LEVELS];
        __shared__ int g_off[HV_MAX_LEVELS], g_pitch[HV_MAX_LEVELS];
        if (threadIdx.x == 0 || threadIdx.x == 32) {
            const bool isX = threadIdx.x == 0;
            const int t = isX ? tx : ty;
            Span* g = isX ? g_sx : g_sy;
            int a = 0, b = 0;
            for (int k = top; k >= 0; k--) {
                const int len = isX ? P.lv[k].w : P.lv[k].h;
                Span sp;
                sp.o0 = (t * HV_PYR_TILE) >> k; sp.o1 = min(len, ((t + 1) * HV_PYR_TILE) >> k);
                int na = sp.o0 - 1, nb = sp.o1;
                if (k < top) { na = min(na, 2 * a - 2); nb = max(nb, 2 * b + 2); }
                span_close(sp, na, nb, len);
                a = sp.s0; b = sp.s1;
                g[k] = sp;
            }
        } else if (threadIdx.x == 64) {
            int off = 0;
            for (int k = 0; k < nl; k++) {
                int rw = (HV_PYR_TILE >> k) + 2 * halo_of(k, top) + 4;
                rw = (rw + 3) & ~3;
                // level 0: the TMA box is as wide as the buffer (multiple of 16 bytes) and starts on a 16-byte boundary of the image row,
                // up to 12 bytes left of the 4-byte-aligned origin the staging by loads uses: 16 spare columns (+16: the row pitch stays
                // off a multiple of 128 bytes)
                const int pitch = k == 0 ? ((rw + 15) & ~15) + 32 : rw;
                g_off[k] = off; g_pitch[k] = pitch;
                off += pitch * rw; off = (off + 127) & ~127;
            }
        }
        __syncthreads();
        for (int k = 0; k < nl; k++) {
            sx[k] = g_sx[k]; sy[k] = g_sy[k]; bufOff[k] = g_off[k]; bufPitch[k] = g_pitch[k];
            // the owned tile starts on a 4-byte boundary of the buffer (32-bit shared-memory loads of the strips)
            ox[k] = sx[k].s0 - ((sx[k].s0 - sx[k].o0) & 3);
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
        int xorg = x0a;                               // image column of buffer column 0
#ifndef HV_EMU
        if (list.useTma[blockIdx.z]) {
            // ONE bulk tensor copy for the whole region: rows below / columns right of the image arrive as zeros (they are never read:
            // the spans are closed under reflection inside the image), the box is as wide as the buffer pitch. The box must start on a
            // 16-BYTE boundary of the row: with u8 elements any other inner coordinate traps as an illegal instruction on B200
            // (tools/tma_probe.cu, profiles/r02_tma_probe.md), so the region begins up to 12 columns further left.
            xorg = x0a & ~15;
            if (threadIdx.x == 0) pyr_tma_load_2d(b0, &list.tmap[blockIdx.z], xorg, sy[0].s0, &s_bar, (unsigned)(bp * (HV_PYR_TILE + 2 * halo_of(0, top))));
            pyr_mbar_wait(&s_bar, 0);
        } else
#endif
        if (aligned && (ext == nullptr || x0a + words * 4 <= pitch)) {
            // four rows per warp in flight: the loads of a staging pass are independent, but one load -> store pair per iteration
            // costs a full L2 / HBM round trip per row (13 per warp for the 108 rows of a 4-level tile)
            constexpr int NW = PYR_NT / 32;
            for (int c0 = 0; c0 < words; c0 += 32) {
                const int c = c0 + lane;
                for (int r = wrp; r < rows; r += 4 * NW) {
                    uint32_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int rr = r + u * NW;
                        v[u] = (rr < rows && c < words) ? __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)(sy[0].s0 + rr) * pitch + x0a) + c) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int rr = r + u * NW;
                        if (rr < rows && c < words) reinterpret_cast<uint32_t*>(b0 + rr * bp)[c] = v[u];
                    }
                }
            }
        } else {   // arbitrary external pitch / alignment: byte loads, never past the last image column
            const int nbytes = min(words * 4, L0.w - x0a);
            for (int r = wrp; r < rows; r += PYR_NT / 32) {
                const uint8_t* src = base + (size_t)(sy[0].s0 + r) * pitch + x0a;
                for (int c = lane; c < nbytes; c += 32) b0[r * bp + c] = __ldg(src + c);
            }
        }
        sx[0].s0 = xorg;  // stored origin is the aligned one (o0 is a multiple of the tile size, so this is ox[0] as well)
        ox[0] = xorg;
    }
#ifndef HV_EMU
    if (!list.useTma[blockIdx.z])
#endif
    __syncthreads();

    // ---- second generation: between two barriers, the Scharr pass of level k and the pyrDown pass that produces level k + 1
    // (both only read the buffer of level k)
    for (int k = 0; k < nl; k++) {
        const HvLevel L = P.lv[k];      // by value: the stores below go through generic pointers and would force re-loads of the table
        const uint8_t* buf = smem + bufOff[k];
        const int bp = bufPitch[k];
        const bool writeGray = k > 0 || ext != nullptr;
        if ((HV_PYR_TILE >> k) >= 4 && L.w >= 8 && L.h >= 8) {
            if (k == 0) emit_fast<4>(L, buf, bp, ox[k], sy[k].s0, sx[k], sy[k], writeGray);
            else if (k == 1) emit_fast<2>(L, buf, bp, ox[k], sy[k].s0, sx[k], sy[k], writeGray);
            else emit_fast<1>(L, buf, bp, ox[k], sy[k].s0, sx[k], sy[k], writeGray);
        } else {
            emit_level(L, buf, bp, ox[k], sy[k].s0, sx[k], sy[k], writeGray, min(4, HV_PYR_TILE >> k));
        }
        if (k + 1 == nl) break;
        uint8_t* dst = smem + bufOff[k + 1];
        if (L.w >= 8 && L.h >= 8) {
            if (k == 0) pyrdown_fast<4>(buf, bp, ox[k], sy[k].s0, L.w, L.h, dst, bufPitch[k + 1], ox[k + 1], sx[k + 1], sy[k + 1]);
            else pyrdown_fast<2>(buf, bp, ox[k], sy[k].s0, L.w, L.h, dst, bufPitch[k + 1], ox[k + 1], sx[k + 1], sy[k + 1]);
        } else {
            pyrdown_generic(buf, bp, ox[k], sy[k].s0, L.w, L.h, dst, bufPitch[k + 1], ox[k + 1], sx[k + 1], sy[k + 1]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(PYR_NT, 4) hv_pyr_fused2_kernel(const __grid_constant__ PyrBuildList list)
{
    extern __shared__ __align__(128) uint8_t smem[];
    pyr_body(list, smem);
}

// Host-side launch helper (called from capi.cu). Shared memory is sized for the deepest pyramid in the list.
size_t hv_pyr_smem_bytes(int nlevels)
{
    size_t off = 0;
    for (int k = 0; k < nlevels; k++) {
        int h = 1; for (int i = nlevels - 1; i > k; --i) h = 2 * h + 2;
        int rw = (HV_PYR_TILE >> k) + 2 * h + 4; rw = (rw + 3) & ~3;
        const int pitch = k == 0 ? ((rw + 15) & ~15) + 32 : rw;
        off += (size_t)pitch * rw; off = (off + 127) & ~(size_t)127;
    }
    return off;
}

// ---- TMA descriptors (host): cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
#include "hv_device_once.cuh"
#include <cuda.h>
typedef CUresult (*PyrEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PyrEncodeFn pyr_encode_fn()
{
    static PyrEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (getenv("HV_PYR_NO_TMA")) return (PyrEncodeFn) nullptr;          // A/B switch: stage with 32-bit loads instead
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (PyrEncodeFn)p;
    }();
    return fn;
}
// Level-0 source `base` (u8, w x h, `pitch` bytes per row) as a 2-D tensor with a (boxW x boxH) box; false if TMA's rules are not met
static bool pyr_make_tmap(HvTmap& out, const uint8_t* base, int w, int h, int pitch, int boxW, int boxH)
{
    static_assert(sizeof(CUtensorMap) == sizeof(HvTmap), "CUtensorMap is 128 bytes");
    PyrEncodeFn enc = pyr_encode_fn();
    if (!enc || (((size_t)base) & 15) || (pitch & 15) || boxW > 256 || boxH > 256 || (boxW & 15)) return false;
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH}, estr[2] = {1, 1};
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    memcpy(&out, &m, sizeof(m));
    return true;
}

cudaError_t hv_launch_pyr_fused(const HvPyrDesc* table, const unsigned short* idx, const uint8_t* const* src, const int* srcPitch,
                                const uint8_t* const* level0, const int* level0Pitch, const int* nlevels,
                                int n, int w0, int h0, int maxNlevels, cudaStream_t stream)
{
    static bool seen[64];                             // per device (hv_common.cuh)
    size_t smem = hv_pyr_smem_bytes(maxNlevels);
    if (hv_first_use_on_device(seen)) {
        cudaError_t e = cudaFuncSetAttribute(hv_pyr_fused2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
    }
    for (int base = 0; base < n; base += PYR_MAX_BATCH) {
        PyrBuildList list;
        list.table = table; list.n = min(PYR_MAX_BATCH, n - base);
        for (int i = 0; i < list.n; i++) {
            list.idx[i] = idx[base + i];
            list.src[i] = src ? src[base + i] : nullptr;
            list.srcPitch[i] = src ? srcPitch[base + i] : 0;
            // TMA descriptor over the image the CTAs stage from: the external frame if there is one, else the level-0 buffer
            const int nl = nlevels[base + i];
            int halo = 1; for (int q = nl - 1; q > 0; --q) halo = 2 * halo + 2;
            int rw = HV_PYR_TILE + 2 * halo + 4; rw = (rw + 3) & ~3; rw = ((rw + 15) & ~15) + 32;       // = pitch of the level-0 shared-memory buffer
            const uint8_t* img = list.src[i] ? list.src[i] : level0[base + i];
            const int pitch = list.src[i] ? list.srcPitch[i] : level0Pitch[base + i];
            list.useTma[i] = pyr_make_tmap(list.tmap[i], img, w0, h0, pitch, rw, HV_PYR_TILE + 2 * halo) ? 1 : 0;
        }
        dim3 grid((w0 + HV_PYR_TILE - 1) / HV_PYR_TILE, (h0 + HV_PYR_TILE - 1) / HV_PYR_TILE, list.n);
        hv_pyr_fused2_kernel<<<grid, PYR_NT, smem, stream>>>(list);
    }
    return cudaGetLastError();
}
