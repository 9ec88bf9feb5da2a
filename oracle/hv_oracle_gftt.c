/* oracle/hv_oracle_gftt.c -- TEST INFRASTRUCTURE (never on the product path): plain-C restatement of the reference's default corner
 * detector on CPU images, SURVEY.md 8(f) N2:
 *   FeatureDetectorImplementation::detect            src/tracker/feature_detector.cpp:619-640  (sort, the resize quirk, applyMinDistance)
 *     CpuCornerResponse::operator()                  src/tracker/feature_detector.cpp:281-310  -> cv::cornerMinEigenVal(img, blockSize, 3)
 *       cornerEigenValsVecs / calcMinEigenVal        OCV/imgproc/src/corner.cpp:238-320, 52-96 (OCV = 3rdparty/mobile-cv-suite/opencv/modules)
 *       cv::Sobel 8U -> 32F, scale folded into the SMOOTHING kernel   OCV/imgproc/src/deriv.cpp (getSobelKernels, "if( dx == 0 ) kx *= scale; else ky *= scale")
 *       row / column filters of size 3                OCV/imgproc/src/filter.simd.hpp (SymmRowSmallFilter / SymmColumnSmallFilter:
 *                                                     symmetric  S1*k0 + (S0 + S2)*k1,  anti-symmetric  S2 - S0)
 *       cv::boxFilter(cov, 3x3, normalize = false)    OCV/imgproc/src/box_filter.dispatch.cpp
 *     CollectMax::cpuImplementation                  src/tracker/feature_detector.cpp:393-417  (one key point per bs x bs block, GAIN 16)
 *   FeatureDetector::applyMinDistance                src/tracker/feature_detector_legacy.cpp
 *
 * Parity status: the arithmetic up to the covariance products follows the reference's operation order (fp32, no fused multiply-add:
 * compile with -ffp-contract=off). The 3 x 3 box sum is formed directly, (a + b) + c per row and over the rows, whereas OpenCV
 * keeps RUNNING sums along each row and down each column (s += new - old), whose rounding depends on everything to the left / above:
 * that order is not reproducible by a parallel kernel, and the reference itself changes it with the CPU (AVX / FMA dispatch in
 * corner.cpp, filter.simd.hpp). Parity of the response is therefore a float tolerance (|d| <= 1e-6 + 1e-5 |r|), pinned against the
 * compiled reference in tests/test_oracle_gftt.py; key point coordinates are integers and must agree except in blocks whose two best
 * responses are closer than that tolerance (listed by the test).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}

/* cv::cornerMinEigenVal(src 8UC1, blockSize, ksize 3, BORDER_DEFAULT = REFLECT_101) -> response (w x h float) */
void orc_gftt_response(const uint8_t* img, int stride, int w, int h, int block_size, float* response)
{
    const double scale_d = 1.0 / ((double)(1 << 2) * block_size * 255.0);         /* corner.cpp:246-251 */
    const float k1 = (float)(1.0 * scale_d), k0 = (float)(2.0 * scale_d);         /* [1 2 1] * scale as a CV_32F kernel */
    float* dx = (float*)malloc(sizeof(float) * (size_t)w * h);
    float* dy = (float*)malloc(sizeof(float) * (size_t)w * h);
    float* cov = (float*)malloc(sizeof(float) * 3 * (size_t)w * h);
    /* Sobel dx: row pass [-1 0 1] (exact), column pass [1 2 1] * scale;  Sobel dy: row pass [1 2 1] * scale, column pass [-1 0 1] */
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = img + (size_t)reflect101(y - 1, h) * stride;
        const uint8_t* r1 = img + (size_t)y * stride;
        const uint8_t* r2 = img + (size_t)reflect101(y + 1, h) * stride;
        for (int x = 0; x < w; x++) {
            const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
            const float d0 = (float)(r0[xr] - r0[xl]), d1 = (float)(r1[xr] - r1[xl]), d2 = (float)(r2[xr] - r2[xl]);
            dx[(size_t)y * w + x] = d1 * k0 + (d0 + d2) * k1;
            const float s0 = (float)r0[x] * k0 + (float)(r0[xl] + r0[xr]) * k1;
            const float s2 = (float)r2[x] * k0 + (float)(r2[xl] + r2[xr]) * k1;
            dy[(size_t)y * w + x] = s2 - s0;
        }
    }
    for (size_t i = 0; i < (size_t)w * h; i++) { cov[3 * i] = dx[i] * dx[i]; cov[3 * i + 1] = dx[i] * dy[i]; cov[3 * i + 2] = dy[i] * dy[i]; }
    const int r = block_size / 2;                                                /* anchor (-1,-1): centre; block sizes are odd here */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float acc[3] = {0.f, 0.f, 0.f};
            for (int c = 0; c < 3; c++) {
                float tot = 0.f;
                for (int j = -r; j <= r; j++) {
                    const float* row = cov + 3 * (size_t)reflect101(y + j, h) * w;
                    float s = 0.f;
                    for (int i = -r; i <= r; i++) s = (i == -r) ? row[3 * reflect101(x + i, w) + c] : s + row[3 * reflect101(x + i, w) + c];
                    tot = (j == -r) ? s : tot + s;
                }
                acc[c] = tot;
            }
            const float a = acc[0] * 0.5f, b = acc[1], c2 = acc[2] * 0.5f;       /* corner.cpp:88-94 */
            const float t = a - c2;
            response[(size_t)y * w + x] = (a + c2) - sqrtf(b * b + t * t);
        }
    free(dx); free(dy); free(cov);
}

/* CollectMax::cpuImplementation (feature_detector.cpp:393-417): kp = (x, y, response) per block, row-major block order.
 * Returns the number of key points = (w / bs) * (h / bs) (integer division, as the reference's std::ceil of an int quotient). */
int orc_gftt_collect(const float* response, int w, int h, int bs, float min_response, float* kp_xyr)
{
    int n = 0;
    for (int yb = 0; yb < h / bs; yb++)
        for (int xb = 0; xb < w / bs; xb++) {
            float best = -1e10f; int bx = 0, by = 0;
            for (int y = yb * bs; y < (yb + 1) * bs && y < h; y++)
                for (int x = xb * bs; x < (xb + 1) * bs && x < w; x++) {
                    const float r = response[(size_t)y * w + x] * 16.0f;          /* CpuCornerResponse::GAIN */
                    if (r > best && r > min_response) { bx = x; by = y; best = r; }
                }
            kp_xyr[3 * n] = (float)bx; kp_xyr[3 * n + 1] = (float)by; kp_xyr[3 * n + 2] = best; n++;
        }
    return n;
}

/* FeatureDetectorImplementation::detect (feature_detector.cpp:625-638) after collectMax: stable sort by response (descending),
 * `corners.clear(); corners.resize(n); push_back...` (n zero points in front: a quirk of the reference that is kept), then
 * applyMinDistance(corners, prev, mask_radius) when mask_radius > 0. corners_xy: capacity 2 * nkp points. Returns the count. */
int orc_gftt_corners(const float* kp_xyr, int nkp, const float* prev_xy, int nprev, int mask_radius, int max_tracks, float* corners_xy)
{
    int* order = (int*)malloc(sizeof(int) * (size_t)(nkp > 0 ? nkp : 1));
    for (int i = 0; i < nkp; i++) order[i] = i;
    for (int i = 1; i < nkp; i++) {                                              /* insertion sort = stable */
        const int v = order[i]; int j = i - 1;
        while (j >= 0 && kp_xyr[3 * order[j] + 2] < kp_xyr[3 * v + 2]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = v;
    }
    int n = 0;
    for (int i = 0; i < nkp; i++) { corners_xy[2 * n] = 0.f; corners_xy[2 * n + 1] = 0.f; n++; }
    for (int i = 0; i < nkp; i++) { corners_xy[2 * n] = kp_xyr[3 * order[i]]; corners_xy[2 * n + 1] = kp_xyr[3 * order[i] + 1]; n++; }
    free(order);
    if (mask_radius <= 0) return n;
    const float r2 = (float)(mask_radius * mask_radius);
    int out = 0;
    for (int i = 0; i < n; i++) {
        const float cx = corners_xy[2 * i], cy = corners_xy[2 * i + 1];
        int near = 0;
        for (int k = 0; k < nprev && !near; k++) { const float ddx = prev_xy[2 * k] - cx, ddy = prev_xy[2 * k + 1] - cy; if (ddx * ddx + ddy * ddy < r2) near = 1; }
        for (int k = 0; k < out && !near; k++) { const float ddx = corners_xy[2 * k] - cx, ddy = corners_xy[2 * k + 1] - cy; if (ddx * ddx + ddy * ddy < r2) near = 1; }
        if (!near) { corners_xy[2 * out] = cx; corners_xy[2 * out + 1] = cy; out++; }
        if (out >= max_tracks) break;
    }
    return out;
}

/* ---- frame ingest, same test-infrastructure status as above -------------------------------------------------------
 * orc_gray: accelerated-arrays pixelwiseAffineUnary<FixedPoint<uint8_t>> (AA/cpu/operations.cpp:145-177, AA/fixed_point.hpp:16-36) as
 *           src/tracker/image.cpp:360-366 uses it: v = 0; v += coeff[j] * float(in[j]) in fp32 (channel order), out = T(v).
 * orc_remap: UndistorterImplementation::undistort, CPU branch (src/tracker/undistorter.cpp:84-112), with the camera mapping given as a
 *           table {x0, y0, xfrac, yfrac} per output pixel (x0 = -32768: no source pixel); taps beyond the last column / row read the
 *           linear address like cv::Mat::at does, clamped to the last byte of the image. */
typedef struct { int16_t x0, y0; float xfrac, yfrac; } orc_remap_entry;

void orc_gray(const uint8_t* src, int stride, int channels, int w, int h, const float* coeff, uint8_t* dst)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p = src + (size_t)y * stride + (size_t)x * channels;
            float v = 0.0f;
            for (int j = 0; j < channels; j++) { const float in = (float)((double)p[j] / 255.0); v += coeff[j] * in; }
            double d = (double)v;
            d = d < 0.0 ? 0.0 : d > 1.0 ? 1.0 : d;
            dst[(size_t)y * w + x] = (uint8_t)(255.0 * d + 0.5);
        }
}

void orc_remap(const uint8_t* src, int stride, int w, int h, const orc_remap_entry* table, uint8_t* dst)
{
    const long long last = (long long)(h - 1) * stride + (w - 1);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const orc_remap_entry e = table[(size_t)y * w + x];
            float out = 0.0f;
            if (e.x0 != -32768)
                for (int iy = 0; iy < 2; iy++) {
                    const float wy = iy > 0 ? e.yfrac : (1 - e.yfrac);
                    for (int ix = 0; ix < 2; ix++) {
                        const float wx = ix > 0 ? e.xfrac : (1 - e.xfrac);
                        long long a = (long long)(e.y0 + iy) * stride + (e.x0 + ix);
                        if (a > last) a = last;
                        out += src[a] * wx * wy;
                    }
                }
            dst[(size_t)y * w + x] = (uint8_t)(int)(out + 0.5);
        }
}
