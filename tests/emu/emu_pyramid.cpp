// tests/emu/emu_pyramid.cpp -- the REAL body of hv_pyr_fused2_kernel (device part of hybvio_b200/csrc/pyramid.cu: all levels of an image in
// one launch, 64 x 64 level-0 tile per CTA, coarser levels from shared memory) on the host emulator against the C oracle
// (oracle/hv_oracle_lk.c: pyrDown + Scharr as OpenCV computes them): every level's gray and gradient images bit-identical.
// "pyr_device.inc" is cut out of pyramid.cu by the test that builds this file (the `extern __shared__` array becomes a pointer).
#include <algorithm>
#include "cuda_emu.h"
#include "pyr_device.inc"

extern "C" {
struct orc_pyramid;
orc_pyramid* orc_pyr_create(const uint8_t* img, int w, int h, int stride, int win, int maxLevel);
int orc_pyr_levels(const orc_pyramid* p);
void orc_pyr_level_size(const orc_pyramid* p, int level, int* w, int* h);
void orc_pyr_get_level_padded(const orc_pyramid* p, int level, uint8_t* gray, int16_t* deriv);
void orc_pyr_free(orc_pyramid* p);
}

static unsigned hash2(int x, int y) { unsigned h = (unsigned)x * 374761393u + (unsigned)y * 668265263u; h = (h ^ (h >> 13)) * 1274126177u; return h ^ (h >> 16); }

static int g_fill = 0x00;      // initial content of the shared-memory arena: a result that depends on bytes no thread wrote changes with it
static int g_win = 31;        // buildOpticalFlowPyramid stops at the first level that is not larger than the window
static int g_pattern = 0;     // 1: only 0 and 255 (largest gradients: the 16-bit lanes of the second generation at their limits)

static int run(int W, int H, int maxLevel, bool fromSrc, bool gen2)
{
    const int WIN = g_win;
    std::vector<uint8_t> img((size_t)W * H);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) img[(size_t)y * W + x] = g_pattern ? (uint8_t)((hash2(x / 2, y / 2) & 1) ? 255 : 0) : (uint8_t)((hash2(x / 3, y / 5) & 0x7f) + (hash2(x, y) & 0x7f));
    orc_pyramid* po = orc_pyr_create(img.data(), W, H, W, WIN, maxLevel);
    const int nl = orc_pyr_levels(po);
    HvPyrDesc desc; memset(&desc, 0, sizeof(desc));
    desc.nlevels = nl; desc.win = WIN;
    std::vector<uint8_t*> gbuf(nl); std::vector<short2*> dbuf(nl);
    for (int lv = 0; lv < nl; lv++) {
        int w, h; orc_pyr_level_size(po, lv, &w, &h);
        const int gp = lv == 0 && w % 4 == 0 ? w : (w + 127) & ~127, dp = (w + 31) & ~31;
        gbuf[lv] = (uint8_t*)aligned_alloc(128, (((size_t)gp * h + 256) + 127) & ~(size_t)127); memset(gbuf[lv], 0xAB, (size_t)gp * h);
        dbuf[lv] = (short2*)aligned_alloc(128, ((size_t)dp * h * 4 + 256 + 127) & ~(size_t)127); memset(dbuf[lv], 0xCD, (size_t)dp * h * 4);
        HvLevel& L = desc.lv[lv]; L.gray = gbuf[lv]; L.deriv = dbuf[lv]; L.w = w; L.h = h; L.gpitch = gp; L.dpitch = dp;
    }
    if (!fromSrc) for (int y = 0; y < H; y++) memcpy(gbuf[0] + (size_t)y * desc.lv[0].gpitch, img.data() + (size_t)y * W, W);       // the frame was copied into level 0
    PyrBuildList list; memset(&list, 0, sizeof(list));
    list.table = &desc; list.n = 1; list.idx[0] = 0; list.src[0] = fromSrc ? img.data() : nullptr; list.srcPitch[0] = fromSrc ? W : 0;
    std::vector<unsigned char> smem(hv_pyr_smem_bytes(nl) + 64, (unsigned char)g_fill);
    emu_dynamic_smem = (unsigned char*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    const int gx = (W + HV_PYR_TILE - 1) / HV_PYR_TILE, gy = (H + HV_PYR_TILE - 1) / HV_PYR_TILE;
    gridDim.x = gx; gridDim.y = gy; gridDim.z = 1;
    for (int ty = 0; ty < gy; ty++) for (int tx = 0; tx < gx; tx++) {
        emu::block_y = ty; emu::block_z = 0;
        emu::launch_cta(PYR_NT, (unsigned)tx, [&] { hv_pyr_fused2_kernel(list); });
    }
    emu::block_y = 0;
    long long badG = 0, badD = 0, pix = 0;
    for (int lv = 0; lv < nl; lv++) {
        int w, h; orc_pyr_level_size(po, lv, &w, &h);
        const int pw = w + 2 * WIN, ph = h + 2 * WIN;
        std::vector<uint8_t> g((size_t)pw * ph); std::vector<int16_t> dd((size_t)pw * ph * 2);
        orc_pyr_get_level_padded(po, lv, g.data(), dd.data());
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
            const size_t o = (size_t)(y + WIN) * pw + x + WIN;
            badG += gbuf[lv][(size_t)y * desc.lv[lv].gpitch + x] != g[o];
            const short2 d = dbuf[lv][(size_t)y * desc.lv[lv].dpitch + x];
            badD += d.x != dd[2 * o] || d.y != dd[2 * o + 1];
            pix++;
        }
    }
    printf("%s %dx%d, %d levels, frame %s: %lld pixels, %lld gray / %lld gradient differences  %s\n", gen2 ? "gen2" : "gen1", W, H, nl, fromSrc ? "read from a separate buffer" : "copied into level 0", pix, badG, badD,
           badG + badD == 0 ? "ok" : "FAIL");
    for (int lv = 0; lv < nl; lv++) { free(gbuf[lv]); free(dbuf[lv]); }
    orc_pyr_free(po);
    return badG + badD != 0;
}

int main()
{
    int fails = 0;
    {
        const bool gen2 = true;
        fails += run(320, 240, 3, false, gen2);
        fails += run(320, 240, 3, true, gen2);
        fails += run(188, 120, 2, false, gen2);          // widths that are not multiples of the tile, odd level sizes
        fails += run(150, 101, 3, true, gen2);
        fails += run(752, 480, 3, false, gen2);          // BASELINE config 2: 4 levels
    }
    // level widths that are not multiples of 4 (partial items), levels below 8 pixels (per-pixel
    // fallback inside the kernel), deeper pyramids (tiles of 4 and 2 pixels at levels 4 and 5), a single tile, TUM-VI's 512 x 512
    fails += run(203, 77, 3, false, true);
    fails += run(67, 66, 3, true, true);
    fails += run(40, 24, 3, false, true);
    fails += run(130, 70, 5, false, true);
    fails += run(512, 512, 3, false, true);
    g_win = 3;
    fails += run(130, 70, 5, false, true);
    fails += run(97, 45, 4, true, true);
    fails += run(400, 300, 5, false, true);             // 6 levels, 94-pixel halo at level 0: staged rows of more than 32 words
    g_win = 31;
    g_fill = 0xFF;                                      // the packed pyrDown reads one byte past its 7 taps: must not reach a result
    fails += run(752, 480, 3, false, true);
    fails += run(203, 77, 3, true, true);
    fails += run(130, 70, 5, false, true);
    g_fill = 0x00;
    g_pattern = 1;
    fails += run(320, 240, 3, false, true);
    fails += run(203, 77, 3, true, true);
    return fails;
}
