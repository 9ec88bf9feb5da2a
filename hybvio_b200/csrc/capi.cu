// hybvio_b200/csrc/capi.cu -- C ABI (include/hybvio_b200.h): context, image pyramid, Lucas-Kanade.
// The EKF entry points live in ekf_capi.cu.
#include "capi_internal.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>

static thread_local char g_err[512] = "";

void hv_set_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

extern "C" {

const char* hv_version(void) { return "hybvio_b200 0.1 (sm_100a)"; }
const char* hv_last_error(void) { return g_err; }

int hv_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int ctx_create(int device, cudaStream_t stream, bool own, hv_ctx** out)
{
    if (!out) { hv_set_error("hv_ctx_create: out is NULL"); return HV_ERR_INVALID; }
    *out = nullptr;
    int n = hv_device_count();
    if (n <= 0 || device < 0 || device >= n) {
        hv_set_error("hv_ctx_create: no CUDA device %d (found %d). hybvio_b200 has no CPU fallback.", device, n);
        return HV_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    HV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        hv_set_error("hv_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return HV_ERR_NO_DEVICE;
    }
    HV_CUDA(cudaSetDevice(device));
    hv_ctx* c = new hv_ctx;
    c->device = device;
    // every failure below releases what has been created so far (hv_ctx_destroy tolerates a partly initialised context)
    auto fail = [&](cudaError_t e, const char* what) {
        hv_set_error("hv_ctx_create: %s failed: %s", what, cudaGetErrorString(e));
        hv_ctx_destroy(c);
        return e == cudaErrorMemoryAllocation ? HV_ERR_OOM : HV_ERR_CUDA;
    };
    cudaError_t e = cudaSuccess;
    if (own) { e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking); if (e != cudaSuccess) { c->stream = nullptr; return fail(e, "cudaStreamCreate"); } c->ownStream = true; }
    else c->stream = stream;
    if ((e = cudaMalloc(&c->d_table, sizeof(HvPyrDesc) * HV_TABLE_CAPACITY)) != cudaSuccess) { c->d_table = nullptr; return fail(e, "cudaMalloc"); }
    if ((e = cudaMemsetAsync(c->d_table, 0, sizeof(HvPyrDesc) * HV_TABLE_CAPACITY, c->stream)) != cudaSuccess) return fail(e, "cudaMemsetAsync");
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "cudaStreamSynchronize");
    for (int i = HV_TABLE_CAPACITY - 1; i >= 0; --i) c->freeSlots.push_back(i);
    *out = c;
    return HV_OK;
}

int hv_ctx_create(int device, hv_ctx** out) { return ctx_create(device, nullptr, true, out); }
int hv_ctx_create_on_stream(int device, void* s, hv_ctx** out) { return ctx_create(device, (cudaStream_t)s, false, out); }

int hv_ctx_destroy(hv_ctx* c)
{
    if (!c) return HV_OK;
    cudaSetDevice(c->device);
    if (c->stream || !c->ownStream) cudaStreamSynchronize(c->stream);
    if (c->sideStream) { cudaStreamSynchronize(c->sideStream); cudaStreamDestroy(c->sideStream); }
    if (c->covStream) { cudaStreamSynchronize(c->covStream); cudaStreamDestroy(c->covStream); }
    if (c->d_table) cudaFree(c->d_table);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_done) cudaFree(c->d_done);
    if (c->d_ekfStage) cudaFree(c->d_ekfStage);
    if (c->h_ekfStage) cudaFreeHost(c->h_ekfStage);
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return HV_OK;
}

int hv_ctx_sync(hv_ctx* c)
{
    if (!c) { hv_set_error("hv_ctx_sync: NULL ctx"); return HV_ERR_INVALID; }
    HV_CUDA(cudaStreamSynchronize(c->stream));
    if (c->sideStream) HV_CUDA(cudaStreamSynchronize(c->sideStream));
    if (c->covStream) HV_CUDA(cudaStreamSynchronize(c->covStream));
    return HV_OK;
}
void* hv_ctx_stream(hv_ctx* c) { return c ? (void*)c->stream : nullptr; }
long long hv_ctx_launch_count(hv_ctx* c) { return c ? c->launches : 0; }

} // extern "C"

int hv_ctx_reserve_stage(hv_ctx* c, size_t bytes)
{
    if (bytes <= c->stageBytes) return HV_OK;
    HV_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->d_stage = nullptr; c->h_stage = nullptr; c->stageBytes = 0;
    size_t cap = 4096; while (cap < bytes) cap *= 2;
    HV_CUDA(cudaMalloc(&c->d_stage, cap));
    HV_CUDA(cudaHostAlloc(&c->h_stage, cap + 64, cudaHostAllocMapped));          // + 64: the completion flag
    HV_CUDA(cudaHostGetDevicePointer(&c->hd_stage, c->h_stage, 0));
    memset(c->h_stage, 0, cap + 64);
    // (stream-ordered: the context's stream is non-blocking, a cudaMemset on the legacy stream could land after the first kernel's counts)
    if (!c->d_done) { HV_CUDA(cudaMalloc(&c->d_done, sizeof(unsigned))); HV_CUDA(cudaMemsetAsync(c->d_done, 0, sizeof(unsigned), c->stream)); c->doneCount = 0; }
    c->stageBytes = cap;
    return HV_OK;
}

// HV_NO_POLL=1: results come back with a D2H copy + stream synchronisation instead (A/B switch)
bool hv_polling_enabled() { static const bool on = getenv("HV_NO_POLL") == nullptr; return on; }

int hv_poll_flag(volatile unsigned* flag, unsigned seq, cudaStream_t stream, const char* who)
{
    for (unsigned long long spins = 1;; spins++) {
        if (*flag == seq) break;
        if ((spins & 0xfff) == 0) {
            const cudaError_t q = cudaStreamQuery(stream);
            if (q == cudaErrorNotReady) continue;
            if (q != cudaSuccess) { hv_set_error("%s: %s while waiting for the result", who, cudaGetErrorString(q)); return HV_ERR_CUDA; }
            if (*flag == seq) break;
            hv_set_error("%s: the kernel finished without raising its completion flag", who); return HV_ERR_STATE;
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    return HV_OK;
}

// ------------------------------------------------------------------------------------------------ pyramid
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" {

int hv_pyr_create(hv_ctx* c, int w, int h, int win, int maxLevel, hv_pyr** out)
{
    if (!c || !out || w <= 0 || h <= 0 || win <= 2 || maxLevel < 0) {
        hv_set_error("hv_pyr_create: invalid argument (w=%d h=%d win=%d maxLevel=%d)", w, h, win, maxLevel);
        return HV_ERR_INVALID;
    }
    if (maxLevel > HV_MAX_LEVELS - 1) {
        hv_set_error("hv_pyr_create: maxLevel %d > %d unsupported", maxLevel, HV_MAX_LEVELS - 1);
        return HV_ERR_UNSUPPORTED;
    }
    if (c->freeSlots.empty()) { hv_set_error("hv_pyr_create: more than %d live pyramids", HV_TABLE_CAPACITY); return HV_ERR_OOM; }
    HV_CUDA(cudaSetDevice(c->device));
    hv_pyr* p = new hv_pyr;
    p->ctx = c; p->w = w; p->h = h; p->win = win;
    // level geometry (OCV/video/src/lkpyramid.cpp:776-816)
    int lw = w, lh = h, nl = 0;
    size_t off = 0, goff[HV_MAX_LEVELS], doff[HV_MAX_LEVELS];
    memset(&p->desc, 0, sizeof(p->desc));
    for (int level = 0; level <= maxLevel; ++level) {
        HvLevel& L = p->desc.lv[level];
        L.w = lw; L.h = lh;
        // Level 0 IS the input image: with a dense pitch (w % 4 == 0 keeps the kernels' 32-bit loads aligned) a contiguous
        // host frame is ONE 1-D H2D copy; a padded pitch would make it a 2-D copy of h rows, measured at ~5x the time of
        // the 1-D copy of the same 361 KB (B200, PCIe gen5). Coarser levels are only ever written by the kernel.
        L.gpitch = (level == 0 && lw % 4 == 0) ? lw : (int)align_up((size_t)lw, 128);
        L.dpitch = (int)align_up((size_t)lw, 32);
        goff[level] = off; off = align_up(off + (size_t)L.gpitch * lh, 256);
        doff[level] = off; off = align_up(off + (size_t)L.dpitch * lh * sizeof(short2), 256);
        nl = level + 1;
        lw = (lw + 1) / 2; lh = (lh + 1) / 2;
        if (lw <= win || lh <= win) break;
    }
    p->nlevels = nl; p->desc.nlevels = nl; p->desc.win = win;
    p->bytes = off;
    cudaError_t e = cudaMalloc(&p->d_mem, off);
    if (e != cudaSuccess) { delete p; hv_set_error("hv_pyr_create: cudaMalloc(%zu) failed: %s", off, cudaGetErrorString(e)); return HV_ERR_OOM; }
    HV_CUDA(cudaMemsetAsync(p->d_mem, 0, off, c->stream));
    for (int level = 0; level < nl; ++level) {
        p->desc.lv[level].gray = (uint8_t*)p->d_mem + goff[level];
        p->desc.lv[level].deriv = (short2*)((uint8_t*)p->d_mem + doff[level]);
    }
    p->slot = c->freeSlots.back(); c->freeSlots.pop_back();
    HV_CUDA(cudaMemcpyAsync(c->d_table + p->slot, &p->desc, sizeof(HvPyrDesc), cudaMemcpyHostToDevice, c->stream));
    HV_CUDA(cudaStreamSynchronize(c->stream));   // desc is on the stack-owned object; creation is rare
    *out = p;
    return HV_OK;
}

int hv_pyr_release(hv_pyr* p)
{
    if (!p) return HV_OK;
    hv_ctx* c = p->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (p->d_mem) cudaFree(p->d_mem);
    c->freeSlots.push_back(p->slot);
    delete p;
    return HV_OK;
}

int hv_pyr_levels(const hv_pyr* p) { return p ? p->nlevels : HV_ERR_INVALID; }

int hv_pyr_level_size(const hv_pyr* p, int level, int* w, int* h)
{
    if (!p || level < 0 || level >= p->nlevels) { hv_set_error("hv_pyr_level_size: bad level"); return HV_ERR_INVALID; }
    if (w) *w = p->desc.lv[level].w;
    if (h) *h = p->desc.lv[level].h;
    return HV_OK;
}

int hv_pyr_build_batch(hv_pyr* const* pyrs, const uint8_t* const* gray, const size_t* strides, int n, int srcIsDevice)
{
    if (!pyrs || !gray || !strides || n <= 0) { hv_set_error("hv_pyr_build_batch: invalid argument"); return HV_ERR_INVALID; }
    hv_ctx* c = pyrs[0] ? pyrs[0]->ctx : nullptr;
    if (!c) { hv_set_error("hv_pyr_build_batch: NULL pyramid"); return HV_ERR_INVALID; }
    HV_CUDA(cudaSetDevice(c->device));
    std::vector<unsigned short> idx(n);
    std::vector<const uint8_t*> src(n);
    std::vector<int> srcPitch(n), l0Pitch(n), nls(n);
    std::vector<const uint8_t*> l0(n);
    int maxNl = 0;
    for (int i = 0; i < n; i++) {
        hv_pyr* p = pyrs[i];
        if (!p || p->ctx != c || p->w != pyrs[0]->w || p->h != pyrs[0]->h || !gray[i] || strides[i] < (size_t)p->w) {
            hv_set_error("hv_pyr_build_batch: pyramid %d invalid / different context or size", i);
            return HV_ERR_INVALID;
        }
        const HvLevel& L0 = p->desc.lv[0];
        if (srcIsDevice) {      // frame already in HBM: the kernel reads it in place and fills level 0 itself
            src[i] = gray[i]; srcPitch[i] = (int)strides[i];
        } else {                // the frame lands directly in the level-0 buffer: level 0 of the pyramid IS the input image
            if (strides[i] == (size_t)L0.gpitch)
                HV_CUDA(cudaMemcpyAsync(L0.gray, gray[i], (size_t)L0.gpitch * p->h, cudaMemcpyHostToDevice, c->stream));
            else
                HV_CUDA(cudaMemcpy2DAsync(L0.gray, L0.gpitch, gray[i], strides[i], (size_t)p->w, (size_t)p->h, cudaMemcpyHostToDevice, c->stream));
            src[i] = nullptr; srcPitch[i] = 0;
        }
        idx[i] = (unsigned short)p->slot;
        l0[i] = L0.gray; l0Pitch[i] = L0.gpitch; nls[i] = p->nlevels;
        if (p->nlevels > maxNl) maxNl = p->nlevels;
    }
    HV_CUDA(hv_launch_pyr_fused(c->d_table, idx.data(), src.data(), srcPitch.data(), l0.data(), l0Pitch.data(), nls.data(), n, pyrs[0]->w, pyrs[0]->h,
                                maxNl, c->stream));
    c->launches += (n + 31) / 32;
    return HV_OK;
}

int hv_pyr_build(hv_pyr* p, const uint8_t* gray, size_t stride)
{
    return hv_pyr_build_batch(&p, &gray, &stride, 1, 0);
}

int hv_pyr_download_level(hv_pyr* p, int level, uint8_t* gray, int16_t* deriv)
{
    if (!p || level < 0 || level >= p->nlevels) { hv_set_error("hv_pyr_download_level: bad level"); return HV_ERR_INVALID; }
    hv_ctx* c = p->ctx;
    HV_CUDA(cudaSetDevice(c->device));
    const HvLevel& L = p->desc.lv[level];
    if (gray) HV_CUDA(cudaMemcpy2DAsync(gray, L.w, L.gray, L.gpitch, L.w, L.h, cudaMemcpyDeviceToHost, c->stream));
    if (deriv) HV_CUDA(cudaMemcpy2DAsync(deriv, (size_t)L.w * 4, L.deriv, (size_t)L.dpitch * 4, (size_t)L.w * 4, L.h,
                                         cudaMemcpyDeviceToHost, c->stream));
    HV_CUDA(cudaStreamSynchronize(c->stream));
    return HV_OK;
}

int hv_pyr_download_level_padded(hv_pyr* p, int level, uint8_t* gray, int16_t* deriv)
{
    if (!p || level < 0 || level >= p->nlevels) { hv_set_error("hv_pyr_download_level_padded: bad level"); return HV_ERR_INVALID; }
    const HvLevel& L = p->desc.lv[level];
    const int win = p->win, W = L.w + 2 * win, H = L.h + 2 * win;
    std::vector<uint8_t> g((size_t)L.w * L.h);
    std::vector<int16_t> d((size_t)L.w * L.h * 2);
    int rc = hv_pyr_download_level(p, level, gray ? g.data() : nullptr, deriv ? d.data() : nullptr);
    if (rc != HV_OK) return rc;
    // border exactly as the reference materialises it: gray REFLECT_101, gradient CONSTANT 0 (lkpyramid.cpp:761-808)
    for (int y = 0; y < H; y++) {
        const int sy = hv_reflect101(y - win, L.h);
        const bool rowIn = (unsigned)(y - win) < (unsigned)L.h;
        for (int x = 0; x < W; x++) {
            const int sx = hv_reflect101(x - win, L.w);
            const bool in = rowIn && (unsigned)(x - win) < (unsigned)L.w;
            if (gray) gray[(size_t)y * W + x] = g[(size_t)sy * L.w + sx];
            if (deriv) {
                deriv[((size_t)y * W + x) * 2] = in ? d[((size_t)sy * L.w + sx) * 2] : 0;
                deriv[((size_t)y * W + x) * 2 + 1] = in ? d[((size_t)sy * L.w + sx) * 2 + 1] : 0;
            }
        }
    }
    return HV_OK;
}

// ------------------------------------------------------------------------------------------------ LK
static int lk_fill(LkLaunch& L, hv_ctx* c, int maxLevel, int maxIter, double eps, double minEig)
{
    // criteria clamp of SparsePyrLKOpticalFlowImpl::calc (lkpyramid.cpp:1361-1369)
    L.table = c->d_table;
    L.prefetch = 1;
    L.maxLevel = maxLevel;
    L.maxIter = maxIter < 0 ? 0 : (maxIter > 100 ? 100 : maxIter);
    double e = eps < 0. ? 0. : (eps > 10. ? 10. : eps);
    L.eps2 = e * e;
    L.minEig = (float)minEig;
    L.doneCounter = nullptr; L.doneTarget = 0; L.seq = 0; L.hostFlag = nullptr;
    return HV_OK;
}

static int lk_check_pair(const char* who, hv_ctx* c, hv_pyr* a, hv_pyr* b)
{
    if (!a || !b || a->ctx != c || b->ctx != c) { hv_set_error("%s: pyramid NULL or from another context", who); return HV_ERR_INVALID; }
    if (a->w != b->w || a->h != b->h || a->win != b->win) { hv_set_error("%s: pyramids differ in size/window", who); return HV_ERR_INVALID; }
    return HV_OK;
}

static int lk_track_batch_device_on(hv_ctx* c, cudaStream_t stream, const hv_lk_job* jobs, int njobs, int maxIter, double eps, double minEig, const float* initXY = nullptr)
{
    if (!c || !jobs || njobs < 0) { hv_set_error("hv_lk_track_batch_device: invalid argument"); return HV_ERR_INVALID; }
    HV_CUDA(cudaSetDevice(c->device));
    for (int base = 0; base < njobs; base += LK_MAX_JOBS) {
        LkLaunch L;
        int cnt = njobs - base < LK_MAX_JOBS ? njobs - base : LK_MAX_JOBS;
        int win = 0, maxLevel = HV_MAX_LEVELS;
        for (int i = 0; i < cnt; i++) {
            const hv_lk_job& j = jobs[base + i];
            int rc = lk_check_pair("hv_lk_track_batch_device", c, j.prev, j.next);
            if (rc != HV_OK) return rc;
            if (j.n < 0 || (j.n > 0 && (!j.d_prev_xy || !j.d_next_xy || !j.d_status))) {
                hv_set_error("hv_lk_track_batch_device: job %d has NULL buffers", base + i); return HV_ERR_INVALID;
            }
            if (win && win != j.prev->win) { hv_set_error("hv_lk_track_batch_device: mixed window sizes"); return HV_ERR_INVALID; }
            win = j.prev->win;
            LkJob& d = L.jobs[i];
            d.prevIdx = j.prev->slot; d.nextIdx = j.next->slot; d.n = j.n; d.useInitial = j.use_initial;
            d.prevPts = (const float2*)j.d_prev_xy; d.nextPts = (float2*)j.d_next_xy;
            d.status = j.d_status; d.trackStatus = j.d_track_status; d.initPts = initXY && cnt == 1 ? (const float2*)initXY : nullptr;
        }
        L.njobs = cnt;
        lk_fill(L, c, maxLevel, maxIter, eps, minEig);
        cudaError_t e = hv_launch_lk(L, win, stream);
        if (e == cudaErrorInvalidValue) { hv_set_error("hv_lk_track: window size %d unsupported (supported: 11, 15, 21, 31)", win); return HV_ERR_UNSUPPORTED; }
        HV_CUDA(e);
        c->launches += 1;
    }
    return HV_OK;
}
int hv_lk_track_batch_device(hv_ctx* c, const hv_lk_job* jobs, int njobs, int maxIter, double eps, double minEig)
{
    return lk_track_batch_device_on(c, c ? c->stream : nullptr, jobs, njobs, maxIter, eps, minEig);
}

static int lk_track_device_on(hv_ctx* c, cudaStream_t stream, hv_pyr* prev, hv_pyr* next, const float* dPrev, float* dNext, uint8_t* dStatus,
                              int32_t* dTs, int n, int useInitial, int maxIter, double eps, double minEig, const float* dInit = nullptr)
{
    if (!c || n < 0) { hv_set_error("hv_lk_track_device: invalid argument"); return HV_ERR_INVALID; }
    if (n == 0) {                 // optical_flow.cpp:41-44: empty input, empty output (the pyramids are still validated)
        int rc = lk_check_pair("hv_lk_track_device", c, prev, next);
        return rc;
    }
    hv_lk_job j; j.prev = prev; j.next = next; j.d_prev_xy = dPrev; j.d_next_xy = dNext; j.d_status = dStatus;
    j.d_track_status = dTs; j.n = n; j.use_initial = useInitial;
    return lk_track_batch_device_on(c, stream, &j, 1, maxIter, eps, minEig, dInit);
}
int hv_lk_track_device(hv_ctx* c, hv_pyr* prev, hv_pyr* next, const float* dPrev, float* dNext, uint8_t* dStatus,
                       int32_t* dTs, int n, int useInitial, int maxIter, double eps, double minEig)
{
    return lk_track_device_on(c, c ? c->stream : nullptr, prev, next, dPrev, dNext, dStatus, dTs, n, useInitial, maxIter, eps, minEig);
}
int hv_lk_track_device_on_stream(hv_ctx* c, void* cudaStream, hv_pyr* prev, hv_pyr* next, const float* dPrev, const float* dInit, float* dNext,
                                 uint8_t* dStatus, int32_t* dTs, int n, int maxIter, double eps, double minEig)
{
    return lk_track_device_on(c, (cudaStream_t)cudaStream, prev, next, dPrev, dNext, dStatus, dTs, n, dInit ? 1 : 0, maxIter, eps, minEig, dInit);
}

int hv_lk_track(hv_ctx* c, hv_pyr* prev, hv_pyr* next, const float* prevXY, float* nextXY, uint8_t* status,
                int32_t* trackStatus, int n, int useInitial, int maxIter, double eps, double minEig)
{
    if (!c || n < 0 || (n > 0 && (!prevXY || !nextXY))) { hv_set_error("hv_lk_track: invalid argument"); return HV_ERR_INVALID; }
    int rc = lk_check_pair("hv_lk_track", c, prev, next);
    if (rc != HV_OK) return rc;
    if (n == 0) return HV_OK;   // optical_flow.cpp:41-44: empty input, empty output
    HV_CUDA(cudaSetDevice(c->device));
    // staging block: [prev 8n | next 8n | trackStatus 4n | status n]
    const size_t oPrev = 0, oNext = 8 * (size_t)n, oTs = 16 * (size_t)n, oSt = 20 * (size_t)n, total = 21 * (size_t)n;
    rc = hv_ctx_reserve_stage(c, total);
    if (rc != HV_OK) return rc;
    uint8_t* hs = (uint8_t*)c->h_stage; uint8_t* ds = (uint8_t*)c->d_stage;
    memcpy(hs + oPrev, prevXY, 8 * (size_t)n);
    if (useInitial) memcpy(hs + oNext, nextXY, 8 * (size_t)n);
    if (hv_polling_enabled() && hv_lk_uses_cta_kernel(n)) {     // only the CTA-per-feature kernel raises the host flag
        // The kernel reads the points from and writes the results to the mapped pinned block itself (a few KB over PCIe) and
        // raises a flag there when the last feature is done: no H2D / D2H copy calls, no stream synchronisation.
        uint8_t* hd = (uint8_t*)c->hd_stage;
        LkLaunch L;
        LkJob& d = L.jobs[0];
        d.prevIdx = prev->slot; d.nextIdx = next->slot; d.n = n; d.useInitial = useInitial;
        d.prevPts = (const float2*)(hd + oPrev); d.nextPts = (float2*)(hd + oNext);
        d.status = hd + oSt; d.trackStatus = (int32_t*)(hd + oTs); d.initPts = nullptr;
        L.njobs = 1;
        lk_fill(L, c, HV_MAX_LEVELS, maxIter, eps, minEig);
        volatile unsigned* flag = (volatile unsigned*)(hs + c->stageBytes);
        L.doneCounter = c->d_done; c->doneCount += (unsigned)n; L.doneTarget = c->doneCount;
        L.seq = ++c->seq; L.hostFlag = (volatile unsigned*)(hd + c->stageBytes);
        cudaError_t e = hv_launch_lk(L, prev->win, c->stream);
        if (e == cudaErrorInvalidValue) { hv_set_error("hv_lk_track: window size %d unsupported (supported: 11, 15, 21, 31)", prev->win); return HV_ERR_UNSUPPORTED; }
        HV_CUDA(e);
        c->launches += 1;
        rc = hv_poll_flag(flag, L.seq, c->stream, "hv_lk_track");
        if (rc != HV_OK) {      // resynchronise the counter before anybody polls again
            cudaStreamSynchronize(c->stream); cudaMemsetAsync(c->d_done, 0, sizeof(unsigned), c->stream); cudaStreamSynchronize(c->stream); c->doneCount = 0;
            return rc;
        }
    } else {
        HV_CUDA(cudaMemcpyAsync(ds, hs, useInitial ? 16 * (size_t)n : 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        rc = hv_lk_track_device(c, prev, next, (const float*)(ds + oPrev), (float*)(ds + oNext), ds + oSt, (int32_t*)(ds + oTs),
                                n, useInitial, maxIter, eps, minEig);
        if (rc != HV_OK) return rc;
        HV_CUDA(cudaMemcpyAsync(hs + oNext, ds + oNext, total - oNext, cudaMemcpyDeviceToHost, c->stream));
        HV_CUDA(cudaStreamSynchronize(c->stream));
    }
    memcpy(nextXY, hs + oNext, 8 * (size_t)n);
    if (trackStatus) memcpy(trackStatus, hs + oTs, 4 * (size_t)n);
    if (status) memcpy(status, hs + oSt, (size_t)n);
    return HV_OK;
}

// ------------------------------------------------------------------------------------------------ corner detection (N2)
static int gftt_args(const char* who, hv_ctx* c, hv_pyr* pyr, int blockSize, int cell, float minResponse, GfttArgs& a)
{
    if (!c || !pyr || pyr->ctx != c) { hv_set_error("%s: invalid context / pyramid", who); return HV_ERR_INVALID; }
    if (blockSize != 3) { hv_set_error("%s: gfttBlockSize %d unsupported (3 only)", who, blockSize); return HV_ERR_UNSUPPORTED; }
    if (cell < 2 || cell > 32) { hv_set_error("%s: cell size %d unsupported (2..32)", who, cell); return HV_ERR_UNSUPPORTED; }
    const HvLevel& L = pyr->desc.lv[0];
    memset(&a, 0, sizeof(a));
    a.gray = L.gray; a.pitch = L.gpitch; a.w = L.w; a.h = L.h; a.cell = cell; a.minResponse = minResponse;
    const double scale = 1.0 / ((double)(1 << 2) * blockSize * 255.0);          // OCV/imgproc/src/corner.cpp:246-251
    a.k1 = (float)(1.0 * scale); a.k0 = (float)(2.0 * scale);
    return HV_OK;
}

int hv_gftt_cells(const hv_pyr* pyr, int cell, int* cellsX, int* cellsY)
{
    if (!pyr || cell <= 0) { hv_set_error("hv_gftt_cells: invalid argument"); return HV_ERR_INVALID; }
    if (cellsX) *cellsX = pyr->w / cell;
    if (cellsY) *cellsY = pyr->h / cell;
    return HV_OK;
}

int hv_gftt_detect_device(hv_ctx* c, hv_pyr* pyr, int blockSize, int cell, float minResponse, float* dKp)
{
    GfttArgs a;
    int rc = gftt_args("hv_gftt_detect_device", c, pyr, blockSize, cell, minResponse, a);
    if (rc != HV_OK) return rc;
    if (!dKp) { hv_set_error("hv_gftt_detect_device: NULL output"); return HV_ERR_INVALID; }
    HV_CUDA(cudaSetDevice(c->device));
    a.kp = dKp;
    HV_CUDA(hv_launch_gftt(a, c->stream));
    c->launches += 1;
    return HV_OK;
}

int hv_gftt_detect(hv_ctx* c, hv_pyr* pyr, int blockSize, int cell, float minResponse, float* kp)
{
    GfttArgs a;
    int rc = gftt_args("hv_gftt_detect", c, pyr, blockSize, cell, minResponse, a);
    if (rc != HV_OK) return rc;
    if (!kp) { hv_set_error("hv_gftt_detect: NULL output"); return HV_ERR_INVALID; }
    const int cells = (a.w / cell) * (a.h / cell);
    if (cells == 0) return HV_OK;
    HV_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)cells * 3 * sizeof(float);
    rc = hv_ctx_reserve_stage(c, bytes);
    if (rc != HV_OK) return rc;
    uint8_t* hs = (uint8_t*)c->h_stage;
    if (hv_polling_enabled()) {
        // the kernel writes the key points straight into the mapped pinned block and the last cell raises the flag (as the LK kernel does)
        a.kp = (float*)c->hd_stage;
        volatile unsigned* flag = (volatile unsigned*)(hs + c->stageBytes);
        a.doneCounter = c->d_done; c->doneCount += (unsigned)cells; a.doneTarget = c->doneCount;
        a.seq = ++c->seq; a.hostFlag = (volatile unsigned*)((uint8_t*)c->hd_stage + c->stageBytes);
        HV_CUDA(hv_launch_gftt(a, c->stream));
        c->launches += 1;
        rc = hv_poll_flag(flag, a.seq, c->stream, "hv_gftt_detect");
        if (rc != HV_OK) {
            cudaStreamSynchronize(c->stream); cudaMemsetAsync(c->d_done, 0, sizeof(unsigned), c->stream); cudaStreamSynchronize(c->stream); c->doneCount = 0;
            return rc;
        }
    } else {
        a.kp = (float*)c->d_stage;
        HV_CUDA(hv_launch_gftt(a, c->stream));
        c->launches += 1;
        HV_CUDA(cudaMemcpyAsync(hs, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
        HV_CUDA(cudaStreamSynchronize(c->stream));
    }
    memcpy(kp, hs, bytes);
    return HV_OK;
}

// ------------------------------------------------------------------------------------------------ frame ingest (N4)
struct hv_ingest {
    hv_ctx* ctx = nullptr;
    int w = 0, h = 0;
    uint8_t* d_raw = nullptr; size_t rawBytes = 0;     // the frame as it arrived (device)
    uint8_t* d_gray = nullptr;                          // gray before the remap (w x h, pitch w rounded up to 4)
    HvRemapEntry* d_table = nullptr;
};

int hv_ingest_create(hv_ctx* c, int w, int h, hv_ingest** out)
{
    if (!c || !out || w <= 0 || h <= 0 || w > 32767 || h > 32767) { hv_set_error("hv_ingest_create: invalid argument"); return HV_ERR_INVALID; }
    HV_CUDA(cudaSetDevice(c->device));
    hv_ingest* g = new hv_ingest;
    g->ctx = c; g->w = w; g->h = h;
    const size_t gp = (size_t)((w + 3) & ~3);
    cudaError_t e = cudaMalloc(&g->d_gray, gp * h);
    if (e != cudaSuccess) { delete g; hv_set_error("hv_ingest_create: %s", cudaGetErrorString(e)); return HV_ERR_OOM; }
    *out = g;
    return HV_OK;
}
int hv_ingest_destroy(hv_ingest* g)
{
    if (!g) return HV_OK;
    cudaStreamSynchronize(g->ctx->stream);
    cudaFree(g->d_raw); cudaFree(g->d_gray); cudaFree(g->d_table);
    delete g;
    return HV_OK;
}
int hv_ingest_set_remap(hv_ingest* g, const hv_remap_entry* table)
{
    if (!g) { hv_set_error("hv_ingest_set_remap: NULL handle"); return HV_ERR_INVALID; }
    static_assert(sizeof(hv_remap_entry) == sizeof(HvRemapEntry) && sizeof(HvRemapEntry) == 12, "remap entry layout");
    HV_CUDA(cudaSetDevice(g->ctx->device));
    HV_CUDA(cudaStreamSynchronize(g->ctx->stream));
    if (!table) { cudaFree(g->d_table); g->d_table = nullptr; return HV_OK; }
    const size_t bytes = sizeof(HvRemapEntry) * (size_t)g->w * g->h;
    if (!g->d_table) HV_CUDA(cudaMalloc(&g->d_table, bytes));
    HV_CUDA(cudaMemcpy(g->d_table, table, bytes, cudaMemcpyHostToDevice));
    return HV_OK;
}
int hv_ingest_frame(hv_ingest* g, const uint8_t* src, size_t stride, int channels, const double* coeff, hv_pyr* dst, uint8_t* grayOut)
{
    if (!g || !src || !dst || channels < 1 || channels > 4 || dst->ctx != g->ctx || dst->w != g->w || dst->h != g->h || stride < (size_t)g->w * channels) {
        hv_set_error("hv_ingest_frame: invalid argument"); return HV_ERR_INVALID;
    }
    hv_ctx* c = g->ctx;
    HV_CUDA(cudaSetDevice(c->device));
    const int w = g->w, h = g->h;
    const HvLevel& L0 = dst->desc.lv[0];
    const int gp = (w + 3) & ~3;
    const bool colour = channels > 1, remap = g->d_table != nullptr;
    if (!colour && !remap) {                                     // plain gray frame: exactly hv_pyr_build
        int rc = hv_pyr_build(dst, src, stride);
        if (rc != HV_OK) return rc;
    } else {
        const size_t need = stride * h;
        if (need > g->rawBytes) { cudaStreamSynchronize(c->stream); cudaFree(g->d_raw); g->d_raw = nullptr; g->rawBytes = 0; HV_CUDA(cudaMalloc(&g->d_raw, need)); g->rawBytes = need; }
        HV_CUDA(cudaMemcpyAsync(g->d_raw, src, need, cudaMemcpyHostToDevice, c->stream));          // the only trip of the frame over PCIe
        const uint8_t* cur = g->d_raw; int curPitch = (int)stride;
        if (colour) {
            float cf[4] = {0.299f, 0.587f, 0.114f, 0.0f};                                          // image.cpp:360-366
            if (coeff) for (int i = 0; i < 4; i++) cf[i] = i < channels ? (float)coeff[i] : 0.0f;
            uint8_t* out = remap ? g->d_gray : L0.gray; const int op = remap ? gp : L0.gpitch;
            HV_CUDA(hv_launch_gray(cur, curPitch, channels, w, h, cf, out, op, c->stream));
            c->launches += 1;
            cur = out; curPitch = op;
        }
        if (remap) { HV_CUDA(hv_launch_remap(cur, curPitch, w, h, g->d_table, L0.gray, L0.gpitch, c->stream)); c->launches += 1; }
        // level 0 of the pyramid now holds the ingested image: build the rest in place (no second copy of the frame)
        unsigned short idx = (unsigned short)dst->slot;
        const uint8_t* l0 = L0.gray; const int l0p = L0.gpitch, nl = dst->nlevels;
        HV_CUDA(hv_launch_pyr_fused(c->d_table, &idx, nullptr, nullptr, &l0, &l0p, &nl, 1, w, h, nl, c->stream));
        c->launches += 1;
    }
    if (grayOut) HV_CUDA(cudaMemcpy2DAsync(grayOut, (size_t)w, L0.gray, (size_t)L0.gpitch, (size_t)w, (size_t)h, cudaMemcpyDeviceToHost, c->stream));
    return HV_OK;
}

} // extern "C"
