// hybvio_b200/host/cuda_tracker_backends.cpp -- tracker::ImagePyramid / ImagePyramid::Factory / tracker::OpticalFlow
// implemented on top of the hv_pyr_* / hv_lk_* C ABI (include/hybvio_b200.h).
//
// Plugs into the single seam where the reference chooses its back ends, ImageImplementation::SharedData
// (src/tracker/image.cpp:55-56): replace
//     ImagePyramid::Factory::buildOpenCv(parameters.tracker) -> tracker::buildCudaImagePyramidFactory(parameters.tracker)
//     OpticalFlow::buildOpenCv(parameters.tracker)           -> tracker::buildCudaOpticalFlow(parameters.tracker)
// (a two-line change, or compile this file instead of image_pyramid.cpp / optical_flow.cpp and keep the buildOpenCv
// names: define HV_REPLACE_OPENCV_BACKENDS). tracker.cpp, image.cpp and everything above them are unchanged: the
// pyramid stays an opaque tracker::ImagePyramid handed from Image::opticalFlow to OpticalFlow::compute
// (src/tracker/image.cpp:87-106), and getOpenCv() is only ever called by the OpenCV flow back end
// (src/tracker/optical_flow.cpp:96-97), which this file replaces.
#include "image_pyramid.hpp"
#include "optical_flow.hpp"
#include "parameters.hpp"
#include "../../include/hybvio_b200.h"
#include "cuda_context.hpp"

#include <accelerated-arrays/cpu/image.hpp>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace cv { class Mat; }

namespace tracker {
namespace {
using hybvio_b200::sharedContext;   // ONE context / stream per process, shared with CudaEKF (cuda_context.hpp)
[[noreturn]] void fail(const char* what) { hybvio_b200::hvFail(what); }
#define HV(call) do { if ((call) != HV_OK) fail(#call); } while (0)

// Pool of device pyramids, recycled like the reference's util::Allocator ring (image_pyramid.cpp:31,37): a pyramid
// returns to the pool when the last tracker::Image holding it dies (prevImage, SLAM queue, ...).
struct PyramidPool {
    std::mutex mutex;
    std::vector<hv_pyr*> free;
    int w = 0, h = 0, win = 0, maxLevel = 0;
    hv_pyr* acquire(int width, int height, int win_, int maxLevel_) {
        std::lock_guard<std::mutex> lock(mutex);
        if (width != w || height != h || win_ != win || maxLevel_ != maxLevel) {
            for (hv_pyr* p : free) hv_pyr_release(p);
            free.clear(); w = width; h = height; win = win_; maxLevel = maxLevel_;
        }
        if (!free.empty()) { hv_pyr* p = free.back(); free.pop_back(); return p; }
        hv_pyr* p = nullptr;
        HV(hv_pyr_create(sharedContext(), width, height, win_, maxLevel_, &p));
        return p;
    }
    void release(hv_pyr* p, int width, int height) {
        std::lock_guard<std::mutex> lock(mutex);
        if (width == w && height == h) free.push_back(p); else hv_pyr_release(p);
    }
};

// Which device pyramid holds a given host image (key: address of the image data): lets the corner detector adapter
// (cuda_feature_detector.cpp) run on the level-0 image that is already in HBM instead of uploading the frame a second time. The
// tracker builds an image's pyramid (optical flow, tracker.cpp:395-438) before it detects new features on it (tracker.cpp:531).
struct PyramidRegistry {
    std::mutex mutex;
    std::vector<std::pair<const void*, hv_pyr*>> entries;
    void add(const void* key, hv_pyr* p) { std::lock_guard<std::mutex> l(mutex); remove_locked(key); entries.emplace_back(key, p); }
    void remove(const void* key, hv_pyr* p) {
        std::lock_guard<std::mutex> l(mutex);
        for (size_t i = 0; i < entries.size(); i++) if (entries[i].first == key && entries[i].second == p) { entries.erase(entries.begin() + i); return; }
    }
    hv_pyr* find(const void* key) { std::lock_guard<std::mutex> l(mutex); for (auto& e : entries) if (e.first == key) return e.second; return nullptr; }
private:
    void remove_locked(const void* key) { for (size_t i = 0; i < entries.size(); i++) if (entries[i].first == key) { entries.erase(entries.begin() + i); return; } }
};
PyramidRegistry& registry() { static PyramidRegistry r; return r; }

struct CudaImagePyramid : ImagePyramid {
    std::shared_ptr<PyramidPool> pool;
    hv_pyr* pyr = nullptr;
    int width = 0, height = 0;
    const void* key = nullptr;
    std::shared_ptr<accelerated::Image> source;   // keeps the host gray image alive until the async H2D has been consumed

    ~CudaImagePyramid() override { if (pyr) { registry().remove(key, pyr); pool->release(pyr, width, height); } }
    // The device layout is not an accelerated::Image; like the reference's CPU pyramid (image_pyramid.cpp:17-25) these
    // two accessors are not used by any caller.
    accelerated::Image& getGrayLevel(std::size_t) final { assert(false && "device-resident pyramid"); std::abort(); }
    accelerated::Image& getGradientLevel(std::size_t) final { assert(false && "device-resident pyramid"); std::abort(); }
    const std::vector<cv::Mat>& getOpenCv() final { assert(false && "device-resident pyramid: no cv::Mat view"); std::abort(); }
};

// One pool per process (the adapters share one context): the pyramid factory and the ingest adapter (cuda_undistorter.cpp) draw from it.
std::shared_ptr<PyramidPool> globalPool() { static std::shared_ptr<PyramidPool> p = std::make_shared<PyramidPool>(); return p; }

// Pyramids that the frame ingest has ALREADY built on the device (the frame was undistorted there and never needs to be uploaded again),
// keyed by the address of the host copy of the ingested image; taken over by the pyramid factory when the tracker asks for that image's
// pyramid. At most 8 wait here: a frame whose pyramid is never requested goes back to the pool.
struct PrebuiltPyramids {
    std::mutex mutex;
    std::vector<std::pair<const void*, hv_pyr*>> entries;
    void add(const void* key, hv_pyr* p, int w, int h) {
        std::lock_guard<std::mutex> l(mutex);
        for (size_t i = 0; i < entries.size(); i++) if (entries[i].first == key) { globalPool()->release(entries[i].second, w, h); entries.erase(entries.begin() + i); break; }
        if (entries.size() >= 8) { globalPool()->release(entries.front().second, w, h); entries.erase(entries.begin()); }
        entries.emplace_back(key, p);
    }
    hv_pyr* take(const void* key) {
        std::lock_guard<std::mutex> l(mutex);
        for (size_t i = 0; i < entries.size(); i++) if (entries[i].first == key) { hv_pyr* p = entries[i].second; entries.erase(entries.begin() + i); return p; }
        return nullptr;
    }
};
PrebuiltPyramids& prebuilt() { static PrebuiltPyramids p; return p; }

class CudaImagePyramidFactory : public ImagePyramid::Factory {
    const odometry::ParametersTracker& parameters;
    std::shared_ptr<PyramidPool> pool = globalPool();
public:
    explicit CudaImagePyramidFactory(const odometry::ParametersTracker& p) : parameters(p) {}
    std::shared_ptr<ImagePyramid> compute(std::shared_ptr<accelerated::Image> img) final {
        assert(img->channels == 1 && img->bytesPerChannel() == 1);   // gray u8 (ImagePyramid::GrayType)
        auto& cpu = accelerated::cpu::Image::castFrom(*img);
        auto pyramid = std::make_shared<CudaImagePyramid>();
        pyramid->pool = pool; pyramid->width = img->width; pyramid->height = img->height; pyramid->source = img;
        pyramid->key = cpu.getDataRaw();
        if (hv_pyr* pre = prebuilt().take(pyramid->key)) {
            pyramid->pyr = pre;                 // built on the device by the frame ingest (cuda_undistorter.cpp): nothing to copy, nothing to launch
        } else {
            pyramid->pyr = pool->acquire(img->width, img->height, parameters.pyrLKWindowSize, parameters.pyrLKMaxLevel);
            // H2D copy + one fused kernel, asynchronous on the context stream (cv::buildOpticalFlowPyramid in the reference)
            // (raw pointer: the reference's gray type is FixedPoint<uint8_t>, ImagePyramid::GrayType; getData<uint8_t>() would reject it)
            HV(hv_pyr_build(pyramid->pyr, cpu.getDataRaw(), static_cast<size_t>(cpu.bytesPerRow())));
        }
        registry().add(pyramid->key, pyramid->pyr);
        return pyramid;
    }
};

class CudaOpticalFlow : public OpticalFlow {
    const odometry::ParametersTracker& parameters;
    std::vector<std::int32_t> status32;
public:
    explicit CudaOpticalFlow(const odometry::ParametersTracker& p) : parameters(p) {}
    void compute(ImagePyramid& prevImagePyramid, ImagePyramid& imagePyramid, const std::vector<Feature::Point>& prevCorners,
                 std::vector<Feature::Point>& corners, std::vector<Feature::Status>& trackStatus, bool useInitialCorners,
                 int overrideMaxIterations) final {
        auto& prev = static_cast<CudaImagePyramid&>(prevImagePyramid);
        auto& next = static_cast<CudaImagePyramid&>(imagePyramid);
        const int n = static_cast<int>(prevCorners.size());
        trackStatus.clear();
        trackStatus.resize(prevCorners.size(), Feature::Status::FAILED_FLOW);
        if (n == 0) { corners.clear(); return; }                       // optical_flow.cpp:41-44
        if (!useInitialCorners) corners.resize(prevCorners.size());
        assert(corners.size() == prevCorners.size());
        status32.resize(n);
        static_assert(sizeof(Feature::Point) == 2 * sizeof(float), "Feature::Point is {float x, y}");
        const int maxIter = overrideMaxIterations > 0 ? overrideMaxIterations : parameters.pyrLKMaxIter;
        // TRACKED / FAILED_FLOW / FLOW_OUT_OF_RANGE are derived on the device exactly as optical_flow.cpp:52-58 does
        HV(hv_lk_track(sharedContext(), prev.pyr, next.pyr, reinterpret_cast<const float*>(prevCorners.data()),
                       reinterpret_cast<float*>(corners.data()), nullptr, status32.data(), n, useInitialCorners ? 1 : 0, maxIter,
                       parameters.pyrLKEpsilon, parameters.pyrLKMinEigThreshold));
        for (int i = 0; i < n; i++) trackStatus[i] = static_cast<Feature::Status>(status32[i]);
    }
};
} // namespace

std::unique_ptr<ImagePyramid::Factory> buildCudaImagePyramidFactory(const odometry::ParametersTracker& p) {
    return std::unique_ptr<ImagePyramid::Factory>(new CudaImagePyramidFactory(p));
}
std::unique_ptr<OpticalFlow> buildCudaOpticalFlow(const odometry::ParametersTracker& p) {
    return std::unique_ptr<OpticalFlow>(new CudaOpticalFlow(p));
}

// frame ingest (cuda_undistorter.cpp): a pyramid from the shared pool / hand-over of a pyramid already built from the image at `hostData`
hv_pyr* cudaAcquirePyramid(int w, int h, int win, int maxLevel) { return globalPool()->acquire(w, h, win, maxLevel); }
void cudaRegisterPrebuiltPyramid(const void* hostData, hv_pyr* p, int w, int h) { prebuilt().add(hostData, p, w, h); }

// the device pyramid whose level 0 is the host image at `data`, or NULL (cuda_feature_detector.cpp)
hv_pyr* cudaPyramidOfHostImage(const void* data) { return registry().find(data); }

// test harnesses only (oracle/ref_build/pipeline): the device pyramid behind a tracker::ImagePyramid built by this file
hv_pyr* cudaPyramidHandle(ImagePyramid& p) { return static_cast<CudaImagePyramid&>(p).pyr; }

#ifdef HV_REPLACE_OPENCV_BACKENDS
// compiled INSTEAD of src/tracker/image_pyramid.cpp and optical_flow.cpp: same symbols, CUDA back ends
std::unique_ptr<ImagePyramid::Factory> ImagePyramid::Factory::buildOpenCv(const odometry::ParametersTracker& p) { return buildCudaImagePyramidFactory(p); }
std::unique_ptr<OpticalFlow> OpticalFlow::buildOpenCv(const odometry::ParametersTracker& p) { return buildCudaOpticalFlow(p); }
ImagePyramid::~ImagePyramid() = default;
ImagePyramid::Factory::~Factory() = default;
OpticalFlow::~OpticalFlow() = default;
#endif
} // namespace tracker
