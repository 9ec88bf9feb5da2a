// hybvio_b200/host/cuda_feature_detector.cpp -- tracker::FeatureDetector (src/tracker/feature_detector.hpp:20-77) on top of
// hv_gftt_detect (include/hybvio_b200.h): SURVEY.md 8(f) N2.
//
// Drop-in for the detector FeatureDetector::build("GPU-GFTT") hands out for CPU images (src/tracker/feature_detector.cpp:566-682):
// the corner response (cv::cornerMinEigenVal) and the per-cell maxima (CollectMax) run on the GPU, on the level-0 image that the
// pyramid build of the same frame has already put into HBM (cuda_tracker_backends.cpp keeps a host-pointer -> pyramid registry; a
// frame without a pyramid is uploaded once into a scratch pyramid); the rest of detect() -- stable sort by response, the
// `corners.resize(n)` + push_back sequence, applyMinDistance (feature_detector.cpp:625-638) -- is restated line by line on the host and
// calls the reference's own FeatureDetector::applyMinDistance (src/tracker/feature_detector_legacy.cpp, linked unchanged).
// Plug in: tracker::buildCudaFeatureDetector(w, h, parameters.tracker) where image.cpp:52 calls FeatureDetector::build, or compile this
// file with -DHV_REPLACE_FEATURE_DETECTOR instead of feature_detector.cpp (keep feature_detector_legacy.cpp): then it defines
// FeatureDetector::build itself ("FAST" / "GFTT" still go to the legacy builders).
#include "feature_detector.hpp"
#include "image.hpp"
#include "parameters.hpp"
#include "../../include/hybvio_b200.h"
#include "cuda_context.hpp"

#include <accelerated-arrays/cpu/image.hpp>
#include <accelerated-arrays/future.hpp>
#include <algorithm>
#include <cassert>
#include <vector>

namespace tracker {
hv_pyr* cudaPyramidOfHostImage(const void* data);      // cuda_tracker_backends.cpp

namespace {
using hybvio_b200::sharedContext;
#define HV(call) do { if ((call) != HV_OK) hybvio_b200::hvFail(#call); } while (0)

struct KeyPoint : Feature::Point { float response; };

class CudaFeatureDetector : public FeatureDetector {
    const int width, height;
    std::unique_ptr<accelerated::Processor> instant = accelerated::Processor::createInstant();
    hv_pyr* scratch = nullptr;
    std::vector<float> raw;
    std::vector<KeyPoint> keypoints;

    int cell() const {            // CollectMax: reduceFactors {4,4,2} / {4,4} / {4,2} (feature_detector.cpp:425-433)
        const int target = int(parameters.gfttMinDistance);
        return target >= 32 ? 32 : target >= 16 ? 16 : 8;
    }
public:
    CudaFeatureDetector(int w, int h, const odometry::ParametersTracker& p) : FeatureDetector(p), width(w), height(h) {}
    ~CudaFeatureDetector() override { if (scratch) hv_pyr_release(scratch); }

    void detect(Image& image, std::vector<Feature::Point>& corners, const std::vector<Feature::Point>& prevCorners, int maskRadius) final {
        detect(image.getAccImage(), corners, prevCorners, maskRadius).wait();
    }
    accelerated::Future detect(accelerated::Image& image, std::vector<Feature::Point>& corners, const std::vector<Feature::Point>& prevCorners,
                               int maskRadius) final {
        auto& cpu = accelerated::cpu::Image::castFrom(image);
        hv_pyr* pyr = cudaPyramidOfHostImage(cpu.getDataRaw());
        if (!pyr) {               // no pyramid of this frame on the device (detector used outside the tracker's order): upload it once
            if (!scratch) HV(hv_pyr_create(sharedContext(), width, height, parameters.pyrLKWindowSize, 0, &scratch));
            HV(hv_pyr_build(scratch, cpu.getDataRaw(), static_cast<size_t>(cpu.bytesPerRow())));
            pyr = scratch;
        }
        int cx = 0, cy = 0;
        const int bs = cell();
        HV(hv_gftt_cells(pyr, bs, &cx, &cy));
        raw.resize(static_cast<size_t>(cx) * cy * 3);
        HV(hv_gftt_detect(sharedContext(), pyr, parameters.gfttBlockSize, bs, parameters.gfttMinResponse, raw.data()));
        keypoints.resize(static_cast<size_t>(cx) * cy);
        for (size_t i = 0; i < keypoints.size(); i++) { keypoints[i].x = raw[3 * i]; keypoints[i].y = raw[3 * i + 1]; keypoints[i].response = raw[3 * i + 2]; }
        // feature_detector.cpp:625-638
        std::stable_sort(keypoints.begin(), keypoints.end(), [](const KeyPoint& a, const KeyPoint& b) -> bool { return a.response > b.response; });
        corners.clear();
        corners.resize(keypoints.size());
        for (const auto& kp : keypoints) corners.push_back(kp);
        if (maskRadius > 0) applyMinDistance(corners, prevCorners, maskRadius);
        return instant->enqueue([]() {});
    }
    bool supportsAsync() const final { return false; }
    void debugVisualize(cv::Mat&) final {}
};
} // namespace

std::unique_ptr<FeatureDetector> buildCudaFeatureDetector(int w, int h, const odometry::ParametersTracker& p) {
    return std::unique_ptr<FeatureDetector>(new CudaFeatureDetector(w, h, p));
}

#ifdef HV_REPLACE_FEATURE_DETECTOR
// compiled INSTEAD of src/tracker/feature_detector.cpp
FeatureDetector::~FeatureDetector() = default;
FeatureDetector::FeatureDetector(const odometry::ParametersTracker& p) : parameters(p) {}
std::unique_ptr<FeatureDetector> FeatureDetector::build(int w, int h, accelerated::Processor&, accelerated::Image::Factory&,
                                                        accelerated::operations::StandardFactory&, const odometry::ParametersTracker& p) {
    if (p.featureDetector == "FAST") return buildLegacyFAST(w, h, p);
    if (p.featureDetector == "GFTT") return buildLegacyGFTT(w, h, p);
    return buildCudaFeatureDetector(w, h, p);
}
#endif
} // namespace tracker
