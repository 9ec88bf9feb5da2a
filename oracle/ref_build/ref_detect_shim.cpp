// TEST INFRASTRUCTURE (oracle/_ref/libref_detect.so): the reference's OWN corner detector -- src/tracker/feature_detector.cpp,
// feature_detector_legacy.cpp, compiled unmodified -- on a CPU image, i.e. the path FeatureDetector::build("GPU-GFTT") takes when the
// images live in CPU memory (feature_detector.cpp:666-669: CpuCornerResponse + CollectMax::cpuImplementation), plus the response map
// cv::cornerMinEigenVal it is built on (feature_detector.cpp:296). Pins oracle/hv_oracle_gftt.c (tests/test_oracle_gftt.py).
#include "feature_detector.hpp"
#include "parameters.hpp"
#include <accelerated-arrays/cpu/image.hpp>
#include <accelerated-arrays/cpu/operations.hpp>
#include <accelerated-arrays/future.hpp>
#include <opencv2/imgproc.hpp>
#include <cstring>

extern "C" {
// corners after detect(image, corners, prev, maskRadius): returns the count (<= cap), out = x0 y0 x1 y1 ...
int hv_ref_detect(const uint8_t* img, int w, int h, int maxTracks, double minDistance, float minResponse, const float* prev, int nprev, int maskRadius,
                  float* out, int cap)
{
    odometry::Parameters params;
    params.tracker.featureDetector = "GPU-GFTT";
    params.tracker.maxTracks = maxTracks; params.tracker.gfttMinDistance = minDistance; params.tracker.gfttMinResponse = minResponse;
    auto queue = accelerated::Processor::createQueue();
    auto images = accelerated::cpu::Image::createFactory();
    auto ops = accelerated::cpu::operations::createFactory(*queue);
    auto det = tracker::FeatureDetector::build(w, h, *queue, *images, *ops, params.tracker);
    auto acc = accelerated::cpu::Image::createReference(w, h, 1, accelerated::ImageTypeSpec::DataType::UFIXED8, const_cast<uint8_t*>(img));
    std::vector<tracker::Feature::Point> corners, prevCorners(nprev);
    for (int i = 0; i < nprev; i++) prevCorners[i] = { prev[2 * i], prev[2 * i + 1] };
    det->detect(*acc, corners, prevCorners, maskRadius).wait();
    queue->processAll();
    const int n = (int)corners.size() < cap ? (int)corners.size() : cap;
    for (int i = 0; i < n; i++) { out[2 * i] = corners[i].x; out[2 * i + 1] = corners[i].y; }
    return (int)corners.size();
}
void hv_ref_corner_min_eigen_val(const uint8_t* img, int w, int h, int blockSize, float* response)
{
    cv::Mat src(h, w, CV_8UC1, const_cast<uint8_t*>(img)), dst(h, w, CV_32FC1, response);
    cv::cornerMinEigenVal(src, dst, blockSize, 3);                       // feature_detector.cpp:296
}
void hv_ref_detect_set_threads(int n) { cv::setNumThreads(n); }
}
