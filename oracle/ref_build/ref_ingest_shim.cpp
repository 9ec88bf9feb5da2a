// TEST INFRASTRUCTURE (oracle/_ref/libref_ingest.so): the reference's OWN frame ingest on CPU images -- colour -> gray through
// accelerated-arrays' pixelwiseAffine exactly as src/tracker/image.cpp:360-366 builds it, and Undistorter::buildMono(...)->undistort
// (src/tracker/undistorter.cpp, unmodified) for a distorted pinhole / fisheye camera -- plus the camera mapping table
// (hybvio_b200/host/undistort_table.hpp, which calls the reference's Camera classes). Pins oracle orc_gray / orc_remap and provides the
// golden vectors of tests/golden/ingest_golden.npz.
#include "undistorter.hpp"
#include "camera.hpp"
#include "parameters.hpp"
#include "undistort_table.hpp"
#include <accelerated-arrays/cpu/image.hpp>
#include <accelerated-arrays/cpu/operations.hpp>
#include <accelerated-arrays/future.hpp>
#include <accelerated-arrays/standard_ops.hpp>
#include <cstring>

namespace {
struct Env {
    std::unique_ptr<accelerated::Queue> queue = accelerated::Processor::createQueue();
    std::unique_ptr<accelerated::Image::Factory> images = accelerated::cpu::Image::createFactory();
    std::unique_ptr<accelerated::operations::StandardFactory> ops = accelerated::cpu::operations::createFactory(*queue);
};
std::shared_ptr<const tracker::Camera> makeCamera(int fisheye, double fx, double fy, double cx, double cy, const double* dist, int ndist, int w, int h) {
    api::CameraParameters k; k.focalLengthX = fx; k.focalLengthY = fy; k.principalPointX = cx; k.principalPointY = cy;
    std::vector<double> d(dist, dist + ndist);
    if (fisheye) return tracker::Camera::buildFisheye(k, d);
    return tracker::Camera::buildPinhole(k, d, w, h, nullptr);
}
}

extern "C" {
void hv_ref_gray(const uint8_t* src, int w, int h, int channels, uint8_t* out)
{
    Env env;
    auto in = accelerated::cpu::Image::createReference(w, h, channels, accelerated::ImageTypeSpec::DataType::UFIXED8, const_cast<uint8_t*>(src));
    auto dst = accelerated::cpu::Image::createReference(w, h, 1, accelerated::ImageTypeSpec::DataType::UFIXED8, out);
    const auto graySpec = env.images->getSpec(1, accelerated::ImageTypeSpec::DataType::UFIXED8);
    std::vector<double> coeff = { 0.299, 0.587, 0.114 };                       // image.cpp:360-364
    if (channels == 4) coeff.push_back(0);
    auto op = env.ops->pixelwiseAffine({ coeff }).build(*in, graySpec);
    accelerated::operations::callUnary(op, *in, *dst);
    env.queue->processAll();
}
// Undistorter::buildMono + undistort: original camera = (fisheye ? Kannala-Brandt : distorted pinhole)(fx, fy, cx, cy, dist)
int hv_ref_undistort_mono(const uint8_t* src, int w, int h, int fisheye, double fx, double fy, double cx, double cy, const double* dist, int ndist,
                          double zoom, uint8_t* out, hv_remap_entry* table)
{
    Env env;
    odometry::Parameters params;
    params.tracker.useRectification = true; params.tracker.rectificationZoom = (float)zoom;
    auto cam = makeCamera(fisheye, fx, fy, cx, cy, dist, ndist, w, h);
    auto und = tracker::Undistorter::buildMono(w, h, (float)cam->getFocalLength(), *env.images, *env.ops, params.tracker);
    if (!und) return -1;
    auto in = accelerated::cpu::Image::createReference(w, h, 1, accelerated::ImageTypeSpec::DataType::UFIXED8, const_cast<uint8_t*>(src));
    auto res = und->undistort(*in, cam);
    res.future.wait();
    auto& o = accelerated::cpu::Image::castFrom(*res.image);
    std::memcpy(out, o.getDataRaw(), (size_t)w * h);
    if (table) {
        std::vector<hv_remap_entry> t;
        hybvio_b200::buildUndistortTable(*res.camera, *cam, w, h, t);
        std::memcpy(table, t.data(), t.size() * sizeof(hv_remap_entry));
    }
    return 0;
}
}
