// hybvio_b200/host/cuda_undistorter.cpp -- tracker::Undistorter (src/tracker/undistorter.hpp:14-40) on top of hv_ingest_* (include/hybvio_b200.h):
// SURVEY.md 8(f) N4. Same interface, same result image (host, from the same kind of allocator ring) and the same rectified camera as the
// reference's CPU branch (src/tracker/undistorter.cpp:77-118); the interpolation runs on the device and its result is ALSO left in level 0 of a
// device pyramid that is registered under the address of the output image, so that CudaImagePyramidFactory::compute (cuda_tracker_backends.cpp)
// finds the frame already in HBM, pyramid built, and does not upload it again.
// Plug in: tracker::buildCudaUndistorter(...) where image.cpp:323-336 calls Undistorter::buildRectified / buildMono, or compile with
// -DHV_REPLACE_UNDISTORTER instead of undistorter.cpp.
#include "undistorter.hpp"
#include "camera.hpp"
#include "parameters.hpp"
#include "undistort_table.hpp"
#include "cuda_context.hpp"
#include "../util/allocator.hpp"

#include <accelerated-arrays/cpu/image.hpp>
#include <accelerated-arrays/future.hpp>
#include <string>

namespace tracker {
hv_pyr* cudaAcquirePyramid(int w, int h, int win, int maxLevel);                          // cuda_tracker_backends.cpp
void cudaRegisterPrebuiltPyramid(const void* hostData, hv_pyr* p, int w, int h);

namespace {
using hybvio_b200::sharedContext;
#define HV(call) do { if ((call) != HV_OK) hybvio_b200::hvFail(#call); } while (0)

class CudaUndistorter : public Undistorter {
    accelerated::Image::Factory& ifac;
    util::Allocator<accelerated::Image> imageAllocator;
    std::shared_ptr<const Camera> undistortedCamera;
    const int width, height, win, maxLevel;
    hv_ingest* ingest = nullptr;
    std::string tableFor;                       // serialisation of the original camera the table was built for
    std::vector<hv_remap_entry> table;
public:
    CudaUndistorter(int w, int h, std::shared_ptr<const Camera> camera, accelerated::Image::Factory& f, const odometry::ParametersTracker& p)
        : ifac(f), imageAllocator([&f, w, h]() { return f.create(w, h, 1, accelerated::ImageTypeSpec::DataType::UFIXED8); }),
          undistortedCamera(camera), width(w), height(h), win(p.pyrLKWindowSize), maxLevel(p.pyrLKMaxLevel) {
        HV(hv_ingest_create(sharedContext(), w, h, &ingest));
    }
    ~CudaUndistorter() override { hv_ingest_destroy(ingest); }

    Result undistort(accelerated::Image& image, std::shared_ptr<const Camera> camera) final {
        auto outImage = imageAllocator.next();
        const std::string key = camera->serialize();
        if (key != tableFor) {                  // per-frame intrinsics are allowed by the interface: rebuild only when they change
            hybvio_b200::buildUndistortTable(*undistortedCamera, *camera, width, height, table);
            HV(hv_ingest_set_remap(ingest, table.data()));
            tableFor = key;
        }
        auto& in = accelerated::cpu::Image::castFrom(image);
        auto& out = accelerated::cpu::Image::castFrom(*outImage);
        hv_pyr* pyr = cudaAcquirePyramid(width, height, win, maxLevel);
        HV(hv_ingest_frame(ingest, in.getDataRaw(), static_cast<size_t>(in.bytesPerRow()), image.channels, nullptr, pyr, out.getDataRaw()));
        HV(hv_ctx_sync(sharedContext()));       // the host copy is consumed by CPU code (cornerSubPix, SLAM) right away
        cudaRegisterPrebuiltPyramid(out.getDataRaw(), pyr, width, height);   // the tracker's pyramid request for this image finds it built
        return { .camera = undistortedCamera, .image = outImage, .future = accelerated::Future::instantlyResolved() };
    }
};
} // namespace

std::unique_ptr<Undistorter> buildCudaUndistorter(int w, int h, std::shared_ptr<const Camera> rectifiedCamera, accelerated::Image::Factory& ifac,
                                                  const odometry::ParametersTracker& p) {
    return std::unique_ptr<Undistorter>(new CudaUndistorter(w, h, rectifiedCamera, ifac, p));
}

#ifdef HV_REPLACE_UNDISTORTER
// compiled INSTEAD of src/tracker/undistorter.cpp
std::unique_ptr<Undistorter> Undistorter::buildRectified(int w, int h, std::shared_ptr<const Camera> rectifiedCamera, accelerated::Image::Factory& ifac,
                                                         accelerated::operations::StandardFactory&, const odometry::ParametersTracker& p) {
    if (!p.useRectification) return {};
    return buildCudaUndistorter(w, h, rectifiedCamera, ifac, p);
}
std::unique_ptr<Undistorter> Undistorter::buildMono(int w, int h, float focalLength, accelerated::Image::Factory& ifac,
                                                    accelerated::operations::StandardFactory&, const odometry::ParametersTracker& p) {
    if (!p.useRectification) return {};
    api::CameraParameters i;                    // undistorter.cpp:157-163
    i.focalLengthX = focalLength * p.rectificationZoom; i.focalLengthY = focalLength * p.rectificationZoom;
    i.principalPointX = w * 0.5f; i.principalPointY = h * 0.5f;
    std::shared_ptr<const Camera> camera = Camera::buildPinhole(i, {}, w, h, nullptr);
    return buildCudaUndistorter(w, h, camera, ifac, p);
}
Undistorter::~Undistorter() = default;
#endif
} // namespace tracker
