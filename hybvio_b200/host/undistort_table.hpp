// hybvio_b200/host/undistort_table.hpp -- the camera mapping of UndistorterImplementation::undistort (src/tracker/undistorter.cpp:53-57,
// 84-96) evaluated with the reference's OWN Camera classes, once per (rectified camera, original camera) pair, into the table the
// device kernel interpolates from (hv_remap_entry, include/hybvio_b200.h). Pure host code (no CUDA): shared by the adapter
// (cuda_undistorter.cpp) and by the test shim oracle/ref_build/ref_ingest_shim.cpp.
#ifndef HYBVIO_B200_HOST_UNDISTORT_TABLE_HPP_
#define HYBVIO_B200_HOST_UNDISTORT_TABLE_HPP_
#include "camera.hpp"
#include "../../include/hybvio_b200.h"
#include <cmath>
#include <vector>

namespace hybvio_b200 {
inline void buildUndistortTable(const tracker::Camera& rectifiedCamera, const tracker::Camera& origCamera, int w, int h, std::vector<hv_remap_entry>& table)
{
    table.resize(static_cast<size_t>(w) * h);
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            hv_remap_entry e { HV_REMAP_INVALID_X0, 0, 0.0f, 0.0f };
            Eigen::Vector2d pixRect(x, y), pixOrig;
            Eigen::Vector3d ray;
            if (rectifiedCamera.pixelToRay(pixRect, ray) && origCamera.rayToPixel(ray, pixOrig)) {              // undistortCpu
                if (pixOrig(0) >= 0 && pixOrig(0) < w && pixOrig(1) >= 0 && pixOrig(1) < h) {                   // undistorter.cpp:91
                    const int x0 = int(std::floor(pixOrig(0))), y0 = int(std::floor(pixOrig(1)));
                    const float xfrac = pixOrig(0) - x0, yfrac = pixOrig(1) - y0;                               // undistorter.cpp:94
                    e.x0 = static_cast<int16_t>(x0); e.y0 = static_cast<int16_t>(y0); e.xfrac = xfrac; e.yfrac = yfrac;
                }
            }
            table[static_cast<size_t>(y) * w + x] = e;
        }
    }
}
} // namespace hybvio_b200
#endif
