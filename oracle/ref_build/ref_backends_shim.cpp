// TEST INFRASTRUCTURE (oracle/_ref/libref_backends.so): the reference's OWN back ends -- src/tracker/image_pyramid.cpp,
// src/tracker/optical_flow.cpp, src/odometry/ekf.cpp, compiled unmodified -- reachable through three factory functions, so
// that ONE process can hold the reference back ends and the CUDA back ends side by side (oracle/ref_build/pipeline/). The
// version script hides every other symbol (the reference's ImagePyramid::Factory::buildOpenCv, OpticalFlow::buildOpenCv and
// EKF::build would otherwise collide with the harness's own definitions of the same names).
#include "image_pyramid.hpp"
#include "optical_flow.hpp"
#include "ekf.hpp"
#include "parameters.hpp"
#include <opencv2/core.hpp>

extern "C" {
tracker::ImagePyramid::Factory* hv_ref_build_pyramid_factory(const odometry::ParametersTracker* p) {
    return tracker::ImagePyramid::Factory::buildOpenCv(*p).release();       // src/tracker/image_pyramid.cpp:51-53
}
tracker::OpticalFlow* hv_ref_build_optical_flow(const odometry::ParametersTracker* p) {
    return tracker::OpticalFlow::buildOpenCv(*p).release();                 // src/tracker/optical_flow.cpp:105-107
}
odometry::EKF* hv_ref_build_ekf(const odometry::Parameters* p) {
    return odometry::EKF::build(*p).release();                              // src/odometry/ekf.cpp:1087-1092
}
void hv_ref_set_num_threads(int n) { cv::setNumThreads(n); }
int hv_ref_get_num_threads(void) { return cv::getNumThreads(); }
}
