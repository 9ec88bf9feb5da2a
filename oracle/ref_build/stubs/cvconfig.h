/* Hand-written build configuration for compiling the vendored OpenCV 4.3 sources of the
 * reference directly with g++ (no cmake). Test infrastructure only (oracle/_ref). */
#ifndef HV_REF_CVCONFIG_H
#define HV_REF_CVCONFIG_H
#define HAVE_PTHREAD 1
#define HAVE_PTHREADS_PF 1
#define OPENCV_TRACE 1
#endif
