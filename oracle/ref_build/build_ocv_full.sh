#!/bin/bash
# TEST INFRASTRUCTURE: the vendored OpenCV 4.3 (third-party dependency of the reference, not the reference's own code)
# configured with its own CMakeLists exactly as SURVEY.md Appendix A step 2/5 verified:
#   BUILD_LIST=core,imgproc,video,calib3d,features2d,flann  -> static libs under oracle/_ref/ocv_full/install
# Needed only by the WHOLE-CORE pipeline harness (build_pipeline.sh: tracker.cpp, backend.cpp, control.cpp ... unmodified),
# which needs calib3d / features2d / flann and all of imgproc; the pyramid / LK / EKF oracles keep their plain-g++ recipes.
# Runs only where /root/reference exists (this container); the GPU box uses the prebuilt binaries.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${OUT:-$HERE/../_ref}
M=$REF/3rdparty/mobile-cv-suite
B=$OUT/ocv_full
[ -f "$B/install/lib/libopencv_calib3d.a" ] && { echo "ocv_full already built"; exit 0; }
mkdir -p "$B"
cmake -S "$M/opencv" -B "$B/build" -G Ninja -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
  -DBUILD_LIST=core,imgproc,video,calib3d,features2d,flann -DBUILD_SHARED_LIBS=OFF -DCMAKE_INSTALL_PREFIX="$B/install" \
  -DCMAKE_POSITION_INDEPENDENT_CODE=ON \
  -DWITH_IPP=OFF -DWITH_ITT=OFF -DWITH_OPENCL=OFF -DWITH_TBB=OFF -DWITH_OPENMP=OFF -DWITH_EIGEN=OFF -DWITH_LAPACK=OFF \
  -DWITH_PROTOBUF=OFF -DWITH_QUIRC=OFF -DWITH_TIFF=OFF -DWITH_JASPER=OFF -DWITH_JPEG=OFF -DWITH_PNG=OFF -DWITH_WEBP=OFF \
  -DWITH_OPENEXR=OFF -DWITH_FFMPEG=OFF -DWITH_GSTREAMER=OFF -DWITH_GTK=OFF -DWITH_V4L=OFF -DWITH_1394=OFF -DWITH_ADE=OFF \
  -DBUILD_TESTS=OFF -DBUILD_PERF_TESTS=OFF -DBUILD_EXAMPLES=OFF -DBUILD_opencv_apps=OFF -DBUILD_JAVA=OFF \
  -DBUILD_opencv_python2=OFF -DBUILD_opencv_python3=OFF > "$B/cmake.log" 2>&1
ninja -C "$B/build" install > "$B/ninja.log" 2>&1
echo built "$B/install"
