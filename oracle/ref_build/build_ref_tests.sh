#!/bin/bash
# TEST INFRASTRUCTURE: builds oracle/_ref/run_ref_ekf_tests = the reference's OWN Catch2 unit tests for the EKF
# (test/ekf.cpp: "chi-squared innovation test", "der_predict", "tranformTo"; test/test_main.cpp, test/helpers.cpp),
# compiled UNMODIFIED, but linked against hybvio_b200/host/cuda_ekf.cpp (+ libhybvio_b200.so) instead of the
# reference's src/odometry/ekf.cpp. Passing it on a B200 shows that CudaEKF is a drop-in for odometry::EKF.
# The fixtures test/data/{P,m}.csv are copied next to the binary (build output, git-ignored).
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=${OUT:-$HERE/../_ref}
M=$REF/3rdparty/mobile-cv-suite
OCV=$M/opencv/modules
[ -f "$OUT/gen/output/parameters.hpp" ] || "$HERE/build_ekf.sh"
mkdir -p "$OUT/obj_tests" "$OUT/inc" "$OUT/test/data"
ln -sfn "$M/accelerated-arrays/src" "$OUT/inc/accelerated-arrays"
ln -sfn "$M/jsonl-recorder" "$OUT/inc/jsonl-recorder"
# test/ekf.cpp includes "../src/odometry/parameters.hpp", a dangling symlink in the read-only tree: shadow it
mkdir -p "$OUT/inc/src/odometry" "$OUT/inc/shadow" "$OUT/inc/src/shadow"
ln -sfn "$OUT/gen/output/parameters.hpp" "$OUT/inc/src/odometry/parameters.hpp"
cp "$REF/test/data/P.csv" "$REF/test/data/m.csv" "$OUT/test/data/"
FL="-std=c++17 -O1 -w -DEIGEN_MPL2_ONLY -DEIGEN_DONT_PARALLELIZE"
INC="-I$OUT/gen/output -I$M/eigen -I$M/json/single_include -I$REF/src/odometry -I$REF/src/tracker -I$REF/src -I$REF/test -I$OUT/inc -I$OUT/inc/shadow -I$OUT/inc/src/shadow \
  -I$HERE/stubs -I$OCV/core/include -I$OCV/imgproc/include -I$OCV/video/include -I$OCV/calib3d/include -I$OCV/features2d/include \
  -I$OCV/flann/include -I$OCV/highgui/include -I$OCV/imgcodecs/include -I$OCV/videoio/include -I$OCV/../include"
O=$OUT/obj_tests
g++ $FL $INC -c "$REF/test/ekf.cpp" -o $O/test_ekf.o &
g++ $FL $INC -c "$REF/test/test_main.cpp" -o $O/test_main.o &
g++ $FL $INC -c "$REF/test/helpers.cpp" -o $O/helpers.o &
g++ $FL $INC -I"$ROOT/hybvio_b200/host" -c "$ROOT/hybvio_b200/host/cuda_ekf.cpp" -o $O/cuda_ekf.o &
wait
# parameters.o / odo_util.o ... come from build_ekf.sh (the reference's own support objects, minus ekf.o)
g++ -o "$OUT/run_ref_ekf_tests" $O/test_ekf.o $O/test_main.o $O/helpers.o $O/cuda_ekf.o \
    $OUT/obj_ekf/parameters.o $OUT/obj_ekf/odo_util.o $OUT/obj_ekf/timer.o $OUT/obj_ekf/util_util.o $OUT/obj_ekf/parameter_parser.o \
    -Wl,--gc-sections -L"$ROOT/hybvio_b200" -lhybvio_b200 -Wl,-rpath,'$ORIGIN/../../hybvio_b200' -lpthread
echo built $OUT/run_ref_ekf_tests

# Second suite: the reference's triangulation tests (test/triangulation.cpp: "visual", "stereo_visual", pinv, two-camera
# triangulation and the derivative checks), which use odometry::EKF as the state store behind extractCameraPoseTrail /
# prepareVisualUpdate (src/odometry/triangulation.cpp, unmodified) -- again linked against CudaEKF instead of ekf.cpp.
mkdir -p "$OUT/inc/fake/tracker" "$OUT/inc/fake/odometry"
ln -sfn "$OUT/gen/output/parameters.hpp" "$OUT/inc/fake/odometry/parameters.hpp"     # "../odometry/parameters.hpp" seen from src/tracker
INC2="$INC -I$OUT/inc/fake/tracker -I$OCV/photo/include -I$OCV/dnn/include -I$OCV/ml/include -I$OCV/objdetect/include -I$OCV/stitching/include"
[ -f "$OUT/obj_lk/video_lkpyramid.o" ] || make -C "$HERE" -f Makefile.lk -j8
g++ $FL $INC2 -c "$REF/test/triangulation.cpp" -o $O/test_triangulation.o &
g++ $FL $INC2 -c "$REF/src/odometry/triangulation.cpp" -o $O/ref_triangulation.o &
g++ $FL $INC2 -c "$REF/src/tracker/camera.cpp" -o $O/ref_camera.o &
g++ $FL $INC2 -c "$REF/src/tracker/util.cpp" -o $O/ref_tracker_util.o &
wait
OCVOBJ=$(ls $OUT/obj_lk/*.o $OUT/obj_lk/core_utils/*.o | grep -v shim.o)
g++ -o "$OUT/run_ref_triangulation_tests" $O/test_triangulation.o $O/test_main.o $O/helpers.o $O/ref_triangulation.o $O/ref_camera.o $O/ref_tracker_util.o \
    $O/cuda_ekf.o $OUT/obj_ekf/parameters.o $OUT/obj_ekf/odo_util.o $OUT/obj_ekf/timer.o $OUT/obj_ekf/util_util.o $OUT/obj_ekf/parameter_parser.o $OCVOBJ \
    -Wl,--gc-sections -L"$ROOT/hybvio_b200" -lhybvio_b200 -Wl,-rpath,'$ORIGIN/../../hybvio_b200' -lpthread -ldl -lz
echo built $OUT/run_ref_triangulation_tests

# Third binary: the per-track measurement model through the reference's interfaces, host path (reference triangulation.cpp,
# unmodified) against the device path (hybvio_b200/host/cuda_track_model.hpp), both on CudaEKF.
g++ $FL $INC2 -I"$ROOT/hybvio_b200/host" -c "$HERE/track_model_iface_test.cpp" -o $O/track_model_iface_test.o
g++ -o "$OUT/run_track_model_iface_test" $O/track_model_iface_test.o $O/ref_triangulation.o $O/ref_camera.o $O/ref_tracker_util.o \
    $O/cuda_ekf.o $OUT/obj_ekf/parameters.o $OUT/obj_ekf/odo_util.o $OUT/obj_ekf/timer.o $OUT/obj_ekf/util_util.o $OUT/obj_ekf/parameter_parser.o $OCVOBJ \
    -Wl,--gc-sections -L"$ROOT/hybvio_b200" -lhybvio_b200 -Wl,-rpath,'$ORIGIN/../../hybvio_b200' -lpthread -ldl -lz
echo built $OUT/run_track_model_iface_test
