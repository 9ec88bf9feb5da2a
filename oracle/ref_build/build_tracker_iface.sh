#!/bin/bash
# TEST INFRASTRUCTURE: builds oracle/_ref/run_tracker_iface_test = oracle/ref_build/tracker_iface_test.cpp linked with
#   * the reference's own CPU back ends src/tracker/image_pyramid.cpp + optical_flow.cpp (UNMODIFIED) over the vendored
#     OpenCV objects of Makefile.lk and the accelerated-arrays CPU image sources,
#   * hybvio_b200/host/cuda_tracker_backends.cpp + libhybvio_b200.so (the CUDA back ends behind the same interfaces).
# Also the compile / link check of the tracker adapter (build() runs this where /root/reference exists).
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=${OUT:-$HERE/../_ref}
M=$REF/3rdparty/mobile-cv-suite
OCV=$M/opencv/modules
AA=$M/accelerated-arrays/src
[ -f "$OUT/gen/output/parameters.hpp" ] || "$HERE/build_ekf.sh"
[ -f "$OUT/obj_lk/video_lkpyramid.o" ] || make -C "$HERE" -f Makefile.lk -j8
mkdir -p "$OUT/obj_iface" "$OUT/inc" "$OUT/inc/fake/tracker" "$OUT/inc/fake/odometry"
ln -sfn "$AA" "$OUT/inc/accelerated-arrays"
ln -sfn "$M/jsonl-recorder" "$OUT/inc/jsonl-recorder"
# "../odometry/parameters.hpp" (relative to src/tracker, a dangling symlink in the read-only tree) -> the generated header
ln -sfn "$OUT/gen/output/parameters.hpp" "$OUT/inc/fake/odometry/parameters.hpp"
FL="-std=c++17 -O1 -w -fPIC -DEIGEN_MPL2_ONLY -DEIGEN_DONT_PARALLELIZE"
INC="-I$OUT/gen/output -I$M/eigen -I$M/json/single_include -I$REF/src/odometry -I$REF/src/tracker -I$REF/src -I$OUT/inc -I$OUT/inc/fake/tracker \
  -I$HERE/stubs -I$OCV/core/include -I$OCV/imgproc/include -I$OCV/video/include -I$OCV/calib3d/include -I$OCV/features2d/include \
  -I$OCV/flann/include -I$OCV/highgui/include -I$OCV/imgcodecs/include -I$OCV/videoio/include -I$OCV/photo/include -I$OCV/dnn/include \
  -I$OCV/ml/include -I$OCV/objdetect/include -I$OCV/stitching/include -I$OCV/../include"
O=$OUT/obj_iface
g++ $FL $INC -c "$REF/src/tracker/image_pyramid.cpp" -o $O/ref_image_pyramid.o &
g++ $FL $INC -c "$REF/src/tracker/optical_flow.cpp" -o $O/ref_optical_flow.o &
g++ $FL $INC -c "$ROOT/hybvio_b200/host/cuda_tracker_backends.cpp" -o $O/cuda_tracker_backends.o &
g++ $FL $INC -c "$HERE/tracker_iface_test.cpp" -o $O/tracker_iface_test.o &
for f in cpu/image image future log_and_assert; do g++ $FL -I$AA -c "$AA/$f.cpp" -o $O/aa_$(echo $f | tr / _).o & done
wait
OCVOBJ=$(ls $OUT/obj_lk/*.o $OUT/obj_lk/core_utils/*.o | grep -v shim.o)
g++ -o "$OUT/run_tracker_iface_test" $O/tracker_iface_test.o $O/ref_image_pyramid.o $O/ref_optical_flow.o $O/cuda_tracker_backends.o $O/aa_*.o \
    $OUT/obj_ekf/parameters.o $OUT/obj_ekf/parameter_parser.o $OUT/obj_ekf/util_util.o $OCVOBJ \
    -Wl,--gc-sections -L"$ROOT/hybvio_b200" -lhybvio_b200 -Wl,-rpath,'$ORIGIN/../../hybvio_b200' -lpthread -ldl -lz
echo built $OUT/run_tracker_iface_test
