#!/bin/bash
# TEST INFRASTRUCTURE: builds oracle/_ref/libref_tri.so = the reference's own triangulation + prepareVisualUpdate
# (src/odometry/triangulation.cpp, unmodified) with its EKF and camera model, plus the C shim ref_tri_shim.cpp.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${OUT:-$HERE/../_ref}
M=$REF/3rdparty/mobile-cv-suite
[ -f "$OUT/obj_ekf/ekf.o" ] || "$HERE/build_ekf.sh"
mkdir -p "$OUT/obj_tri" "$OUT/inc/fake/tracker" "$OUT/inc/fake/odometry"
ln -sfn "$M/jsonl-recorder" "$OUT/inc/jsonl-recorder"
ln -sfn "$OUT/gen/output/parameters.hpp" "$OUT/inc/fake/odometry/parameters.hpp"
FL="-std=c++17 -O2 -w -fPIC -ffunction-sections -fdata-sections -DEIGEN_MPL2_ONLY -DEIGEN_DONT_PARALLELIZE"
INC="-I$OUT/gen/output -I$M/eigen -I$M/json/single_include -I$M/yaml-cpp/include -I$REF/src/odometry -I$REF/src -I$OUT/inc -I$OUT/inc/fake/tracker"
O=$OUT/obj_tri
cc() { [ "$2" -nt "$1" ] || g++ $FL $INC -c "$1" -o "$2"; }
cc $REF/src/odometry/triangulation.cpp $O/triangulation.o &
cc $REF/src/tracker/camera.cpp $O/camera.o &
g++ $FL $INC -c $HERE/ref_tri_shim.cpp -o $O/shim.o &
wait
g++ -shared -o $OUT/libref_tri.so $O/triangulation.o $O/camera.o $O/shim.o $OUT/obj_ekf/ekf.o $OUT/obj_ekf/parameters.o $OUT/obj_ekf/odo_util.o \
    $OUT/obj_ekf/timer.o $OUT/obj_ekf/util_util.o $OUT/obj_ekf/parameter_parser.o -Wl,--gc-sections -Wl,--version-script=$HERE/exports.map
echo built $OUT/libref_tri.so
