#!/bin/bash
# TEST INFRASTRUCTURE: builds oracle/_ref/libref_ekf.so = the reference's own odometry::EKF (src/odometry/ekf.cpp,
# unmodified, vendored Eigen 3.3.90) compiled straight from /root/reference with g++ (no cmake), plus a C shim.
# The reference's parameter struct is produced by its own codegen script (inputs symlinked, nothing copied).
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${OUT:-$HERE/../_ref}
M=$REF/3rdparty/mobile-cv-suite
GEN=$OUT/gen
OBJ=$OUT/obj_ekf
mkdir -p "$GEN" "$OBJ"; ln -sfn "$REF/src" "$OUT/src"
for f in parameters_base.hpp parameters_base.cpp parameter_definitions.c; do ln -sf "$REF/codegen/$f" "$GEN/$f"; done
(cd "$GEN" && python3 "$REF/codegen/generate_parameters.py" > /dev/null)
# include root that makes  #include "parameters.hpp" (a dangling symlink in the read-only tree) and
# "../codegen/output/..." style paths resolve to the generated files
FL="-std=c++17 -O2 -w -fPIC -ffunction-sections -fdata-sections -DEIGEN_MPL2_ONLY -DEIGEN_DONT_PARALLELIZE"
INC="-I$GEN/output -I$M/eigen -I$M/json/single_include -I$M/yaml-cpp/include -I$REF/src/odometry -I$REF/src"
cc() { [ "$2" -nt "$1" ] || g++ $FL $INC -c "$1" -o "$2"; }
cc $REF/src/odometry/ekf.cpp $OBJ/ekf.o &
cc $GEN/output/parameters.cpp $OBJ/parameters.o &
cc $REF/src/odometry/util.cpp $OBJ/odo_util.o &
cc $REF/src/util/timer.cpp $OBJ/timer.o &
cc $REF/src/util/util.cpp $OBJ/util_util.o &
cc $REF/src/util/parameter_parser.cpp $OBJ/parameter_parser.o &
g++ $FL $INC -I$REF/src/odometry -c $HERE/ref_ekf_shim.cpp -o $OBJ/shim.o &
wait
for o in ekf parameters odo_util timer util_util parameter_parser shim; do [ -f $OBJ/$o.o ] || { echo "missing $o.o"; exit 1; }; done
g++ -shared -o $OUT/libref_ekf.so $OBJ/*.o -Wl,--gc-sections -Wl,--version-script=$HERE/exports.map -Wl,-z,defs
echo built $OUT/libref_ekf.so
