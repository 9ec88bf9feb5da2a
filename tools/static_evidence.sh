#!/bin/bash
# Static evidence for the opt-in tracker kernels (ptxas -v, SASS instruction counts, one row of the Scharr strip) from the objects `make`
# built: bash tools/static_evidence.sh > profiles/rNN_static_tracker_variants.txt
set -e
cd "$(dirname "$0")/.."
echo "# Static evidence for the opt-in tracker kernels (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3; no GPU involved)"
echo "# cuobjdump -sass build/obj/pyramid.o / lk.o, ptxas -v from build/obj/*.ptxas.log; regenerate with tools/static_evidence.sh"
echo; echo "## ptxas"
grep -A3 "fused2_kernel\|fused_kernel" build/obj/pyramid.ptxas.log | grep -v "^--\|Compile time"
grep -A3 "cta_kernelILi31ELi8\|cta_kernelILi31ELi4" build/obj/lk.ptxas.log | grep -v "^--\|Compile time"
echo; echo "## static SASS instruction counts per kernel"
cuobjdump -sass build/obj/pyramid.o | awk '/Function :/{f=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@!]/{c[f]++} END{for(k in c) print k, c[k]}' | sort
cuobjdump -sass build/obj/lk.o | awk '/Function :/{f=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@!]/{c[f]++} END{for(k in c) print k, c[k]}' | grep "cta_kernelILi31" | sort
echo; echo "## hv_pyr_fused2_kernel: one row (4 pixels: Ix, Iy as int16 pairs + the gray word) of the Scharr strip, between two 16-byte stores"
T=$(mktemp)
cuobjdump -sass build/obj/pyramid.o | awk '/Function : _Z20/{p=1} /Function : _Z19/{p=0} p' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/^ *\/\*[0-9a-f]*\*\/ *//; s/ *\/\*.*//' > "$T"
a=$(grep -n "ST.E.128" "$T" | sed -n 1p | cut -d: -f1); b=$(grep -n "ST.E.128" "$T" | sed -n 2p | cut -d: -f1)
echo "($((b-a)) instructions)"; sed -n "$((a+1)),${b}p" "$T"; rm -f "$T"
