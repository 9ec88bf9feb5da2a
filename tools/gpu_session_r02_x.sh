#!/bin/bash
# Round 2, GPU session X2: the tree with the bulk S exchange for n >= 57 only and the one-pass symmetrisation
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/x2_gpu_tests.log
HV_EKF_NO_PDL=1 HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/x2_ekf_phases.txt 2>&1; tail -7 gpurun_out/x2_ekf_phases.txt | cut -c1-420
timeout 900 python bench.py > gpurun_out/x2_bench.json 2> gpurun_out/x2_bench.err; tail -c 300 gpurun_out/x2_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/x2_bench.json") if l.startswith("{")][-1])
print("x2_bench value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"))
for q, v in (d.get("kernels") or {}).items():
    if "ekf" in q: print("   ", q[:90], v.get("us_per_launch"))
PY
