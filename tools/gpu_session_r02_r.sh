#!/bin/bash
# Round 2, GPU session R: tableau width = 4 (mod 16) in the kernel: whole GPU suite, bench config 2, phase stamps.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; tail -c 300 gpurun_out/r_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r_bench.json") if l.startswith("{")][-1])
k = d.get("kernels") or {}
print("r_bench value", d["value"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
for q, v in k.items(): print("   ", q[:90], v.get("us_per_launch"))
PY
HV_EKF_NO_PDL=1 HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/r_ekf_phases.txt 2>&1; tail -4 gpurun_out/r_ekf_phases.txt | cut -c1-400
