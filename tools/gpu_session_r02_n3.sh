#!/bin/bash
# Round 2, GPU session N3: 16 sessions per GPU in throughput mode: hardware queues 8 / 32, pyramid staging with / without TMA
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {
  HV_BENCH_NO_EXTRAS=1 timeout 400 python bench.py --sessions 16 --no-cpu-baseline --e2e-steps 20 > gpurun_out/n3_$1.json 2> gpurun_out/n3_$1.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/n3_$1.json") if l.startswith("{")][-1])
print("$1: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
}
CUDA_DEVICE_MAX_CONNECTIONS=8 run conn8
CUDA_DEVICE_MAX_CONNECTIONS=16 run conn16
CUDA_DEVICE_MAX_CONNECTIONS=32 HV_PYR_NO_TMA=1 run conn32_notma
