#!/bin/bash
# Round 2, GPU session N: 8 sessions per GPU -- hardware queue aliasing? (CUDA_DEVICE_MAX_CONNECTIONS)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for conn in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$conn HV_BENCH_NO_EXTRAS=1 timeout 400 python bench.py --sessions 8 --no-cpu-baseline --e2e-steps 50 > gpurun_out/n_bench_8s_conn$conn.json 2> gpurun_out/n_bench_8s_conn$conn.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/n_bench_8s_conn$conn.json") if l.startswith("{")][-1])
print("connections $conn: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
CUDA_DEVICE_MAX_CONNECTIONS=32 HV_BENCH_NO_EXTRAS=1 timeout 400 python bench.py --sessions 16 --no-cpu-baseline --e2e-steps 50 > gpurun_out/n_bench_16s_conn32.json 2> gpurun_out/n_bench_16s_conn32.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/n_bench_16s_conn32.json") if l.startswith("{")][-1])
print("16 sessions, connections 32: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
