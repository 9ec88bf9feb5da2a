#!/bin/bash
# Round 2, GPU session N2: 8 / 16 sessions per GPU with 32 hardware queues and the outlier checks back on the main stream in throughput mode
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for ns in 8 16; do
  HV_BENCH_NO_EXTRAS=1 timeout 400 python bench.py --sessions $ns --no-cpu-baseline --e2e-steps 50 > gpurun_out/n2_bench_${ns}s.json 2> gpurun_out/n2_bench_${ns}s.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/n2_bench_${ns}s.json") if l.startswith("{")][-1])
print("$ns sessions: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
