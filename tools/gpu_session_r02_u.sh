#!/bin/bash
# Round 2, GPU session U: predicted end points read where the predictor left them (no copy into the result buffer): GPU suite, bench.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/u_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; tail -c 300 gpurun_out/u_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/u_bench.json") if l.startswith("{")][-1])
print("u_bench value", d["value"], "ms/step", d["ms_per_step"], "launches/step", d.get("gpu_launches_per_step"), "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|ekf_' -c 800 --csv --log-file gpurun_out/u_launches.csv \
  python bench.py --steps 10 --warmup 3 --step-only > gpurun_out/u_launches_bench.log 2>&1; tail -c 200 gpurun_out/u_launches_bench.log
