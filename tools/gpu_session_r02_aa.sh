#!/bin/bash
# Round 2, GPU session AA: the adapter refreshes its inertial mirror from the mean launch; the adapter-level driver reads the pose where the
# flow predictor does. Whole GPU suite (the reference's Catch2 suites and the pipeline runs go through the adapter), bench.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/aa_gpu_tests.log
timeout 900 python bench.py > gpurun_out/aa_bench.json 2> gpurun_out/aa_bench.err; tail -c 300 gpurun_out/aa_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/aa_bench_reference.json 2> gpurun_out/aa_bench_reference.err
python - <<'PY'
import json
for n in ("aa_bench", "aa_bench_reference"):
    d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
    print(n, "value", d["value"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
PY
