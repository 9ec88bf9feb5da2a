#!/bin/bash
# Round 2, GPU session W3: racecheck after the warp barriers; EKF tests; bench (do the barriers cost anything?)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool racecheck --print-limit 6 --error-exitcode 0 python -m pytest -q -x -p no:cacheprovider -m gpu tests/test_gpu_ekf.py -k "test_cuda_fused_check_update_matches_reference_golden or predicted_mean" > gpurun_out/w3_racecheck.log 2>&1
grep -v "Host Frame\|^$" gpurun_out/w3_racecheck.log | grep -v "^=========         " | tail -24 | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_ekf.py tests/test_pipeline.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/w3_bench.json 2> gpurun_out/w3_bench.err; tail -c 300 gpurun_out/w3_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/w3_bench.json") if l.startswith("{")][-1])
print("w3_bench value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"))
for q, v in (d.get("kernels") or {}).items():
    if "update" in q: print("   ", q[:90], v.get("us_per_launch"))
PY
