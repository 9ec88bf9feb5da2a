#!/bin/bash
# Round 2, GPU session T: randomised mix of the asynchronous paths against the oracle; bench configs 4 and 1 and the reference arm on the
# tree with the single-stream chain.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_ekf.py -q -m gpu -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/t_gpu_tests_ekf.log
timeout 600 python bench.py --impl reference > gpurun_out/t_bench_reference.json 2> gpurun_out/t_bench_reference.err; tail -c 200 gpurun_out/t_bench_reference.err
timeout 900 python bench.py --config 4 > gpurun_out/t_bench_config4.json 2> gpurun_out/t_bench_config4.err; tail -c 300 gpurun_out/t_bench_config4.err
timeout 900 python bench.py --config 1 > gpurun_out/t_bench_config1.json 2> gpurun_out/t_bench_config1.err; tail -c 300 gpurun_out/t_bench_config1.err
python - <<'PY'
import json
for n in ("t_bench_reference", "t_bench_config4", "t_bench_config1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "value", d["value"], "ms/step", d.get("ms_per_step"), "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
