#!/bin/bash
# Round 2, GPU session G: outlier checks of a device-resident list on the library's side stream, LK(k) after the IMU burst of frame k.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 1. GPU tests"
timeout 2400 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/g_gpu_tests.log
echo "==== 2. bench: configs 2, 4, 1"
timeout 900 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; tail -c 300 gpurun_out/g_bench.err
timeout 900 python bench.py --config 4 > gpurun_out/g_bench_config4.json 2> gpurun_out/g_bench_config4.err; tail -c 300 gpurun_out/g_bench_config4.err
timeout 900 python bench.py --config 1 > gpurun_out/g_bench_config1.json 2> gpurun_out/g_bench_config1.err; tail -c 300 gpurun_out/g_bench_config1.err
python - <<'PY'
import json
for n in ("g_bench", "g_bench_config4", "g_bench_config1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
        if n == "g_bench":
            for q, v in k.items(): print("   ", q[:90], v.get("us_per_launch"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
echo "==== 3. launch list of the step"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|ekf_' -c 800 --csv --log-file gpurun_out/g_launches.csv \
  python bench.py --steps 10 --warmup 3 --step-only > gpurun_out/g_launches_bench.log 2>&1; tail -c 200 gpurun_out/g_launches_bench.log
