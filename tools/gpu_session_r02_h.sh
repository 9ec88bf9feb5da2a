#!/bin/bash
# Round 2, GPU session H: mean part of the IMU burst as its own launch ahead of the tracker, programmatic dependent launch for the predict
# and LK kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 1. GPU tests"
timeout 2400 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/h_gpu_tests.log
echo "==== 2. bench: config 2 (+ without programmatic dependent launch)"
timeout 900 python bench.py > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; tail -c 300 gpurun_out/h_bench.err
HV_BENCH_NO_EXTRAS=1 HV_EKF_NO_PDL=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/h_bench_nopdl.json 2> gpurun_out/h_bench_nopdl.err
python - <<'PY'
import json
for n in ("h_bench", "h_bench_nopdl"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
        if n == "h_bench":
            for q, v in k.items(): print("   ", q[:90], v.get("us_per_launch"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
echo "==== 3. launch list of the step"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|ekf_' -c 800 --csv --log-file gpurun_out/h_launches.csv \
  python bench.py --steps 10 --warmup 3 --step-only > gpurun_out/h_launches_bench.log 2>&1; tail -c 200 gpurun_out/h_launches_bench.log
