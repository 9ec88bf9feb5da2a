#!/bin/bash
# Round 2, GPU session S: the dependent chain of a frame on ONE stream (optical flow launched on the filter's stream, covariance part of the
# IMU burst on a stream of the library): whole GPU suite, bench config 2 (+ 8 sessions), launch list.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/s_gpu_tests.log
timeout 900 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; tail -c 300 gpurun_out/s_bench.err
HV_BENCH_NO_EXTRAS=1 timeout 400 python bench.py --sessions 8 --no-cpu-baseline --e2e-steps 50 > gpurun_out/s_bench_8s.json 2> gpurun_out/s_bench_8s.err; tail -c 200 gpurun_out/s_bench_8s.err
python - <<'PY'
import json
for n in ("s_bench", "s_bench_8s"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "python_harness", (d.get("python_harness") or {}).get("value"), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|ekf_' -c 800 --csv --log-file gpurun_out/s_launches.csv \
  python bench.py --steps 10 --warmup 3 --step-only > gpurun_out/s_launches_bench.log 2>&1; tail -c 200 gpurun_out/s_launches_bench.log
