#!/bin/bash
# Round 2, GPU session E: where the tensor map has to live for the bulk tensor copy to run (tools/tma_probe), the 16-pivot
# division-free elimination (tools/ubench_elim2 + EKF tests + bench), check batch and augmentation in one launch.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 0. tensor-map placement probe"
for args in "0 0 0" "0 60 40" "0 700 440" "3 60 40" "1 60 40" "2 60 40" "0 60 40 128 108" "0 64 40 112 108"; do
  timeout 60 tools/tma_probe $args 2>&1 | tail -1
done | tee gpurun_out/e_tma_probe.log
echo "==== 1. elimination microbenchmark"
timeout 120 tools/ubench_elim2 2>&1 | tee gpurun_out/e_ubench_elim2.txt
export HV_PYR_NO_TMA=1
echo "==== 2. GPU tests: EKF, pipeline"
timeout 1500 python -m pytest tests/test_gpu_ekf.py tests/test_pipeline.py tests/test_gpu_track_model.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/e_gpu_tests.log
echo "==== 3. bench"
timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; tail -c 300 gpurun_out/e_bench.err
python - <<'PY'
import json
for n in ("e_bench",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"))
        for q, v in k.items(): print("   ", q[:90], v.get("us_per_launch"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
echo "==== 4. phase timers"
HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/e_ekf_phases.txt 2>&1; tail -12 gpurun_out/e_ekf_phases.txt
echo "==== 5. launch list of the bench step (own kernels only)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|ekf_' -c 800 --csv --log-file gpurun_out/e_launches.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/e_launches_bench.log 2>&1; tail -c 200 gpurun_out/e_launches_bench.log
