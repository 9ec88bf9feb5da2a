#!/bin/bash
# Round 2, GPU session L: augmentation with pushed (instead of fetched) exchanges; issue / wait split of the host-buffer op list.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 1. GPU tests: EKF, pipeline"
timeout 1500 python -m pytest tests/test_gpu_ekf.py tests/test_pipeline.py tests/test_gpu_track_model.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/l_gpu_tests.log
echo "==== 2. phase timers (HV_EKF_NO_PDL=1)"
HV_EKF_NO_PDL=1 HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/l_ekf_phases.txt 2>&1; tail -3 gpurun_out/l_ekf_phases.txt
echo "==== 3. bench"
timeout 900 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; tail -c 300 gpurun_out/l_bench.err
python - <<'PY'
import json
for n in ("l_bench",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("host_phase_us_per_step"), "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
        for q, v in k.items(): print("   ", q[:90], v.get("us_per_launch"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
