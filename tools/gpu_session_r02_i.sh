#!/bin/bash
# Round 2, GPU session I: where the augmentation spends its last 11 us (phase marks inside the Joseph / symmetrise region, PDL off so that
# the stamps of a launch do not include waiting for its predecessor), e2e with the IMU burst issued ahead of the optical flow.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 1. phase timers (HV_EKF_NO_PDL=1)"
HV_EKF_NO_PDL=1 HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/i_ekf_phases.txt 2>&1; tail -12 gpurun_out/i_ekf_phases.txt
echo "==== 2. EKF tests, bench"
timeout 900 python -m pytest tests/test_gpu_ekf.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; tail -c 300 gpurun_out/i_bench.err
python - <<'PY'
import json
for n in ("i_bench",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("host_phase_us_per_step"), "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
