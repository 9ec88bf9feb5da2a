#!/bin/bash
# Round 2, GPU session M: the whole GPU suite and every bench line on the tree as it stands.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 1. GPU tests (all)"
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/m_gpu_tests.log
echo "==== 2. smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "==== 3. bench: reference arm, configs 2, 4, 1, 8 sessions"
timeout 600 python bench.py --impl reference > gpurun_out/m_bench_reference.json 2> gpurun_out/m_bench_reference.err; tail -c 200 gpurun_out/m_bench_reference.err
timeout 900 python bench.py > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; tail -c 300 gpurun_out/m_bench.err
timeout 900 python bench.py --config 4 > gpurun_out/m_bench_config4.json 2> gpurun_out/m_bench_config4.err; tail -c 300 gpurun_out/m_bench_config4.err
timeout 900 python bench.py --config 1 > gpurun_out/m_bench_config1.json 2> gpurun_out/m_bench_config1.err; tail -c 300 gpurun_out/m_bench_config1.err
HV_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --sessions 8 --no-cpu-baseline > gpurun_out/m_bench_8sessions.json 2> gpurun_out/m_bench_8sessions.err; tail -c 300 gpurun_out/m_bench_8sessions.err
python - <<'PY'
import json
for n in ("m_bench_reference", "m_bench", "m_bench_config4", "m_bench_config1", "m_bench_8sessions"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"), "| clocks", d.get("clocks"))
        if n == "m_bench":
            for q, v in k.items(): print("   ", q[:90], v.get("us_per_launch"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
echo "==== 4. launch list of the step; phase timers"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|ekf_' -c 800 --csv --log-file gpurun_out/m_launches.csv \
  python bench.py --steps 10 --warmup 3 --step-only > gpurun_out/m_launches_bench.log 2>&1; tail -c 200 gpurun_out/m_launches_bench.log
HV_EKF_NO_PDL=1 HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/m_ekf_phases.txt 2>&1; tail -3 gpurun_out/m_ekf_phases.txt
