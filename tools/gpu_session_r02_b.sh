#!/bin/bash
# Round 2, GPU session B: everything written since session A (corner detector, TMA staging of the pyramid kernel, one-sync hv_ekf_run_host,
# bench --config 4 / 1, e2e_adapter), phase timers, A/Bs, ncu of the TMA pyramid kernel.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session_r02_b.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 1. GPU suite (incl. pipeline parity with the device corner detector)"
timeout 1800 python -m pytest tests -q -m gpu -x -s -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/b_gpu_tests.log
echo "==== 2. phase timers of the update kernel"
HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/b_ekf_phases.txt 2>&1; tail -12 gpurun_out/b_ekf_phases.txt
echo "==== 3. bench: default (config 2), configs 4 and 1, A/Bs"
timeout 900 python bench.py > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; tail -c 600 gpurun_out/b_bench.err
timeout 600 python bench.py --config 4 > gpurun_out/b_bench_config4.json 2> gpurun_out/b_bench_config4.err
timeout 600 python bench.py --config 1 > gpurun_out/b_bench_config1.json 2> gpurun_out/b_bench_config1.err
HV_BENCH_NO_EXTRAS=1 HV_PYR_NO_TMA=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/b_bench_notma.json 2> gpurun_out/b_bench_notma.err
HV_BENCH_NO_EXTRAS=1 HV_NO_POLL=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/b_bench_nopoll.json 2> gpurun_out/b_bench_nopoll.err
timeout 400 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/b_bench_reference.json 2> gpurun_out/b_bench_reference.err
python - <<'PY'
import json
for n in ("b_bench", "b_bench_config4", "b_bench_config1", "b_bench_notma", "b_bench_nopoll", "b_bench_reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "launches/step", d.get("gpu_launches_per_step"),
              "| pyr", [v["us_per_launch"] for q, v in k.items() if "pyr" in q], "lk", [v["us_per_launch"] for q, v in k.items() if "lk" in q],
              "| batched", [(q[:12], v["us_per_launch"], v.get("frac_of_hbm_peak")) for q, v in (d.get("kernels_batched") or {}).items()],
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
echo "==== 4. ncu: pyramid with TMA staging / without, corner detector, launch list"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'hv_pyr' -s 4 -c 2 -o gpurun_out/b_pyr_tma_full -f python tools/prof_kernels.py 2 > gpurun_out/b_prof_pyr.log 2>&1
HV_PYR_NO_TMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'hv_pyr' -s 4 -c 2 -o gpurun_out/b_pyr_ldg_full -f python tools/prof_kernels.py 2 >> gpurun_out/b_prof_pyr.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/b_launches.csv \
    python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/b_launches_bench.log 2>&1
ls -la gpurun_out | grep " b_"
