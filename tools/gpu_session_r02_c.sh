#!/bin/bash
# Round 2, GPU session C: TMA pyramid staging (fixed instruction form), LK with cp.async region prefetch, corner detector, one-sync
# hv_ekf_run_host with a copy stream, bench for configs 2 / 4 / 1, A/Bs, ncu captures.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session_r02_c.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 0. does the TMA staging run?"
timeout 300 python -m pytest tests/test_gpu_pyramid_lk.py -q -m gpu -x -p no:cacheprovider -k "pyramid" 2>&1 | tail -5 | tee gpurun_out/c_tma_probe.log
if ! grep -q " passed" gpurun_out/c_tma_probe.log || grep -q "failed" gpurun_out/c_tma_probe.log; then echo "TMA staging FAILED: everything below runs with HV_PYR_NO_TMA=1"; export HV_PYR_NO_TMA=1; fi
echo "==== 1. GPU suite (incl. pipeline parity with the device corner detector)"
timeout 1800 python -m pytest tests -q -m gpu -s -p no:cacheprovider 2>&1 | tail -45 | tee gpurun_out/c_gpu_tests.log
echo "==== 2. bench: default (config 2), configs 4 and 1, A/Bs"
timeout 900 python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; tail -c 400 gpurun_out/c_bench.err
timeout 600 python bench.py --config 4 > gpurun_out/c_bench_config4.json 2> gpurun_out/c_bench_config4.err
timeout 600 python bench.py --config 1 > gpurun_out/c_bench_config1.json 2> gpurun_out/c_bench_config1.err
HV_BENCH_NO_EXTRAS=1 HV_PYR_NO_TMA=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/c_bench_notma.json 2> gpurun_out/c_bench_notma.err
HV_BENCH_NO_EXTRAS=1 HV_NO_POLL=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/c_bench_nopoll.json 2> gpurun_out/c_bench_nopoll.err
timeout 400 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/c_bench_reference.json 2> gpurun_out/c_bench_reference.err
python - <<'PY'
import json
for n in ("c_bench", "c_bench_config4", "c_bench_config1", "c_bench_notma", "c_bench_nopoll", "c_bench_reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("host_phase_us_per_step"), "adapter", (d.get("e2e_adapter") or {}).get("value"), "launches/step", d.get("gpu_launches_per_step"),
              "| pyr", [v["us_per_launch"] for q, v in k.items() if "pyr" in q], "lk", [v["us_per_launch"] for q, v in k.items() if "lk" in q],
              "| batched", [(q[:12], v["us_per_launch"], v.get("frac_of_hbm_peak")) for q, v in (d.get("kernels_batched") or {}).items()],
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
echo "==== 3. ncu: pyramid with TMA staging / without, LK, corner detector, launch list"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'hv_pyr|hv_lk' -s 4 -c 4 -o gpurun_out/c_tracker_full -f python tools/prof_kernels.py 2 > gpurun_out/c_prof_tracker.log 2>&1
HV_PYR_NO_TMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'hv_pyr' -s 2 -c 2 -o gpurun_out/c_pyr_ldg_full -f python tools/prof_kernels.py 2 >> gpurun_out/c_prof_tracker.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'hv_gftt' -c 2 -o gpurun_out/c_gftt_full -f python tools/prof_gftt.py > gpurun_out/c_prof_gftt.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c_launches.csv \
    python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/c_launches_bench.log 2>&1
echo "==== 4. phase timers"
HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/c_ekf_phases.txt 2>&1; tail -12 gpurun_out/c_ekf_phases.txt
ls -la gpurun_out | grep " c_"
