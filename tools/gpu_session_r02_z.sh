#!/bin/bash
# Round 2, GPU session Z: launch list of the default bench command (own kernels, every phase of bench.py: value, python harness, e2e, adapter, kernel rows); bookkeeping test
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_ekf.py -q -m gpu -p no:cacheprovider -k "bookkeeping or run_device or device_op_list" 2>&1 | tail -3
HV_BENCH_NO_EXTRAS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|ekf_' -c 6000 --csv --log-file gpurun_out/z_launches_default_bench.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/z_launches_default_bench.log 2>&1; tail -c 300 gpurun_out/z_launches_default_bench.log | cut -c1-300
wc -l gpurun_out/z_launches_default_bench.csv
