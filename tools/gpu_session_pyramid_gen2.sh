#!/bin/bash
# GPU session for the second-generation pyramid kernel. Run under gpurun from the repo root:
#   gpurun --timeout 900 -- 'bash tools/gpu_session_pyramid_gen2.sh'
# Everything lands in gpurun_out/ (scratch; copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== parity: every pyramid + LK GPU test with HV_PYR_V2=1"
timeout 600 python -m pytest tests/test_zzz_gpu_pyramid_gen2.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pyr2_tests.log
echo "== timing A/B (CUDA events, 2 images and 32 images per launch)"
timeout 200 python tests/tools/pyr_time.py 2>&1 | tail -2 | tee gpurun_out/pyr_time_gen1.txt
HV_PYR_V2=1 timeout 200 python tests/tools/pyr_time.py 2>&1 | tail -2 | tee gpurun_out/pyr_time_gen2.txt
echo "== ncu --set full of both kernels (one repetition of tools/prof_kernels.py each)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hv_pyr_fused -c 4 -o gpurun_out/pyr_gen1_full -f python tools/prof_kernels.py 1 > gpurun_out/pyr_prof.log 2>&1
HV_PYR_V2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:hv_pyr_fused -c 4 -o gpurun_out/pyr_gen2_full -f python tools/prof_kernels.py 1 >> gpurun_out/pyr_prof.log 2>&1
echo "== bench with the switch set (the default run attaches the same A/B as pyramid_gen2_ab)"
HV_PYR_V2=1 HV_BENCH_NO_EXTRAS=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_pyr2.json 2> gpurun_out/bench_pyr2.err; tail -c 800 gpurun_out/bench_pyr2.json
