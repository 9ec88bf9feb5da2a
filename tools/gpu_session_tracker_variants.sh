#!/bin/bash
# GPU session for the opt-in tracker kernels: second-generation pyramid, 8-warp LK. Run under gpurun from the repo root:
#   gpurun --timeout 900 -- 'bash tools/gpu_session_tracker_variants.sh'
# Everything lands in gpurun_out/ (scratch; copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export HV_GPU_FIRST_RUN_STRICT=1      # plain pass / fail for the tests that have not run on hardware yet (tests/conftest.py)
echo "== parity: every pyramid + LK GPU test with HV_PYR_V2=1"
timeout 900 python -m pytest tests/test_zzz_gpu_tracker_variants.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pyr2_tests.log
echo "== timing A/B (CUDA events, 2 images and 32 images per launch)"
timeout 200 python tests/tools/pyr_time.py 2>&1 | tail -2 | tee gpurun_out/pyr_time_gen1.txt
HV_PYR_V2=1 timeout 200 python tests/tools/pyr_time.py 2>&1 | tail -2 | tee gpurun_out/pyr_time_gen2.txt
echo "== ncu --set full of both kernels (one repetition of tools/prof_kernels.py each)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hv_pyr_fused -c 4 -o gpurun_out/pyr_gen1_full -f python tools/prof_kernels.py 1 > gpurun_out/pyr_prof.log 2>&1
HV_PYR_V2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:hv_pyr_fused -c 4 -o gpurun_out/pyr_gen2_full -f python tools/prof_kernels.py 1 >> gpurun_out/pyr_prof.log 2>&1
echo "== bench with the switch set (the default run attaches the same A/B as tracker_variants_ab)"
HV_PYR_V2=1 HV_BENCH_NO_EXTRAS=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_pyr2.json 2> gpurun_out/bench_pyr2.err; tail -c 800 gpurun_out/bench_pyr2.json
echo "== LK: 8 warps per feature (tools/lk_time.py), CTA kernel for the batched launch"
timeout 200 python tools/lk_time.py 2>&1 | tail -3 | tee gpurun_out/lk_time_default.txt
HV_LK_CTA_WARPS=8 timeout 200 python tools/lk_time.py 2>&1 | tail -3 | tee gpurun_out/lk_time_8warps.txt
# batched launches (8 sessions x 150 features): HV_LK_CTA_MAX=100000 python bench.py --sessions 8 --no-cpu-baseline  (kernels_batched row)
