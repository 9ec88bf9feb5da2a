#!/bin/bash
# Round 2, GPU session W2: racecheck again with the warp barrier in the in-place rows-solve, full hazard records for what remains
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 6 --error-exitcode 0 python -m pytest -q -x -p no:cacheprovider -m gpu tests/test_gpu_ekf.py -k "test_cuda_fused_check_update_matches_reference_golden or predicted_mean" > gpurun_out/w2_racecheck_full.log 2>&1
grep -v "Host Frame\|^$" gpurun_out/w2_racecheck_full.log | grep -v "^=========         " | tail -60 | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_ekf.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
