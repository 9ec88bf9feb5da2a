#!/bin/bash
# Round 2, GPU session W: compute-sanitizer (memcheck, racecheck) over the EKF paths that changed this round.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== memcheck: device op list, predicted mean, random mix (seed 0), pyramid with TMA + LK on a caller stream"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 --error-exitcode 0 python -m pytest -q -x -p no:cacheprovider -m gpu tests/test_gpu_ekf.py tests/test_gpu_pyramid_lk.py \
  -k "device_op_list_matches_oracle or predicted_mean or (random_mix and 0) or device_frame_and_device_lk" 2>&1 | grep -v "^=========     \|Host Frame\|^$" | tail -25 | tee gpurun_out/w_memcheck.log
echo "== racecheck: cluster update kernel (shared-memory hazards), one frame's list"
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 --error-exitcode 0 python -m pytest -q -x -p no:cacheprovider -m gpu tests/test_gpu_ekf.py -k "test_cuda_fused_check_update_matches_reference_golden or predicted_mean" 2>&1 | grep -v "^=========     \|Host Frame\|^$" | tail -25 | tee gpurun_out/w_racecheck.log
