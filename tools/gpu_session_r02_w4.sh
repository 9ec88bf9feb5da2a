#!/bin/bash
# Round 2, GPU session W4: racecheck over the tracker kernels (pyramid with TMA staging, LK with cp.async), the corner detector and the track model
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool racecheck --print-limit 6 --error-exitcode 0 python -m pytest -q -x -p no:cacheprovider -m gpu tests/test_gpu_pyramid_lk.py tests/test_gpu_gftt.py \
  -k "device_frame_and_device_lk or gftt" > gpurun_out/w4_racecheck_tracker.log 2>&1
grep -v "Host Frame\|^$" gpurun_out/w4_racecheck_tracker.log | grep -v "^=========         " | tail -16 | cut -c1-260
timeout 1500 compute-sanitizer --tool racecheck --print-limit 6 --error-exitcode 0 python -m pytest -q -x -p no:cacheprovider -m gpu tests/test_gpu_track_model.py -k "not chain and not suite" > gpurun_out/w4_racecheck_tm.log 2>&1
grep -v "Host Frame\|^$" gpurun_out/w4_racecheck_tm.log | grep -v "^=========         " | tail -12 | cut -c1-260
