#!/bin/bash
# First GPU call of round 2: everything that was written after the last GPU minutes of round 1, in priority order, each step under
# its own timeout (about 25 minutes of box time in total). From the repo root:
#   gpurun --timeout 2400 -- 'bash tools/gpu_session_round2.sh'
# Read gpurun_out/*.log / *.json afterwards; copy what should be judged into profiles/ (r02_*).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export HV_GPU_FIRST_RUN_STRICT=1
echo "==== 1. validated suite (must stay green)"
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_zz_gpu_track_model.py --deselect tests/test_zzz_gpu_persistent.py \
    --deselect tests/test_zzz_gpu_tracker_variants.py 2>&1 | tail -5 | tee gpurun_out/r2_validated_tests.log
echo "==== 2. opt-in tracker kernels (pyramid gen 2, 8-warp LK)"
bash tools/gpu_session_tracker_variants.sh 2>&1 | tee gpurun_out/r2_tracker_variants.log
echo "==== 3. track model + device-gated chain + persistent updates"
timeout 600 python -m pytest tests/test_zzz_gpu_persistent.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r2_persistent_tests.log
bash tools/gpu_session_track_model.sh 2>&1 | tee gpurun_out/r2_track_model.log
