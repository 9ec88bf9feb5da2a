#!/bin/bash
# First GPU session after the track-model row was written (end of section 8). Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_session_track_model.sh'
# Everything lands in gpurun_out/ (scratch; copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export HV_GPU_FIRST_RUN_STRICT=1      # plain pass / fail for the tests that have not run on hardware yet (tests/conftest.py)
echo "== new GPU tests"; timeout 900 python -m pytest tests/test_zz_gpu_track_model.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/tm_tests.log
echo "== measurement tool"; timeout 300 python tests/tools/track_model_bench.py > gpurun_out/tm_bench.json 2> gpurun_out/tm_bench.err; tail -c 3000 gpurun_out/tm_bench.json
echo "== A/B: separate check / update launches, no PDL"
HV_CHAIN_SEPARATE=1 timeout 300 python tests/tools/track_model_bench.py > gpurun_out/tm_bench_separate.json 2>/dev/null
HV_EKF_NO_PDL=1 timeout 300 python tests/tools/track_model_bench.py > gpurun_out/tm_bench_nopdl.json 2>/dev/null
HV_CHAIN_PERSIST=1 timeout 300 python tests/tools/track_model_bench.py > gpurun_out/tm_bench_persist.json 2>/dev/null
echo "== launch list of the model kernel and one chain"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hv_track_model\|ekf_update_cluster2 -c 200 --csv --log-file gpurun_out/tm_launches.csv \
    python tools/prof_track_model.py 1 > gpurun_out/tm_prof.log 2>&1
echo "== ncu --set full of the same launches"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hv_track_model\|ekf_update_cluster2 -c 20 -o gpurun_out/tm_full -f \
    python tools/prof_track_model.py 1 >> gpurun_out/tm_prof.log 2>&1
echo "== all GPU tests, bench"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/all_tests.log
echo "== A/B: persistent sequence of updates"; HV_EKF_PERSIST=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_persist.json 2> gpurun_out/bench_persist.err; tail -c 600 gpurun_out/bench_persist.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
