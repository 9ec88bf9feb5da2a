#!/bin/bash
# Round 2, GPU session A: pipeline-level parity first, then the validated suite, the A/Bs that are still open (persistent update
# sequence, cluster of 16, chain variants) with phase timers, and the ncu captures of the default tracker kernels.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session_r02_a.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/a_gpu.txt
echo "==== 1. pipeline-level parity (lock-step + free-running), configs 2 / 4 / 1"
timeout 1500 python -m pytest tests/test_pipeline.py -q -m gpu -s -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/a_pipeline_tests.log
for m in cuda ref; do timeout 300 oracle/_ref/run_pipeline --mode $m --config 2 --frames 300 --out gpurun_out/pipeline_${m}_config2.json > /dev/null 2> gpurun_out/a_pipeline_$m.err; done
echo "==== 2. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider --deselect tests/test_pipeline.py 2>&1 | tail -15 | tee gpurun_out/a_gpu_tests.log
echo "==== 3. phase timers of the update kernel: default, cluster of 16"
HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/a_ekf_phases_c8.txt 2>&1; tail -12 gpurun_out/a_ekf_phases_c8.txt
HV_EKF_CLUSTER=16 HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/a_ekf_phases_c16.txt 2>&1; tail -12 gpurun_out/a_ekf_phases_c16.txt
echo "==== 4. bench: default, persistent sequence, cluster 16"
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -c 2500 gpurun_out/a_bench.json
HV_BENCH_NO_EXTRAS=1 HV_EKF_PERSIST=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/a_bench_persist.json 2> gpurun_out/a_bench_persist.err
HV_BENCH_NO_EXTRAS=1 HV_EKF_CLUSTER=16 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/a_bench_c16.json 2> gpurun_out/a_bench_c16.err
python - <<'EOF'
import json
for n in ("a_bench", "a_bench_persist", "a_bench_c16"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "launches/step", d.get("gpu_launches_per_step"))
    except Exception as ex:
        print(n, "failed", ex)
EOF
echo "==== 5. chain variants (N1)"
timeout 200 python tests/tools/track_model_bench.py > gpurun_out/a_tm_bench.json 2> gpurun_out/a_tm_bench.err
HV_CHAIN_SEPARATE=1 timeout 200 python tests/tools/track_model_bench.py > gpurun_out/a_tm_bench_separate.json 2>/dev/null
HV_CHAIN_PERSIST=1 timeout 200 python tests/tools/track_model_bench.py > gpurun_out/a_tm_bench_persist.json 2>/dev/null
python - <<'EOF'
import json
for n in ("a_tm_bench", "a_tm_bench_separate", "a_tm_bench_persist"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, json.dumps({k: v for k, v in d.items() if k in ("kernel", "chain", "loop", "variant")})[:900])
    except Exception as ex:
        print(n, "failed", ex)
EOF
echo "==== 6. ncu: launch list of the bench step, full captures of pyramid / LK / update / track model"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/a_launches_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'hv_pyr|hv_lk' -s 4 -c 4 -o gpurun_out/a_tracker_full -f python tools/prof_kernels.py 2 > gpurun_out/a_prof_tracker.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'ekf_update_cluster2' -s 8 -c 9 -o gpurun_out/a_update_full -f python tools/prof_kernels.py 2 > gpurun_out/a_prof_update.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'hv_track_model' -c 3 -o gpurun_out/a_track_model_full -f python tools/prof_track_model.py 1 > gpurun_out/a_prof_tm.log 2>&1
ls -la gpurun_out | tail -40
