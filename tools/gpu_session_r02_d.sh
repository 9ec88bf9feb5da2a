#!/bin/bash
# Round 2, GPU session D: why does the TMA staging fail (compute-sanitizer), LK prefetch A/B in one session, speculative update behind inlier checks.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 0. TMA staging: error text and compute-sanitizer"
cat > /tmp/tma_probe.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from hybvio_b200 import capi, synth
hv = capi.Context(0)
img, _ = synth.stereo_frame(1, 752, 480)
p = hv.pyramid(752, 480, 31, 3)
try:
    p.build(np.ascontiguousarray(img)); hv.sync()
    g, d = p.download(0)
    print("TMA staging ok, level 0 equals input:", bool((g == img).all()), "gradient checksum", int(d.astype(np.int64).sum()))
except Exception as ex:
    print("TMA staging failed:", ex)
PY
timeout 120 python /tmp/tma_probe.py 2>&1 | tail -3 | tee gpurun_out/d_tma_probe.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/tma_probe.py 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|in /" | head -40 | tee gpurun_out/d_tma_sanitizer.log
if ! grep -q "TMA staging ok" gpurun_out/d_tma_probe.log; then echo "TMA staging FAILED: everything below runs with HV_PYR_NO_TMA=1"; export HV_PYR_NO_TMA=1; fi
echo "==== 1. GPU tests: EKF, pipeline (speculative update behind inlier checks is on by default)"
timeout 1500 python -m pytest tests/test_gpu_ekf.py tests/test_pipeline.py tests/test_gpu_pyramid_lk.py tests/test_gpu_tracker_iface.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/d_gpu_tests.log
echo "==== 2. bench: default, LK without prefetch"
timeout 900 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; tail -c 300 gpurun_out/d_bench.err
HV_BENCH_NO_EXTRAS=1 HV_LK_NO_PREFETCH=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/d_bench_lk_noprefetch.json 2> gpurun_out/d_bench_lk_noprefetch.err
HV_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/d_bench_lk_prefetch.json 2> gpurun_out/d_bench_lk_prefetch.err
timeout 400 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/d_bench_reference.json 2> gpurun_out/d_bench_reference.err
python - <<'PY'
import json
for n in ("d_bench", "d_bench_lk_noprefetch", "d_bench_lk_prefetch", "d_bench_reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("host_phase_us_per_step"), "adapter", (d.get("e2e_adapter") or {}).get("value"),
              "| lk", [v["us_per_launch"] for q, v in k.items() if "lk" in q], "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
echo "==== 3. phase timers"
HV_LIB_PATH=hybvio_b200/libhybvio_b200_timing.so timeout 200 python tools/ekf_phases.py > gpurun_out/d_ekf_phases.txt 2>&1; tail -12 gpurun_out/d_ekf_phases.txt
