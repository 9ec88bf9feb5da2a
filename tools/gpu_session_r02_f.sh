#!/bin/bash
# Round 2, GPU session F: pyramid TMA staging with the box on a 16-byte boundary, elimination variants (tools/elim_variants),
# check batch + augmentation in one launch with the 8-pivot elimination, ncu captures.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "==== 0. TMA staging"
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
timeout 120 python /tmp/tma_probe.py 2>&1 | tail -3 | tee gpurun_out/f_tma_probe.log
if ! grep -q "TMA staging ok" gpurun_out/f_tma_probe.log; then echo "TMA staging FAILED: everything below runs with HV_PYR_NO_TMA=1"; export HV_PYR_NO_TMA=1; fi
echo "==== 1. elimination variants (0: production 8x8 rsqrt chain; 1: 8x8 division-free shuffles; 2: 8x8 division-free smem; 3: 16x16 shuffles; 4: 16x16 smem)"
for v in 0 1 2 3 4; do echo "-- variant $v"; timeout 120 tools/ubench_elim2_v$v 2>&1 | cut -c1-330; done | tee gpurun_out/f_ubench_elim_variants.txt
echo "==== 2. GPU tests"
timeout 2400 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/f_gpu_tests.log
echo "==== 3. bench"
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -c 300 gpurun_out/f_bench.err
HV_BENCH_NO_EXTRAS=1 HV_PYR_NO_TMA=1 timeout 300 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 50 > gpurun_out/f_bench_notma.json 2> gpurun_out/f_bench_notma.err
python - <<'PY'
import json
for n in ("f_bench", "f_bench_notma"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        k = d.get("kernels") or {}
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"))
        for q, v in k.items(): print("   ", q[:90], v.get("us_per_launch"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
echo "==== 4. ncu: launch list of the step (own kernels), pyramid with TMA"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|ekf_' -c 800 --csv --log-file gpurun_out/f_launches.csv \
  python bench.py --steps 10 --warmup 3 --step-only > gpurun_out/f_launches_bench.log 2>&1; tail -c 200 gpurun_out/f_launches_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'hv_pyr' -s 2 -c 2 -o gpurun_out/f_pyr_tma_full -f python tools/prof_kernels.py 2 > gpurun_out/f_prof_pyr.log 2>&1; tail -2 gpurun_out/f_prof_pyr.log
