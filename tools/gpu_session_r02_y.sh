#!/bin/bash
# Round 2, GPU session Y: everything once more on the final tree.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/y_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/y_bench_reference.json 2> gpurun_out/y_bench_reference.err; tail -c 200 gpurun_out/y_bench_reference.err
timeout 900 python bench.py > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err; tail -c 300 gpurun_out/y_bench.err
timeout 900 python bench.py --config 4 > gpurun_out/y_bench_config4.json 2> gpurun_out/y_bench_config4.err; tail -c 300 gpurun_out/y_bench_config4.err
timeout 900 python bench.py --config 1 > gpurun_out/y_bench_config1.json 2> gpurun_out/y_bench_config1.err; tail -c 300 gpurun_out/y_bench_config1.err
for ns in 8 16; do HV_BENCH_NO_EXTRAS=1 timeout 400 python bench.py --sessions $ns --no-cpu-baseline --e2e-steps 50 > gpurun_out/y_bench_${ns}s.json 2> gpurun_out/y_bench_${ns}s.err; done
python - <<'PY'
import json
for n in ("y_bench_reference", "y_bench", "y_bench_config4", "y_bench_config1", "y_bench_8s", "y_bench_16s"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "value", d["value"], "ms/step", d.get("ms_per_step"), "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "chain", (d.get("e2e_chain") or {}).get("value"),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("e2e_adapter") or {}).get("value"), "| clocks", (d.get("clocks") or {}).get("reasons"))
        if n == "y_bench":
            for q, v in (d.get("kernels") or {}).items(): print("   ", q[:90], v.get("us_per_launch"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:200])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hy_|ekf_' -c 800 --csv --log-file gpurun_out/y_launches.csv \
  python bench.py --steps 10 --warmup 3 --step-only > gpurun_out/y_launches_bench.log 2>&1; tail -c 200 gpurun_out/y_launches_bench.log
