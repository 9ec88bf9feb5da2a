#!/bin/bash
# Round 2, GPU session O4 (4 GPUs): our arm under torchrun as the driver launches it
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 300 --warmup 20 > gpurun_out/o4_bench_4gpu.json 2> gpurun_out/o4_bench_4gpu.err; tail -c 400 gpurun_out/o4_bench_4gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/o4_bench_4gpu.json") if l.startswith("{")][-1])
    print("4 GPUs: value", d["value"], "ms/step", d.get("ms_per_step"), "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "numa", d["config"].get("host_numa_node_pinned"), "| clocks", (d.get("clocks") or {}).get("reasons"))
except Exception as ex:
    print("failed", repr(ex)[:300])
PY
