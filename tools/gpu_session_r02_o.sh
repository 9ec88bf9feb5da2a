#!/bin/bash
# Round 2, GPU session O (2 GPUs): bench.py under torchrun as the driver launches it, both arms; the throughput-mode test.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_ekf.py -q -m gpu -x -p no:cacheprovider -k "throughput or device_op_list" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 60 --warmup 5 > gpurun_out/o_bench_2gpu_reference.json 2> gpurun_out/o_bench_2gpu_reference.err; tail -c 300 gpurun_out/o_bench_2gpu_reference.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 400 --warmup 20 > gpurun_out/o_bench_2gpu.json 2> gpurun_out/o_bench_2gpu.err; tail -c 300 gpurun_out/o_bench_2gpu.err
python - <<'PY'
import json
for n in ("o_bench_2gpu_reference", "o_bench_2gpu"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "n_gpus", d.get("n_gpus"), "value", d["value"], "ms/step", d.get("ms_per_step"), "e2e", d["e2e"]["value"], "adapter", (d.get("e2e_adapter") or {}).get("value"), "| clocks", d.get("clocks"))
    except Exception as ex:
        print(n, "failed", repr(ex)[:300])
PY
