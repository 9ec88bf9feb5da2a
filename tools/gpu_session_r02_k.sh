#!/bin/bash
# Round 2, GPU session K: how much of the update kernel is instruction fetch (the body several times inside one launch).
set -u
mkdir -p gpurun_out
timeout 120 tools/ubench_update_twice 2>&1 | tee gpurun_out/k_update_repeated.txt
