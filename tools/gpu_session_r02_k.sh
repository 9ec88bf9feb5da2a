#!/bin/bash
# Round 2, GPU session K: how much of the update kernel is instruction fetch (the body several times inside one launch); ncu --set full of
# the update / augmentation / predict kernels as they are now (bulk-copy staging).
set -u
mkdir -p gpurun_out
timeout 120 tools/ubench_update_twice 2>&1 | tee gpurun_out/k_update_repeated.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ekf_' -c 14 -o gpurun_out/k_ekf_full -f python tools/prof_kernels.py 1 > gpurun_out/k_prof_ekf.log 2>&1; tail -2 gpurun_out/k_prof_ekf.log
