#!/bin/bash
# Round 2, GPU session Q: tableau width = 4 (mod 16) (a) against the odd width of round 1 (b): dense products and elimination in isolation
set -u
mkdir -p gpurun_out
for v in a b; do echo "== variant $v (a: W = 4 mod 16, b: W odd)"; tools/ubench_gemm_$v; tools/ubench_elim2_w$v | cut -c1-200; done 2>&1 | tee gpurun_out/q_tableau_width.txt
