#!/bin/bash
# One GPU-box pass: full parity suite, smoke, bench in the three precision modes, tc layer micro-benchmark.
set -u
mkdir -p gpurun_out
bash tools/gpu_check.sh
PRECS="f16x3 bf16" bash tools/gpu_tc.sh
B=1024 timeout 300 python tools/tc_microbench.py > gpurun_out/tc_micro.log 2>&1; tail -4 gpurun_out/tc_micro.log
