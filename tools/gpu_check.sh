#!/bin/bash
# One GPU-box pass: parity tests, smoke, a short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --precision ${BENCH_PREC:-f16x3} --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3} > gpurun_out/bench.log 2>&1; echo "bench exit: $?" >> gpurun_out/bench.log
tail -5 gpurun_out/bench.log
