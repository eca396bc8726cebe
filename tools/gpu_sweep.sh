#!/bin/bash
# Sweep a bench.py option over values: VAR=--chunk VALS="512 1024 2048" bash tools/gpu_sweep.sh
set -u
mkdir -p gpurun_out
for v in ${VALS}; do
timeout 300 python bench.py --steps 3 --warmup 2 --precision ${PREC:-f16x3} --no-cpu-baseline ${VAR} $v ${EXTRA:-} > gpurun_out/bench_sweep_$v.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_sweep_$v.log").read().strip().splitlines()[-1])
    print("${VAR} $v", "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), "fwd_ms", round(d["stages"]["forward_ms"],2))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_sweep_$v.log").read()[-1500:])
PY
done
