#!/bin/bash
set -u
mkdir -p gpurun_out
for model in ${MODELS:-eqtransformer phasenet}; do
for chunk in ${CHUNKS:-512 1024 2048 0}; do
timeout 200 python bench.py --steps 5 --warmup 3 --model $model --no-cpu-baseline --chunk $chunk > gpurun_out/bench_c.log 2>gpurun_out/bench_c.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c.log").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$model chunk $chunk", "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), k)
except Exception as e:
    print("$model $chunk parse failed", e); print(open("gpurun_out/bench_c.err").read()[-500:])
PY
done
done
