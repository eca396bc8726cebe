#!/bin/bash
set -u
mkdir -p gpurun_out
for lanes in ${LANES:-1 2 3 4}; do
VP_LANES=$lanes timeout 200 python bench.py --steps 6 --warmup 3 --model ${MODEL:-eqtransformer} --no-cpu-baseline --quick > gpurun_out/bench_l.log 2>gpurun_out/bench_l.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_l.log").read().strip().splitlines()[-1])
    print("lanes $lanes value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2))
except Exception as e:
    print("lanes $lanes parse failed", e); print(open("gpurun_out/bench_l.err").read()[-600:])
PY
done
