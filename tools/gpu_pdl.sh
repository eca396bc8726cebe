#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider -k "tensor_core or annotate_matches_oracle or golden or tcconv or bit_identical" > gpurun_out/pytest_pdl.log 2>&1
echo "pytest exit: $?"; tail -3 gpurun_out/pytest_pdl.log
for pdl in 1 0; do
for model in eqtransformer phasenet; do
VP_PDL=$pdl timeout 200 python bench.py --steps 6 --warmup 3 --model $model --no-cpu-baseline > gpurun_out/bench_pdl.log 2>gpurun_out/bench_pdl.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pdl.log").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("PDL=$pdl $model", "value", round(d["value"],2), "seq", round(d["sequential"]["value"],2), "e2e", round(d["e2e"]["value"],2), k)
except Exception as e:
    print("PDL=$pdl $model parse failed", e); print(open("gpurun_out/bench_pdl.err").read()[-600:])
PY
done
done
