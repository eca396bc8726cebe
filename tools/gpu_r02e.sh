#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=300 -p no:cacheprovider -k "tensor_core or forward_range or annotate_tensor or golden" > gpurun_out/pytest_decb.log 2>&1
echo "pytest exit: $?"; tail -2 gpurun_out/pytest_decb.log
for cfg in ${CFGS:-X=0}; do
for prec in ${PRECS:-f16x3}; do
env $cfg timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --precision $prec --model ${MODEL:-eqtransformer} > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_e.json").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$cfg $prec", "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), k)
except Exception as e:
    print("$cfg parse failed", e, open("gpurun_out/bench_e.err").read()[-600:])
PY
done
done
