#!/bin/bash
# Quick iteration pass: tensor-core layer + forward parity tests, then short benches with per-kernel-class times.
# ENVS: space-separated list of "NAME=VALUE" toggles to compare (each is run for both models), e.g. "X=0 VP_TC_COAL=0".
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcconv.py -q -x --timeout=120 -p no:cacheprovider > gpurun_out/pytest_tcconv.log 2>&1
echo "tcconv pytest exit: $?"; tail -3 gpurun_out/pytest_tcconv.log
timeout 500 python -m pytest tests/test_gpu_parity.py -q -k "${KEXPR:-tensor_core or annotate}" --timeout=150 -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "parity pytest exit: $?"; tail -5 gpurun_out/pytest_tc.log
for env in ${ENVS:-X=0}; do
for model in ${MODELS:-eqtransformer phasenet}; do
env $env timeout 300 python bench.py --steps ${STEPS:-4} --warmup 3 --model $model --precision ${PREC:-f16x3} --no-cpu-baseline > gpurun_out/bench_iter_${model}_${env}.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_iter_${model}_${env}.log").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$env $model", "value", round(d["value"],2), "seq", round(d["sequential"]["value"],2), "e2e", round(d["e2e"]["value"],2), k)
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_iter_${model}_${env}.log").read()[-1500:])
PY
done
done
