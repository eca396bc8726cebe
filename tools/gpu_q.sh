#!/bin/bash
# quick pass: selected parity tests, then EQT + PhaseNet benches with per-class times
set -u
mkdir -p gpurun_out
timeout ${T1:-400} python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider -k "${KEXPR:-tensor_core or slice or annotate_matches_oracle or golden}" > gpurun_out/pytest_q.log 2>&1
echo "pytest exit: $?"; tail -3 gpurun_out/pytest_q.log
for model in ${MODELS:-eqtransformer phasenet}; do
for prec in ${PRECS:-f16x3}; do
timeout 200 python bench.py --steps ${STEPS:-4} --warmup 3 --model $model --precision $prec --no-cpu-baseline > gpurun_out/bench_q_${model}_$prec.log 2>gpurun_out/bench_q_${model}_$prec.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_q_${model}_$prec.log").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$model $prec", "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), k)
except Exception as e:
    print("$model bench parse failed", e); print(open("gpurun_out/bench_q_${model}_$prec.err").read()[-800:])
PY
done
done
