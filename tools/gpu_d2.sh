#!/bin/bash
# decb2 iteration pass: EQT tensor-core parity tests (under a short timeout: a hang must not eat the box), then quick benches of both decoder-tail kernels.
set -u
mkdir -p gpurun_out
timeout ${T1:-240} python -m pytest tests/test_gpu_parity.py -q -x --timeout=100 -p no:cacheprovider -k "${KEXPR:-eqt_forward_tensor_core or forward_range}" > gpurun_out/pytest_d2.log 2>&1
echo "pytest exit: $?"; tail -${TAILN:-15} gpurun_out/pytest_d2.log
for env in ${ENVS:-VP_DECB_V1=0 VP_DECB_V1=1}; do
for prec in ${PRECS:-f16x3 bf16}; do
env $env timeout 200 python bench.py --steps ${STEPS:-4} --warmup 3 --model eqtransformer --precision $prec --no-cpu-baseline > gpurun_out/bench_d2_${prec}_${env}.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_d2_${prec}_${env}.log").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$env $prec", "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "roofline", round(d["roofline"]["frac"],3), k)
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_d2_${prec}_${env}.log").read()[-1500:])
PY
done
done
