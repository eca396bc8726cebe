#!/bin/bash
# Tensor-core path pass: layer tests, forward/annotate parity in both precisions, short benches.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcconv.py -q -x --timeout=120 -p no:cacheprovider > gpurun_out/pytest_tcconv.log 2>&1
echo "tcconv pytest exit: $?"; tail -3 gpurun_out/pytest_tcconv.log
timeout 400 python -m pytest tests/test_gpu_parity.py -q -s -k "tensor_core" --timeout=150 -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_tc.log
grep -E "max\||passed|failed|Error|error|assert" gpurun_out/pytest_tc.log | head -40
for prec in ${PRECS:-f16x3 bf16}; do
timeout 300 python bench.py --steps 3 --warmup 3 --precision $prec --no-cpu-baseline > gpurun_out/bench_$prec.log 2>&1; echo "bench $prec exit: $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$prec.log").read().strip().splitlines()[-1])
    print("$prec", "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), "fwd_ms", round(d["stages"]["forward_ms"],2), "TF/s", round(d["roofline"]["achieved"],1))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$prec.log").read()[-2000:])
PY
done
