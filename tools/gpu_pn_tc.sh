#!/bin/bash
# PhaseNet tensor-core path: parity tests (layer taps + probabilities + annotate), then short benches.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "pn_ or pick" --timeout=200 -p no:cacheprovider > gpurun_out/pytest_pn_tc.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_pn_tc.log
grep -E "max\||passed|failed|Error|error|assert|exit" gpurun_out/pytest_pn_tc.log | head -80
for prec in ${PRECS:-f16x3 bf16}; do
timeout 300 python bench.py --model phasenet --steps 3 --warmup 3 --precision $prec --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/bench_pn_$prec.log 2>&1; echo "bench $prec exit: $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pn_$prec.log").read().strip().splitlines()[-1])
    print("$prec", "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), "fwd_ms", round(d["stages"]["forward_ms"],2), {k: round(v["ms_per_step"],3) for k,v in d["kernels"]["per_class"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_pn_$prec.log").read()[-2000:])
PY
done
