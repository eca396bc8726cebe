#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --timeout=300 -p no:cacheprovider -k "slice or annotate_matches_oracle or edge_records or golden" > gpurun_out/pytest_k1.log 2>&1
echo "pytest exit: $?"; tail -2 gpurun_out/pytest_k1.log
for cfg in "VP_K1_TMA=0" "VP_K1_TMA=1"; do
for model in eqtransformer phasenet; do
env $cfg timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --model $model > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_j.json").read().strip().splitlines()[-1])
    print("$cfg $model slice_ms", round(d["stages"]["slice_ms"],3), "frac", round(d["stages"]["slice_hbm"]["frac"],3), "value", round(d["value"],2))
except Exception as e:
    print("$cfg parse failed", e, open("gpurun_out/bench_j.err").read()[-600:])
PY
done
done
for cfg in "VP_LANES=2" "VP_LANES=3" "VP_LANES=4"; do
for chunk in 0 2048 3072; do
env $cfg timeout 300 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --quick --chunk $chunk > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_j.json").read().strip().splitlines()[-1])
    print("$cfg chunk $chunk value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2))
except Exception as e:
    print("$cfg parse failed", e, open("gpurun_out/bench_j.err").read()[-600:])
PY
done
done
