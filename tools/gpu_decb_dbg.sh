#!/bin/bash
# Role elimination of the fused decoder tail: decb class time per station-day with parts of decb_kernel disabled
# (results are wrong with a non-zero mask; timing only).  1 no MMAs, 2 no layer epilogues, 4 no head, 8 no input loads.
set -u
mkdir -p gpurun_out
for mask in ${MASKS:-0 1 2 4 6 7 8 15}; do
VP_DECB_DBG=$mask timeout 300 python bench.py --steps 3 --warmup 2 --precision ${PREC:-f16x3} --no-cpu-baseline > gpurun_out/bench_decb_$mask.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_decb_$mask.log").read().strip().splitlines()[-1])
    print("VP_DECB_DBG=$mask", "decb_ms", round(d["kernels"]["per_class"]["decb"]["ms_per_step"],3), "value", round(d["value"],2))
except Exception as e:
    print("parse failed", e)
PY
done
