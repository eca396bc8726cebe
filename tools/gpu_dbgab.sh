#!/bin/bash
# A/B of VP_DECB_DBG values on the EQT bench (only timing-neutral bits give right results): DBGS="0 32"
set -u
mkdir -p gpurun_out
for dbg in ${DBGS:-0 32}; do
VP_DECB_DBG=$dbg timeout 200 python bench.py --steps 4 --warmup 3 --model eqtransformer --no-cpu-baseline > gpurun_out/bench_dbg.log 2>gpurun_out/bench_dbg.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_dbg.log").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("dbg $dbg value", round(d["value"],2), "decb", k.get("decb"))
except Exception as e:
    print("dbg $dbg parse failed", e); print(open("gpurun_out/bench_dbg.err").read()[-600:])
PY
done
