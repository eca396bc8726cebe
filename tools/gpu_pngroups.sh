#!/bin/bash
# A/B of the N-group counts of PhaseNet's three wide layers (VP_PN_GROUPS="dd,us,ut"): parity test + bench per setting
set -u
mkdir -p gpurun_out
for G in ${GROUPS_LIST:-4,4,4 2,2,2}; do
  export VP_PN_GROUPS=$G
  timeout 300 python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider -k "pn_forward or phasenet" > gpurun_out/pytest_png.log 2>&1; echo "groups $G pytest exit $?: $(tail -1 gpurun_out/pytest_png.log)"
  timeout 200 python bench.py --steps 6 --warmup 3 --model phasenet --no-cpu-baseline > gpurun_out/bench_png.log 2> gpurun_out/bench_png.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_png.log").read().strip().splitlines()[-1])
    print("groups $G value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()})
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_png.err").read()[-600:])
PY
done
