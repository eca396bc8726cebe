#!/bin/bash
# A/B of the attention kernels (VP_ATTN_V1=1: one thread per query; default: two lanes per query) + the EQT parity tests
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider -k "${KEXPR:-eqt or eqtransformer or golden or forward or station_day}" > gpurun_out/pytest_attn.log 2>&1
echo "pytest exit: $?"; tail -3 gpurun_out/pytest_attn.log
for v in 0 1; do
  VP_ATTN_V1=$v timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attn.log 2> gpurun_out/bench_attn.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_attn.log").read().strip().splitlines()[-1])
    print("VP_ATTN_V1=$v value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()})
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_attn.err").read()[-600:])
PY
done
