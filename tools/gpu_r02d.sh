#!/bin/bash
# decb variants: tile rows m and the separate level-2 buffer (VP_DECB_Z)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=300 -p no:cacheprovider -k "tensor_core or forward_range or annotate_tensor" > gpurun_out/pytest_decb.log 2>&1
echo "pytest exit: $?"; tail -2 gpurun_out/pytest_decb.log
for cfg in "VP_DECB_M2=53 VP_DECB_Z=0" "VP_DECB_M2=47 VP_DECB_Z=0" "VP_DECB_M2=47 VP_DECB_Z=1" "VP_DECB_M2=45 VP_DECB_Z=1"; do
env $cfg timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_d.json").read().strip().splitlines()[-1])
    print("$cfg", "decb", round(d["kernels"]["per_class"]["decb"]["ms_per_step"],3), "value", round(d["value"],2))
except Exception as e:
    print("$cfg parse failed", e, open("gpurun_out/bench_d.err").read()[-600:])
PY
done
