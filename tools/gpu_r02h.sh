#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --timeout=300 -p no:cacheprovider -k "tensor_core or forward_range or annotate_tensor or golden or slice_forward_scopes or station_day_matches" > gpurun_out/pytest_decb.log 2>&1
echo "pytest exit: $?"; tail -2 gpurun_out/pytest_decb.log
for lib in libvolpick_b200.so libvolpick_b200_v.so; do
for prec in f16x3 bf16; do
VP_LIB_PATH=$PWD/volpick_b200/$lib timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --precision $prec > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_h.json").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$lib $prec", "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), k)
except Exception as e:
    print("$lib parse failed", e, open("gpurun_out/bench_h.err").read()[-600:])
PY
done
done
