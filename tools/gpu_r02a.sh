#!/bin/bash
# Round 2, pass A: the whole GPU parity suite, then a short EQTransformer bench with the drop-in classify(stream) arm.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt
timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider -s ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu_r02.log 2>&1
echo "pytest exit: $?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu_r02.log | tail -3
grep -E "max\|prob|exposure|match rate|station-day" gpurun_out/pytest_gpu_r02.log | tail -40
for model in ${MODELS:-eqtransformer}; do
timeout 600 python bench.py --steps ${STEPS:-8} --warmup 3 --model $model --classify-stream ${CS:-4} --no-cpu-baseline > gpurun_out/bench_r02a_$model.json 2> gpurun_out/bench_r02a_$model.err
echo "bench exit: $?"; tail -c 400 gpurun_out/bench_r02a_$model.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02a_$model.json").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$model", "value", round(d["value"],2), "seq", round(d["sequential"]["value"],2), "e2e", round(d["e2e"]["value"],2), k)
    print("classify", d.get("e2e_classify")); print("gather", d.get("gather")); print("roofline frac", d["roofline"]["frac"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_r02a_$model.json").read()[-1500:])
PY
done
