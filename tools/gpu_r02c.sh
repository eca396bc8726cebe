#!/bin/bash
# decb iteration: parity of everything that runs the decoder tail, then a short bench with per-class times.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --timeout=300 -p no:cacheprovider -s -k "${KEXPR:-tensor_core or annotate or forward_range or station_day or bf16 or golden}" > gpurun_out/pytest_decb.log 2>&1
echo "pytest exit: $?"; grep -E "passed|failed|error" gpurun_out/pytest_decb.log | tail -3
grep -E "max\|prob|match rate" gpurun_out/pytest_decb.log | tail -12
for prec in ${PRECS:-f16x3}; do
timeout 600 python bench.py --steps ${STEPS:-6} --warmup 3 --precision $prec --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_r02c_$prec.json 2> gpurun_out/bench_r02c_$prec.err
echo "bench exit: $?"; tail -c 300 gpurun_out/bench_r02c_$prec.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02c_$prec.json").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$prec", "value", round(d["value"],2), "seq", round(d["sequential"]["value"],2), "e2e", round(d["e2e"]["value"],2), k)
    print("roofline frac", round(d["roofline"]["frac"],4), "bf16", d.get("bf16"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_r02c_$prec.json").read()[-1500:])
PY
done
