#!/bin/bash
# A/B of library variants on the EQT bench with all per-class times: LIBS="stock NAME ..."
set -u
mkdir -p gpurun_out
for v in ${LIBS:-stock}; do
libp=$PWD/volpick_b200/libvolpick_b200_$v.so; [ $v = stock ] && libp=$PWD/volpick_b200/libvolpick_b200.so
VP_LIB_PATH=$libp timeout 200 python bench.py --steps ${STEPS:-4} --warmup 3 --model ${MODEL:-eqtransformer} --no-cpu-baseline > gpurun_out/bench_lab.log 2>gpurun_out/bench_lab.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_lab.log").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$v value", round(d["value"],2), k)
except Exception as e:
    print("$v parse failed", e); print(open("gpurun_out/bench_lab.err").read()[-600:])
PY
done
if [ -n "${KEXPR:-}" ]; then VP_LIB_PATH=$PWD/volpick_b200/libvolpick_b200_${TESTLIB:-stock}.so timeout 400 python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider -k "$KEXPR" > gpurun_out/pytest_lab.log 2>&1; echo "pytest exit: $?"; tail -2 gpurun_out/pytest_lab.log; fi
