#!/bin/bash
# A/B of library variants on the EQT bench: LIBS="stock sleep20 ..." (stock = the default library)
set -u
mkdir -p gpurun_out
for v in ${LIBS:-stock}; do
for prec in ${PRECS:-f16x3}; do
libp=$PWD/volpick_b200/libvolpick_b200_$v.so; [ $v = stock ] && libp=$PWD/volpick_b200/libvolpick_b200.so
env ${ENVX:-X=0} VP_LIB_PATH=$libp timeout 200 python bench.py --steps ${STEPS:-3} --warmup 2 --model eqtransformer --precision $prec --no-cpu-baseline > gpurun_out/bench_ab_${v}_$prec.log 2>gpurun_out/bench_ab_${v}_$prec.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ab_${v}_$prec.log").read().strip().splitlines()[-1])
    k={a:round(b["ms_per_step"],3) for a,b in d["kernels"]["per_class"].items()}
    print("$v $prec", "value", round(d["value"],2), "decb", k.get("decb"), "maxdiff", d.get("parity"))
except Exception as e:
    print("$v bench parse failed", e); print(open("gpurun_out/bench_ab_${v}_$prec.err").read()[-800:])
PY
done
done
