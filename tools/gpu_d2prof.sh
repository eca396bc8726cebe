#!/bin/bash
set -u
mkdir -p gpurun_out
for dbg in ${DBGS:-0 4}; do
VP_DECB_DBG=$dbg VP_LIB_PATH=$PWD/volpick_b200/libvolpick_b200_prof.so timeout 200 python bench.py --steps 1 --warmup 1 --quick --model eqtransformer --precision ${PREC:-f16x3} --no-cpu-baseline > gpurun_out/bench_d2prof_$dbg.log 2> gpurun_out/bench_d2prof_$dbg.err
echo "== dbg $dbg"; grep -A13 "decb2 prof" gpurun_out/bench_d2prof_$dbg.err | tail -14
done
