#!/bin/bash
set -u
mkdir -p gpurun_out
VP_LIB_PATH=$PWD/volpick_b200/libvolpick_b200_eaprof.so timeout 200 python bench.py --steps 1 --warmup 1 --quick --model eqtransformer --no-cpu-baseline > gpurun_out/bench_eaprof.log 2> gpurun_out/bench_eaprof.err
grep "enca prof" gpurun_out/bench_eaprof.err | tail -3
