#!/bin/bash
# BASELINE.json configs[2] / [3]: 1000 synthetic station-days sharded over N GPUs (one process per GPU, torchrun as the driver
# launches it), 16 distinct pinned records per rank, picks of every record gathered on rank 0.  N = $1.
set -u
N=${1:-2}
TAG=${TAG:-r02}
mkdir -p gpurun_out
STEPS=$((1000 / N))
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py \
      --gpus $N --steps $STEPS --warmup 3 --records 16 --quick "$@" > gpurun_out/scale_${TAG}_${name}_n$N.json 2> gpurun_out/scale_${TAG}_${name}_n$N.err
  echo "$name N=$N exit: $?"; tail -n 1 gpurun_out/scale_${TAG}_${name}_n$N.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('   value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'incl gather', round(d['gather']['e2e_value_incl_gather'],2), 'gather ms', round(d['gather']['ms'],2), 'records', d['gather']['records'], 'triggers', d['gather']['triggers'], 'clocks', d['clocks'])
except Exception as e: print('   parse failed', e)
"; }
run eqt
run pn --model phasenet
if [ "${WITH_BF16:-1}" = "1" ]; then run eqt_bf16 --precision bf16; fi
nvidia-smi topo -m > gpurun_out/scale_${TAG}_topo_n$N.txt 2>&1
