#!/bin/bash
# ncu --set full captures (never a bench value) of named kernels of one EQTransformer/PhaseNet station-day.
#   KERNELS="decb_kernel:0:1 deca_kernel:0:1" TAG=r01b bash tools/gpu_ncu_full.sh      (regex:skip:count)
set -u
mkdir -p gpurun_out
MODEL=${MODEL:-eqtransformer}
TAG=${TAG:-r01}
EXTRA=${BENCH_EXTRA:---precision f16x3}
for spec in ${KERNELS}; do
  IFS=: read -r rx skip cnt <<< "$spec"
  REP=gpurun_out/full_${TAG}_${MODEL}_${rx//[^a-zA-Z0-9_]/_}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${rx} -s ${skip:-0} -c ${cnt:-1} \
      -f -o $REP python bench.py --model $MODEL --profile-steps 1 $EXTRA > ${REP}.log 2>&1
  echo "full capture ${rx} exit: $?"
  ncu -i $REP.ncu-rep --page raw --csv > ${REP}_raw.csv 2>/dev/null
  SZ=$(stat -c %s $REP.ncu-rep)
  if [ "$SZ" -gt 20971520 ]; then rm -f $REP.ncu-rep; echo "dropped $REP.ncu-rep ($SZ bytes)"; fi
done
du -sh gpurun_out
