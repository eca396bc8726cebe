#!/bin/bash
# ncu passes (never a bench value): launch list of one station-day, then a full-set capture.
# The .ncu-rep is converted to CSV on the box; it only travels back when it is small (<40 MiB).
set -u
mkdir -p gpurun_out
MODEL=${MODEL:-eqtransformer}
TAG=${TAG:-r01}
EXTRA=${BENCH_EXTRA:-}
if [ "${SKIP_LIST:-0}" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-1200} --csv \
    --log-file gpurun_out/launches_${TAG}_${MODEL}.csv python bench.py --model $MODEL --profile-steps 1 $EXTRA > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch-list exit: $?"
fi
if [ -n "${FULL_REGEX:-}" ]; then
REP=gpurun_out/prof_${TAG}_${MODEL}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:${FULL_REGEX} -s ${FULL_SKIP:-0} -c ${FULL_COUNT:-3} \
    -f -o $REP python bench.py --model $MODEL --profile-steps 1 $EXTRA > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit: $?"
ncu -i $REP.ncu-rep --page raw --csv > ${REP}_raw.csv 2>/dev/null
SZ=$(stat -c %s $REP.ncu-rep)
if [ "$SZ" -gt 41943040 ]; then rm -f $REP.ncu-rep; echo "dropped $REP.ncu-rep ($SZ bytes)"; fi
fi
du -sh gpurun_out
