#!/bin/bash
# one ncu --set full capture of decb2_kernel (a 4096-window launch), raw + source pages exported as CSV
set -u
mkdir -p gpurun_out
REP=gpurun_out/full_${TAG:-r02b}_decb2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decb2_kernel -s ${SKIP:-1} -c 1 -f -o $REP python bench.py --model eqtransformer --profile-steps 1 --precision ${PREC:-f16x3} > ${REP}.log 2>&1
echo "capture exit: $?"
ncu -i $REP.ncu-rep --page raw --csv > ${REP}_raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --print-source sass > ${REP}_source.csv 2>/dev/null
ncu -i $REP.ncu-rep --page details > ${REP}_details.txt 2>/dev/null
ls -la gpurun_out/full_${TAG:-r02b}_decb2*
