#!/bin/bash
# racecheck on the -DVP_RACECHECK_BARRIERS build (named barriers at the mbarrier hand-over points of the decoder-tail kernels)
set -u
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, ".")
import volpick_b200 as vb
from volpick_b200.synthetic import synthetic_record
x = synthetic_record(7, 30_000)
m = vb.EQTransformer.from_pretrained("volpick").cuda()
for prec in ("f16x3", "bf16"):
    a = m._argdict(dict(P_threshold=0.2, S_threshold=0.2, precision=prec, chunk_windows=16))
    ann, trig, trim = m.annotate_array(x, a, True, m._thresholds(a))
    print("EQT", prec, len(trig), float(np.nanmax(ann)))
PY
VP_LIB_PATH=$PWD/volpick_b200/libvolpick_b200_rc.so timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python /tmp/san.py > gpurun_out/sanitize_d2_racecheck_barriers.log 2>&1
echo "racecheck (debug barriers) exit: $?"; grep -E "RACECHECK SUMMARY|EQT" gpurun_out/sanitize_d2_racecheck_barriers.log | head -5; grep -oE "[a-z_0-9]+\.cu:[0-9]+" gpurun_out/sanitize_d2_racecheck_barriers.log | sort | uniq -c | sort -rn | head
