#!/bin/bash
# Round 2: memcheck over both models (all modes), racecheck (a) on the stock library with decb excluded (VP_FUSED=0 path of
# tools/san_nodecb.py) and (b) on the -DVP_RACECHECK_BARRIERS build of decb (named barriers at the mbarrier hand-over points).
set -u
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, ".")
import volpick_b200 as vb
from volpick_b200.synthetic import synthetic_record
x = synthetic_record(7, 30_000)
kinds = os.environ.get("SAN_KINDS", "eqt,pn").split(",")
precs = os.environ.get("SAN_PRECS", "f16x3,bf16,fp32").split(",")
for cls in ([vb.EQTransformer] if "eqt" in kinds else []) + ([vb.PhaseNet] if "pn" in kinds else []):
    m = cls.from_pretrained("volpick").cuda()
    for prec in precs:
        a = m._argdict(dict(P_threshold=0.2, S_threshold=0.2, precision=prec, chunk_windows=16))
        ann, trig, trim = m.annotate_array(x, a, True, m._thresholds(a))
        print(cls.__name__, prec, len(trig), float(np.nanmax(ann)))
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python /tmp/san.py > gpurun_out/sanitize_r02_memcheck.log 2>&1
echo "memcheck exit: $?"; grep -E "ERROR SUMMARY|Invalid|Error" gpurun_out/sanitize_r02_memcheck.log | head -8
VP_FUSED=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python tools/san_nodecb.py > gpurun_out/sanitize_r02_racecheck_nodecb.log 2>&1
echo "racecheck (no decb) exit: $?"; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/sanitize_r02_racecheck_nodecb.log | head -8
SAN_KINDS=eqt SAN_PRECS=f16x3,bf16 VP_LIB_PATH=$PWD/volpick_b200/libvolpick_b200_rc.so timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python /tmp/san.py > gpurun_out/sanitize_r02_racecheck_decb_barriers.log 2>&1
echo "racecheck (decb with debug barriers) exit: $?"; grep -E "RACECHECK SUMMARY|hazard|fused" gpurun_out/sanitize_r02_racecheck_decb_barriers.log | sort | uniq -c | sort -rn | head -12
SAN_KINDS=eqt SAN_PRECS=f16x3 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python /tmp/san.py > gpurun_out/sanitize_r02_racecheck_stock.log 2>&1
echo "racecheck (stock) exit: $?"; grep -E "RACECHECK SUMMARY" gpurun_out/sanitize_r02_racecheck_stock.log | head -3; grep -oE "[a-z_0-9]+\.cu:[0-9]+" gpurun_out/sanitize_r02_racecheck_stock.log | sort | uniq -c | sort -rn | head -12
