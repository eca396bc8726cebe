#!/bin/bash
# compute-sanitizer on the EQTransformer path with the TMEM-operand decoder tail (memcheck, then racecheck for the record)
set -u
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, ".")
import volpick_b200 as vb
from volpick_b200.synthetic import synthetic_record
x = synthetic_record(7, 30_000)
m = vb.EQTransformer.from_pretrained("volpick").cuda()
for prec in os.environ.get("SAN_PRECS", "f16x3,bf16").split(","):
    a = m._argdict(dict(P_threshold=0.2, S_threshold=0.2, precision=prec, chunk_windows=16))
    ann, trig, trim = m.annotate_array(x, a, True, m._thresholds(a))
    print("EQT", prec, len(trig), float(np.nanmax(ann)))
    xw = torch.randn(5, 3, 6000, device="cuda")
    y = m.forward(xw, precision=prec)
    print("forward", prec, float(y[0].max()))
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python /tmp/san.py > gpurun_out/sanitize_d2_memcheck.log 2>&1
echo "memcheck exit: $?"; grep -E "ERROR SUMMARY|Invalid|Error|EQT|forward" gpurun_out/sanitize_d2_memcheck.log | head -12
SAN_PRECS=f16x3 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python /tmp/san.py > gpurun_out/sanitize_d2_racecheck.log 2>&1
echo "racecheck exit: $?"; grep -E "RACECHECK SUMMARY" gpurun_out/sanitize_d2_racecheck.log | head -3; grep -oE "[a-z_0-9]+\.cu:[0-9]+" gpurun_out/sanitize_d2_racecheck.log | sort | uniq -c | sort -rn | head -20
