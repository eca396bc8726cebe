#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory pipelines) over a small end-to-end run of both models.
set -u
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, ".")
import volpick_b200 as vb
from volpick_b200.synthetic import synthetic_record
x = synthetic_record(7, 30_000)
for cls in (vb.EQTransformer, vb.PhaseNet):
    m = cls.from_pretrained("volpick").cuda()
    for prec in ("f16x3", "bf16", "fp32"):
        a = m._argdict(dict(P_threshold=0.2, S_threshold=0.2, precision=prec, chunk_windows=16))
        ann, trig, trim = m.annotate_array(x, a, True, m._thresholds(a))
        print(cls.__name__, prec, len(trig), float(np.nanmax(ann)))
    w = torch.randn(5, 3, m.in_samples, device="cuda")
    print(cls.__name__, "pick_windows", {k: sum(len(p[0]) for p in v) for k, v in m.pick_windows(w, None, threshold=0.05).items()})
PY
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit: $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|Error" gpurun_out/sanitize_$tool.log | head -12
  grep -E "EQTransformer|PhaseNet" gpurun_out/sanitize_$tool.log | head -8
done
