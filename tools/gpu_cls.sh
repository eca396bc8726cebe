#!/bin/bash
# classify(stream) timing of both models with the host profile, after the stream / classify parity tests
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider -k "${KEXPR:-helper_thread or stream or classify or phasenet or pn_}" 2>&1 | tail -4
for m in ${MODELS:-phasenet eqtransformer}; do
  VP_PROFILE_HOST=1 timeout 250 python bench.py --model $m --steps 6 --warmup 3 --records 16 --classify-stream ${NREC:-8} --no-cpu-baseline > gpurun_out/bench_cls_$m.json 2> gpurun_out/bench_cls_$m.err
  grep 'host profile' gpurun_out/bench_cls_$m.err | tail -2
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cls_$m.json').read().strip().splitlines()[-1])
print("$m", round(d['value'],2), round(d['e2e']['value'],2), d['e2e_classify']['copy_true'], d['e2e_classify']['copy_false'])
print({a:round(b['ms_per_step'],3) for a,b in d['kernels']['per_class'].items()})
PY
done
