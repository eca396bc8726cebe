#!/bin/bash
# Round-final pass on one B200: full parity suite, smoke, bench lines of both models / precisions, the reference arm,
# ncu launch lists and full captures of the top kernels.  TAG names the outputs (gpurun_out/*_${TAG}_*).
set -u
TAG=${TAG:-r01g}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_${TAG}_eqt.json 2> gpurun_out/bench_${TAG}_eqt.err; echo "bench eqt exit: $?"
timeout 300 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/bench_${TAG}_eqt_bf16.json 2>/dev/null; echo "bench eqt bf16 exit: $?"
timeout 300 python bench.py --model phasenet > gpurun_out/bench_${TAG}_pn.json 2>/dev/null; echo "bench pn exit: $?"
timeout 300 python bench.py --model phasenet --precision bf16 --no-cpu-baseline > gpurun_out/bench_${TAG}_pn_bf16.json 2>/dev/null; echo "bench pn bf16 exit: $?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>/dev/null; echo "bench reference exit: $?"
python - <<PY
import json
for n in ("eqt", "eqt_bf16", "pn", "pn_bf16", "reference"):
    try:
        d = json.loads(open("gpurun_out/bench_${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(n, "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "roofline", r.get("kernel"), r.get("frac"), "clocks", d.get("clocks"))
    except Exception as e:
        print(n, "parse failed", e)
PY
for MODEL in eqtransformer phasenet; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_${MODEL}.csv \
    python bench.py --model $MODEL --profile-steps 1 --precision f16x3 > gpurun_out/ncu_launches_${TAG}_${MODEL}.log 2>&1; echo "launch list $MODEL exit: $?"
done
KERNELS="${KERNELS:-decb_kernel:0:1 resstack_kernel:0:1 deca_kernel:0:1}" TAG=$TAG bash tools/gpu_ncu_full.sh
