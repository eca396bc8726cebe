#!/bin/bash
# Round-2 evidence pass on one B200 (TAG names the outputs gpurun_out/*_${TAG}*): full parity suite, smoke, the bench lines of
# every single-GPU BASELINE.json configuration, the reference arm, ncu launch lists and --set full captures of the top kernels.
set -u
TAG=${TAG:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1; nproc >> gpurun_out/gpu_info.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -s > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke_${TAG}.log; tail -3 gpurun_out/smoke_${TAG}.log
fi
B() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "bench $name exit: $?"; }
B eqt --steps 20 --warmup 3 --records 16 --classify-stream 8
B eqt_bf16 --steps 20 --warmup 3 --records 16 --precision bf16 --no-cpu-baseline
B pn --model phasenet --steps 40 --warmup 3 --records 16 --classify-stream 8
B pn_bf16 --model phasenet --steps 40 --warmup 3 --records 16 --precision bf16 --no-cpu-baseline
B pn_hour --model phasenet --samples 360000 --steps 40 --warmup 3 --records 16 --no-cpu-baseline
B reference --impl reference --steps 2 --warmup 1
B reference_pn --impl reference --model phasenet --steps 3 --warmup 1
timeout 900 python bench.py --sweep --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_sweep.jsonl 2> gpurun_out/bench_${TAG}_sweep.err; echo "sweep exit: $?"
python - <<PY
import json
for n in ("eqt", "eqt_bf16", "pn", "pn_bf16", "pn_hour", "reference", "reference_pn"):
    try:
        d = json.loads(open("gpurun_out/bench_${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(n, "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "roofline", r.get("kernel"), r.get("frac"), "clocks", d.get("clocks"))
        if "e2e_classify" in d: print("   classify", {k: v for k, v in d["e2e_classify"].items() if k.startswith("copy")})
        if "bf16" in d: print("   bf16", d["bf16"])
        if "kernels" in d: print("   ", {a: round(b["ms_per_step"], 3) for a, b in d["kernels"]["per_class"].items()})
    except Exception as e:
        print(n, "parse failed", e)
for line in open("gpurun_out/bench_${TAG}_sweep.jsonl"):
    try:
        d = json.loads(line); print("sweep", d["model"], d["precision"], d["windows"], "ms", round(d["ms"], 3), "win/s", round(d["windows_per_s"]), "TF", round(d["tflops"], 1))
    except Exception: pass
PY
if [ "${SKIP_NCU:-0}" != "1" ]; then
for MODEL in eqtransformer phasenet; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_${MODEL}.csv \
    python bench.py --model $MODEL --profile-steps 1 --precision f16x3 > gpurun_out/ncu_launches_${TAG}_${MODEL}.log 2>&1; echo "launch list $MODEL exit: $?"
done
KERNELS="${KERNELS:-decb_kernel:0:1 resstack2_kernel:0:1 deca_kernel:0:1}" TAG=$TAG bash tools/gpu_ncu_full.sh
fi
