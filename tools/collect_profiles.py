#!/usr/bin/env python
"""Copy the evidence of one GPU pass (tools/gpu_r02_final.sh, TAG) from the scratch gpurun_out/ into the tracked profiles/:
bench lines, forward sweep, launch lists (csv + per-kernel shares), ncu --set full summaries, SASS excerpt.

    python tools/collect_profiles.py r02a
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed",
        "lts__t_requests.sum", "lts__t_sectors.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed_pipe_uniform.sum"]


def main(tag):
    os.makedirs(P, exist_ok=True)
    for fn in sorted(os.listdir(G)):
        if fn.startswith(f"bench_{tag}_") and (fn.endswith(".json") or fn.endswith(".jsonl")):
            lines = [l for l in open(os.path.join(G, fn)).read().splitlines() if l.startswith("{")]
            if lines:
                with open(os.path.join(P, fn.replace("bench_" + tag, tag + "_bench")), "w") as f:
                    f.write("\n".join(lines) + "\n")
        if fn.startswith(f"scale_{tag}") and fn.endswith(".json"):
            lines = [l for l in open(os.path.join(G, fn)).read().splitlines() if l.startswith("{")]
            if lines:
                with open(os.path.join(P, fn.replace("scale_" + tag, tag + "_scale")), "w") as f:
                    f.write(lines[-1] + "\n")
        if fn.startswith(f"launches_{tag}_") and fn.endswith(".csv"):
            shutil.copy(os.path.join(G, fn), os.path.join(P, fn.replace("launches_" + tag, tag + "_launches")))
            buf = io.StringIO()
            with redirect_stdout(buf):
                ncu_summary.launches(os.path.join(G, fn), md=True)
            model = fn[len(f"launches_{tag}_"):-4]
            with open(os.path.join(P, f"{tag}_launches_{model}.md"), "w") as f:
                f.write(f"# {tag}: ncu launch list of one {model} station-day (f16x3), `ncu --metrics gpu__time_duration.sum --clock-control none "
                        f"python bench.py --model {model} --profile-steps 1 --precision f16x3`\n\nPer-launch times under ncu are cold-cache and "
                        "serialised: the SHARES are what must agree with the CUDA-event pass of bench.py (`kernels.per_class`).\n\n" + buf.getvalue())
    out = [f"# {tag}: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 0 -c 1 python bench.py --profile-steps 1 "
           "--precision f16x3` (tools/gpu_ncu_full.sh), one 4096-window launch each\n\n```"]
    for fn in sorted(os.listdir(G)):
        if fn.startswith(f"full_{tag}_") and fn.endswith("_raw.csv"):
            rows = list(csv.reader(open(os.path.join(G, fn))))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            for r in rows[2:]:
                out.append("== " + r[hdr.index("Kernel Name")][:100])
                for k in KEYS:
                    if k in hdr:
                        i = hdr.index(k)
                        out.append(f"   {k:84s} {r[i]:>18s} {units[i]}")
    out.append("```")
    if len(out) > 3:
        with open(os.path.join(P, f"{tag}_f16x3_top_kernels_ncu_full.md"), "w") as f:
            f.write("\n".join(out) + "\n")
    sass = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_excerpt.py")], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{tag}_sass_excerpt.txt"), "w") as f:
        f.write(sass)
    for fn in (f"pytest_gpu_{tag}.log", f"smoke_{tag}.log"):
        if os.path.exists(os.path.join(G, fn)):
            txt = open(os.path.join(G, fn)).read()
            with open(os.path.join(P, f"{tag}_" + fn.replace(f"_{tag}", "")), "w") as f:
                f.write(txt[-20000:])
    print("collected", sorted(fn for fn in os.listdir(P) if fn.startswith(tag)))


if __name__ == "__main__":
    main(sys.argv[1])
