#!/usr/bin/env python
"""Summarise ncu CSV outputs brought back in gpurun_out/ (launch list -> per-kernel shares; raw page -> key metrics)."""
import collections
import csv
import re
import sys


def launches(path, md=False):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(r[ui], v)
        name = re.sub(r"\(.*", "", r[ki])[:80]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"Sum of kernel durations: {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches\n")
    if md:
        print("| share | total ms | launches | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if md:
            print(f"| {100 * v[1] / tot:.1f}% | {v[1]:.3f} | {v[0]} | `{k}` |")
        else:
            print(f"{100 * v[1] / tot:5.1f}% {v[1]:9.3f} ms {v[0]:5d}  {k}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_umma.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__inst_executed_pipe_fma.sum",
        "smsp__inst_executed_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ni = hdr.index("Kernel Name")
    for r in data:
        print("==", r[ni][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:72s} {r[i]:>16s} {units[i]}")
        for i, h in enumerate(hdr):
            if "tensor" in h and h not in WANT:
                print(f"   {h:72s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], md="--md" in sys.argv)
    else:
        raw(sys.argv[2])
