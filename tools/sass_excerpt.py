#!/usr/bin/env python
"""Counts of the Blackwell-specific SASS instructions per kernel of the built library (profiles/sass_excerpt.txt).

    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt

tcgen05.mma -> UTCHMMA / UTCQMMA..., tcgen05.ld / st -> LDTM / STTM, tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS...,
cp.async.bulk.tensor -> UTMALDG / UTMASTG, cp.async.bulk -> UBLKCP, cp.async -> LDGSTS, mbarrier -> SYNCS (B200_PROFILING.md).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "volpick_b200", "libvolpick_b200.so")
PAT = re.compile(r"\b(UTC[A-Z0-9]*|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTMAPF|LDGSTS|SYNCS|FENCE|ELECT)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        if cur is None:
            continue
        m = PAT.search(line.split("/*")[1] if line.strip().startswith("/*") and line.count("/*") > 1 else line)
        if m:
            cur[m.group(1)] += 1
    tot = collections.Counter()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} | per-kernel counts of tcgen05 / TMEM / TMA / mbarrier instructions\n")
    for name, c in per.items():
        if not any(k.startswith("UTC") or k in ("LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMASTG") for k in c):
            continue
        tot.update(c)
        print(name)
        print("    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    print("\nTOTAL  " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
