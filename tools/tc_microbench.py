#!/usr/bin/env python
"""Sweep the tcgen05 conv layer micro-benchmark over the EQTransformer layer shapes (GPU box)."""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

LAYERS = [
    # name, CIN, T, COUT, K, mode, crop, pool, out_fmt, groups
    ("enc0", 3, 6000, 8, 11, 0, 0, 2, 0, 1), ("enc1", 8, 3000, 16, 9, 0, 0, 2, 0, 1), ("enc2", 16, 1500, 16, 7, 0, 0, 2, 0, 1),
    ("enc3", 16, 750, 32, 7, 0, 0, 2, 0, 1), ("enc4", 32, 375, 32, 5, 0, 0, 2, 0, 1), ("enc5", 32, 188, 64, 5, 0, 0, 2, 0, 1),
    ("enc6", 64, 94, 64, 3, 0, 0, 2, 1, 1),
    ("dec0", 16, 47, 64, 3, 1, 0, 1, 0, 3), ("dec1", 64, 94, 64, 5, 1, 0, 1, 0, 3), ("dec2", 64, 188, 32, 5, 2, 1, 1, 0, 3),
    ("dec3", 32, 375, 32, 7, 1, 0, 1, 0, 3), ("dec4", 32, 750, 16, 7, 1, 0, 1, 0, 3), ("dec5", 16, 1500, 16, 9, 1, 0, 1, 0, 3),
    ("dec6", 16, 3000, 8, 11, 1, 0, 1, 1, 3), ("res", 64, 47, 64, 3, 0, 0, 1, 1, 1),
]


def run(names, B, precision, iters=5):
    from volpick_b200 import _lib

    lib = _lib.load()
    out = {}
    for (name, cin, T, cout, k, mode, crop, pool, fmt, groups) in LAYERS:
        if names and name not in names:
            continue
        ms = C.c_float(0)
        _lib.check(lib.vp_tcconv_bench(B * groups, cin, T, cout, k, mode, crop, pool, _lib.PRECISION[precision], fmt, iters, C.byref(ms)))
        out[name] = ms.value * 1e3
    return out


if __name__ == "__main__":
    B = int(os.environ.get("B", "1024"))
    names = [a for a in sys.argv[1:] if not a.startswith("-")]
    if os.environ.get("VP_TC_SWEEP"):
        # re-run this script per debug mask (the mask is read at launch time from the environment)
        for mask in (0, 1, 2, 4, 3, 5, 6, 7):
            env = dict(os.environ, VP_TC_DBG=str(mask))
            env.pop("VP_TC_SWEEP")
            print(f"--- VP_TC_DBG={mask} (1: no A loads, 2: no MMAs, 4: no epilogue)", flush=True)
            subprocess.run([sys.executable, __file__] + names, env=env)
        sys.exit(0)
    for prec in os.environ.get("PRECS", "f16x3 bf16").split():
        res = run(names, B, prec)
        print(prec, " ".join(f"{k}={v:.0f}us" for k, v in res.items()), f"total={sum(res.values()):.0f}us", flush=True)
