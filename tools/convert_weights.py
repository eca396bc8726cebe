#!/usr/bin/env python
"""Convert the reference's shipped SeisBench weight pairs into the torch-free VPW1 container.

    python tools/convert_weights.py [/root/reference/Final_models] [volpick_b200/weights]

Source layout  (reference):  Final_models/<set>/<model>/<name>.{pt,json}.v1
Target layout  (SeisBench cache layout, /root/reference/README.md:12): weights/<model>/<name>.{vpw,json}.v1
"""
import os
import shutil
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from volpick_b200 import weights_io  # noqa: E402


def main() -> None:
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/Final_models"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "volpick_b200", "weights")
    for wset in sorted(os.listdir(src)):
        if not os.path.isdir(os.path.join(src, wset)):
            continue
        for model in sorted(os.listdir(os.path.join(src, wset))):
            d = os.path.join(src, wset, model)
            for fn in sorted(os.listdir(d)):
                if ".pt.v" not in fn:
                    continue
                name, ver = fn.split(".pt.v")
                tensors = weights_io.read_pt(os.path.join(d, fn))
                os.makedirs(os.path.join(dst, model), exist_ok=True)
                out = os.path.join(dst, model, f"{name}.vpw.v{ver}")
                weights_io.write_vpw(out, tensors)
                shutil.copyfile(os.path.join(d, f"{name}.json.v{ver}"), os.path.join(dst, model, f"{name}.json.v{ver}"))
                n = sum(v.size for v in tensors.values())
                print(f"{out}: {len(tensors)} tensors, {n} floats")


if __name__ == "__main__":
    main()
