#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ with the CPU oracle.

Run in the build container (it reads the reference's own weight files, so the fixtures also pin
the VPW1 conversion):

    python tools/make_golden.py

SeisBench itself is not importable here, so these vectors are outputs of the
oracle's restatement ("parity unpinned"); they freeze the oracle and let the GPU box check the CUDA
path without /root/reference.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import nets, pipeline  # noqa: E402
from volpick_b200.synthetic import synthetic_record  # noqa: E402

REF = "/root/reference/Final_models/volpick"
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (kind, station seed, n_samples, overlap, blinding, stacking)
    "phasenet_5min": ("phasenet", 0, 30000, 1500, (0, 0), "avg"),
    "phasenet_5min_max": ("phasenet", 3, 20011, 2000, (250, 250), "max"),
    "eqt_2min": ("eqtransformer", 1, 12000, 5500, (500, 500), "avg"),
    "eqt_tail_max": ("eqtransformer", 2, 13777, 3000, (500, 500), "max"),
}


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, (kind, station, n, overlap, blinding, stacking) in CASES.items():
        sd = nets.load_state_dict(f"{REF}/{kind}/volpick.pt.v1")
        with open(f"{REF}/{kind}/volpick.json.v1") as f:
            defaults = json.load(f)["default_args"]
        x = synthetic_record(station, n)
        ann, starts, y = pipeline.annotate_array(kind, sd, x, overlap, blinding, stacking, return_windows=True)
        thr = {"P_threshold": 0.2, "S_threshold": 0.2, "detection_threshold": defaults.get("detection_threshold", 0.3)}
        picks, offsets = pipeline.classify_array(kind, ann, thr)
        rows = [(pipeline.LABELS[kind].index(lab), *p) for lab, ps in picks.items() for p in ps]
        trig = np.array(rows, dtype=np.float64).reshape(-1, 5)
        win = pipeline.prenorm(pipeline.cut_windows(x, starts[:2], pipeline.IN_SAMPLES[kind]), kind)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            kind=kind, station=station, n_samples=n, overlap=overlap, blinding=np.array(blinding), stacking=stacking,
            starts=starts, annotation=ann.astype(np.float32), triggers=trig,
            trim=np.array([offsets[l] for l in pipeline.LABELS[kind]]),
            windows01=win.astype(np.float32), probs01=y[:2].astype(np.float32),
            thresholds=np.array([thr["detection_threshold"] if l == "Detection" else (0.0 if l == "N" else 0.2) for l in pipeline.LABELS[kind]]),
            input_checksum=np.array([float(np.abs(x).sum(dtype=np.float64))]),
        )
        print(name, "windows", len(starts), "annotation", ann.shape, "triggers", len(trig))


if __name__ == "__main__":
    main()
