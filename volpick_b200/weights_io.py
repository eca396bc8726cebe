"""Weight containers for the volpick models.

``from_pretrained`` (SeisBench ``SeisBenchModel.from_pretrained`` / ``load``; reference call sites
/root/reference/README.md:46-47, /root/reference/Final_models/demo.ipynb cells 7-8) looks for a
``<name>.json.v<ver>`` + ``<name>.pt.v<ver>`` pair in ``<cache_root>/<modelclass lowercase>/``.  This
module keeps that layout and adds a torch-free twin of the ``.pt`` file, ``<name>.vpw.v<ver>``:

    bytes 0..3   magic  b"VPW1"
    bytes 4..7   uint32 little endian: length H of the JSON header
    H bytes      JSON   {"tensors": [{"name", "shape", "offset" (in floats)}, ...], "n_floats"}
    then         float32 little endian payload, tensors back to back in state-dict order
                 (BatchNorm ``num_batches_tracked`` counters are dropped)

The four weight sets shipped by the reference (/root/reference/Final_models/*/*/*.pt.v1) are
converted by ``tools/convert_weights.py`` into ``volpick_b200/weights/``.
"""
from __future__ import annotations

import json
import os
import struct
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np

MAGIC = b"VPW1"


def write_vpw(path: str, tensors: "OrderedDict[str, np.ndarray]") -> None:
    entries = []
    offset = 0
    chunks = []
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype="<f4")
        entries.append({"name": name, "shape": list(a.shape), "offset": offset})
        offset += a.size
        chunks.append(a.reshape(-1))
    header = json.dumps({"tensors": entries, "n_floats": offset}).encode()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(header)))
        f.write(header)
        f.write(np.concatenate(chunks).tobytes())


def read_vpw(path: str) -> "OrderedDict[str, np.ndarray]":
    with open(path, "rb") as f:
        if f.read(4) != MAGIC:
            raise ValueError(f"{path}: not a VPW1 weight file")
        (hlen,) = struct.unpack("<I", f.read(4))
        header = json.loads(f.read(hlen).decode())
        data = np.frombuffer(f.read(), dtype="<f4")
    if data.size != header["n_floats"]:
        raise ValueError(f"{path}: payload has {data.size} floats, header says {header['n_floats']}")
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for e in header["tensors"]:
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        out[e["name"]] = data[e["offset"] : e["offset"] + n].reshape(e["shape"]).copy()
    return out


def read_pt(path: str) -> "OrderedDict[str, np.ndarray]":
    """Read a SeisBench ``.pt`` state dict (tensors only) -- needs torch."""
    import torch

    sd = torch.load(path, weights_only=True, map_location="cpu")
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        out[k] = v.detach().to(torch.float32).numpy().copy()
    return out


def default_cache_root() -> str:
    """``VOLPICK_B200_CACHE`` (or ``SEISBENCH_CACHE_ROOT``/models/v3) first, the in-package weights last."""
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")


def search_roots() -> Tuple[str, ...]:
    roots = []
    if os.environ.get("VOLPICK_B200_CACHE"):
        roots.append(os.environ["VOLPICK_B200_CACHE"])
    if os.environ.get("SEISBENCH_CACHE_ROOT"):
        roots.append(os.path.join(os.environ["SEISBENCH_CACHE_ROOT"], "models", "v3"))
    roots.append(default_cache_root())
    return tuple(roots)


def find_weights(model_dir: str, name: str, version_str: Optional[str] = None) -> Tuple[str, str]:
    """Return (json_path, weights_path) for ``<name>`` under ``<root>/<model_dir>/``; newest version wins."""
    tried = []
    for root in search_roots():
        d = os.path.join(root, model_dir)
        tried.append(d)
        if not os.path.isdir(d):
            continue
        versions = []
        for fn in os.listdir(d):
            if fn.startswith(name + ".json.v"):
                versions.append(fn[len(name) + len(".json.v") :])
        if version_str is not None:
            versions = [v for v in versions if v == str(version_str)]
        if not versions:
            continue
        # type-stable key: numeric parts sort before (and among themselves numerically), text parts lexically
        ver = sorted(versions, key=lambda s: [(0, int(p), "") if p.isdigit() else (1, 0, p) for p in s.split(".")])[-1]
        js = os.path.join(d, f"{name}.json.v{ver}")
        for ext in ("vpw", "pt"):
            w = os.path.join(d, f"{name}.{ext}.v{ver}")
            if os.path.exists(w):
                return js, w
    raise ValueError(f"No weights '{name}' (version {version_str or 'latest'}) found; looked in {tried}")


def load_weights(path: str) -> "OrderedDict[str, np.ndarray]":
    return read_vpw(path) if ".vpw." in os.path.basename(path) else read_pt(path)
