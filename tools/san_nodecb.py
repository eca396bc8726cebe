"""racecheck workload without the mbarrier-ordered fused decoder kernels: PhaseNet (all modes), EQTransformer fp32 and the
EQTransformer tensor-core path with the fused decoder kernels disabled (VP_FUSED=0 -> layer-by-layer tcconv)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import volpick_b200 as vb  # noqa: E402
from volpick_b200.synthetic import synthetic_record  # noqa: E402

x = synthetic_record(7, 30_000)
for cls, precs in ((vb.PhaseNet, ("f16x3", "bf16", "fp32")), (vb.EQTransformer, ("fp32", "f16x3"))):
    m = cls.from_pretrained("volpick").cuda()
    for prec in precs:
        a = m._argdict(dict(P_threshold=0.2, S_threshold=0.2, precision=prec, chunk_windows=16))
        ann, trig, trim = m.annotate_array(x, a, True, m._thresholds(a))
        print(cls.__name__, prec, len(trig), float(np.nanmax(ann)))
    w = torch.randn(5, 3, m.in_samples, device="cuda")
    print(cls.__name__, "pick_windows", {k: sum(len(p[0]) for p in v) for k, v in m.pick_windows(w, None, threshold=0.05).items()})
