import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import volpick_b200 as vb
from oracle import nets
from volpick_b200 import weights_io
from volpick_b200.synthetic import synthetic_record

prec = sys.argv[1]
sd = nets.state_dict_from_numpy(weights_io.load_weights(weights_io.find_weights("eqtransformer", "volpick")[1]))
rec = synthetic_record(21, 6000 * 9 + 10)
x = np.stack([rec[:, i * 6000:(i + 1) * 6000] for i in range(9)]).astype(np.float32)
x = x - x.mean(-1, keepdims=True); x = x / (np.abs(x).max(-1, keepdims=True) + 1e-10)
ref = torch.stack(nets.eqtransformer_forward(sd, torch.from_numpy(x)), dim=1).numpy()
m = vb.EQTransformer.from_pretrained("volpick").cuda()
got = torch.stack(m.forward(torch.from_numpy(x).cuda(), precision=prec), dim=1).cpu().numpy()
d = np.abs(got - ref)
print(prec, os.environ.get("VP_DECB_M2"), os.environ.get("VP_DECB_M1"), "max", d.max(), "nan", np.isnan(got).sum())
for g in range(3):
    dd = d[:, g, :]
    bad = np.argwhere(dd > 1e-3)
    print(" group", g, "max", dd.max(), "n_bad", len(bad), "first", bad[:3].tolist(), "last", bad[-3:].tolist())
    if len(bad):
        ts = np.unique(bad[:, 1]); print("   bad t range", ts.min(), ts.max(), "count unique t", len(ts), "windows", np.unique(bad[:,0]).tolist())
if os.environ.get("VP_DBG_DETAIL"):
    thr = float(os.environ["VP_DBG_DETAIL"])
    for g in range(3):
        bad = np.argwhere(d[0, g] > thr)[:, 0]
        print("detail group", g, "n", len(bad))
        print("  t:", bad[:60].tolist())
        print("  d:", np.round(d[0, g, bad[:20]], 3).tolist())
        print("  got:", np.round(got[0, g, bad[:10]], 3).tolist(), "ref:", np.round(ref[0, g, bad[:10]], 3).tolist())
