"""Reference pin, conditional on SeisBench being importable (last row).

SeisBench and ObsPy are not part of this image and cannot be installed here (no network), so on this box every test in
this file SKIPS and the oracle stays "parity unpinned".  Wherever ``seisbench`` imports, the same file pins the oracle --
and through it the CUDA path -- to the real thing on the golden records:

* ``sbm.EQTransformer`` / ``sbm.PhaseNet`` built from the shipped ``model_args`` with the shipped weights: forward on the
  golden windows against ``oracle.nets`` (<= 1e-5) and, with a GPU, against the CUDA forward (<= 1e-4);
* ``annotate_batch_pre`` against ``oracle.pipeline.prenorm`` under both amplitude scopes: reports which reading of
  ``norm="peak"`` SeisBench implements and fails when it is not the oracle's default;
* ``annotate`` / ``classify`` on a synthetic ObsPy stream against ``oracle.pipeline`` (probabilities <= 1e-5, identical
  pick indices) and against the CUDA ``classify``.
"""
import json

import numpy as np
import pytest

sbm = pytest.importorskip("seisbench.models")
obspy = pytest.importorskip("obspy")
import torch  # noqa: E402

from oracle import nets, pipeline  # noqa: E402
from volpick_b200 import weights_io  # noqa: E402
from volpick_b200.synthetic import synthetic_record  # noqa: E402

KINDS = {"eqtransformer": "EQTransformer", "phasenet": "PhaseNet"}


def _sb_model(kind):
    js, wpath = weights_io.find_weights(kind, "volpick")
    with open(js) as f:
        meta = json.load(f)
    model = getattr(sbm, KINDS[kind])(**meta.get("model_args", {}))
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights_io.load_weights(wpath).items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    model.eval()
    return model, meta


def _obspy_stream(x, station="S0000"):
    t0 = obspy.UTCDateTime("2020-01-01T00:00:00")
    return obspy.Stream([
        obspy.Trace(x[i].copy(), header=dict(network="XX", station=station, location="", channel="HH" + c, starttime=t0,
                                             sampling_rate=100.0))
        for i, c in enumerate("ZNE")
    ])


@pytest.mark.parametrize("name", ["eqt_2min", "eqt_tail_max", "phasenet_5min", "phasenet_5min_max"])
def test_oracle_forward_matches_seisbench(name, golden):
    """SeisBench's forward on the golden windows against the oracle's nets and against the committed fixture."""
    g = golden[name]
    kind = str(g["kind"])
    model, _ = _sb_model(kind)
    x = torch.from_numpy(g["windows01"])
    with torch.no_grad():
        out = model(x)
    out = torch.stack(list(out), dim=-1) if isinstance(out, (tuple, list)) else out.permute(0, 2, 1)  # heads-last (B, L, 3)
    sd = nets.state_dict_from_numpy(weights_io.load_weights(weights_io.find_weights(kind, "volpick")[1]))
    ref = pipeline.forward_batches(kind, sd, g["windows01"], 64)
    assert float(np.abs(out.numpy() - ref).max()) <= 1e-5
    assert float(np.abs(out.numpy() - g["probs01"]).max()) <= 1e-5  # the committed fixture itself


@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
def test_annotate_batch_pre_scope(kind):
    """Which amplitude scope does SeisBench's norm="peak" use?  The oracle's default must be the one."""
    model, _ = _sb_model(kind)
    L = pipeline.IN_SAMPLES[kind]
    x = synthetic_record(46, L + 500)
    x[1] *= 0.2
    win = x[None, :, :L].astype(np.float32)
    if hasattr(model, "annotate_batch_pre"):
        got = model.annotate_batch_pre(torch.from_numpy(win.copy()), {}).numpy()
    else:  # SeisBench 0.4.x: per-window NumPy hook
        got = np.stack([model.annotate_window_pre(w.copy(), {}) for w in win]).astype(np.float32)
    err = {scope: float(np.abs(got - pipeline.prenorm(win, kind, "peak", scope)).max()) for scope in ("channel", "window")}
    print(f"{kind}: |seisbench - oracle| per-channel scope {err['channel']:.2e}, per-window scope {err['window']:.2e}")
    assert err["channel"] <= 1e-5, f"SeisBench does not normalise per component: {err}"


@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
def test_annotate_and_classify_match_seisbench(kind):
    model, meta = _sb_model(kind)
    x = synthetic_record(0, 60_000)
    st = _obspy_stream(x)
    kw = dict(overlap=5500, blinding=(500, 500)) if kind == "eqtransformer" else {}
    thr = dict(meta.get("default_args", {}))
    ann = model.annotate(st, **kw)
    sd = nets.state_dict_from_numpy(weights_io.load_weights(weights_io.find_weights(kind, "volpick")[1]))
    ref = pipeline.annotate_array(kind, sd, x, kw.get("overlap"), kw.get("blinding"))
    for li, label in enumerate(pipeline.LABELS[kind]):
        tr = ann.select(channel=f"{KINDS[kind]}_{label}")[0]
        col, f, _ = pipeline.trim_nan(ref[:, li])
        assert tr.stats.npts == len(col)
        assert abs((tr.stats.starttime - st[0].stats.starttime) - f / 100.0) < 1e-6
        assert float(np.abs(tr.data - col).max()) <= 1e-5
    out = model.classify(st, **kw, **{k: v for k, v in thr.items() if k.endswith("_threshold")})
    picks = out.picks if hasattr(out, "picks") else out[0]
    ref_picks, _ = pipeline.classify_array(kind, ref, {**{"detection_threshold": 0.3}, **thr})
    for phase in "PS":
        mine = sorted(p for p in picks if p.phase == phase)
        assert len(mine) == len(ref_picks[phase])
        for p, r in zip(mine, ref_picks[phase]):
            assert round((p.peak_time - st[0].stats.starttime) * 100) == r[2]
            assert abs(p.peak_value - r[3]) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
def test_cuda_classify_matches_seisbench(kind):
    import volpick_b200 as vb

    model, meta = _sb_model(kind)
    mine = getattr(vb, KINDS[kind]).from_pretrained("volpick").cuda()
    x = synthetic_record(0, 60_000)
    st = _obspy_stream(x)
    kw = dict(overlap=5500, blinding=(500, 500)) if kind == "eqtransformer" else {}
    a_ref, a_got = model.annotate(st, **kw), mine.annotate(st, **kw)
    for tr in a_ref:
        got = a_got.select(channel=tr.stats.channel)[0]
        assert got.stats.starttime == tr.stats.starttime and got.stats.npts == tr.stats.npts
        assert float(np.abs(got.data - tr.data).max()) <= 1e-4
    p_ref = model.classify(st, **kw).picks
    p_got = mine.classify(st, **kw).picks
    assert len(p_ref) == len(p_got)
    for a, b in zip(sorted(p_ref), sorted(p_got)):
        assert a.phase == b.phase and abs(a.peak_time - b.peak_time) <= 0.01 + 1e-9


def test_fft_resample_matches_obspy():
    """oracle.pipeline.fft_resample is a restatement from recollection of obspy.Trace.resample(window="hann", no_filter=True): with
    ObsPy present, pin it."""
    from oracle import pipeline

    rng = np.random.default_rng(0)
    for npts, rate in [(4001, 125.0), (4000, 125.0), (3000, 40.0), (2999, 66.6)]:
        x = rng.standard_normal(npts)
        tr = obspy.Trace(data=x.copy(), header={"sampling_rate": rate})
        tr.resample(100.0, no_filter=True)
        got = pipeline.fft_resample(x, rate, 100.0)
        assert got.shape == tr.data.shape
        assert float(np.abs(got - tr.data).max()) <= 1e-9 * float(np.abs(tr.data).max())
